// Lookup-table interpolants and pressure inference on the device.
//
// The reference turns a table (pressure x distance -> displacement) into two scipy LinearNDInterpolator objects
// (jointforces/simulation.py:595-621): piecewise-linear interpolation over the Delaunay triangulation of the
// scattered points (log distance, log pressure) resp. (log distance, displacement), NaN outside the hull.  It then
// evaluates the inverse one for every PIV vector of every frame (force.py:399-432, infer_pressure) and reduces
// the pressures of a frame to mean / median / std and 72 angular sector means (force.py:95-121, reconstruct).
//
// Here the triangulation is made once on the host by the same Qhull call (scipy.spatial.Delaunay, what
// LinearNDInterpolator does internally) and handed over as (simplices, transform, neighbors); the device does
//   * point location: triangles are bucketed by the intervals between the distinct abscissae of the points (a table
//     has one column of points per distance), a query tests the triangles of its interval in ascending index with
//     scipy's inside test (barycentric coordinates from `transform`, eps = 100 * DBL_EPSILON; second pass with
//     sqrt(eps) on the neighbours of degenerate simplices, as scipy's brute-force search does);
//   * interpolation: sum_j c_j * values[simplex_j], operation order of scipy's _evaluate_double, no FMA contraction;
//   * infer_pressure: projection, angle filter, table lookup, exp, fused in one kernel, arithmetic in numpy's order;
//   * per-frame statistics: fixed-order tree reductions, exact median by 8-pass radix selection.
// Values agree with scipy/numpy to rounding (1e-12 relative in the tests: libm vs CUDA log/exp/atan2 differ in the
// last bit, and a query on a shared edge may be assigned to the other triangle); order statistics are exact.
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "jf_internal.h"

struct jf_table {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t n_points = 0, n_simplices = 0;
    int n_buckets = 0, has_degenerate = 0;
    double *d_values = nullptr, *d_transform = nullptr, *d_bounds = nullptr;
    int32_t *d_simplices = nullptr, *d_bucket_ptr = nullptr, *d_bucket_tri = nullptr, *d_broad_ptr = nullptr, *d_broad_tri = nullptr;
    // scratch, grown on demand
    double *d_in = nullptr, *d_out = nullptr;
    size_t cap_in = 0, cap_out = 0;
    int32_t *d_flag = nullptr;  // keep flags, compaction offsets and the kept count
    size_t cap_flag = 0;
    double *d_frame = nullptr;
    size_t cap_frame = 0;
    int64_t launches = 0;
};

namespace {

struct TableDev {
    const double *values, *transform, *bounds;
    const int32_t *simplices, *bucket_ptr, *bucket_tri, *broad_ptr, *broad_tri;
    int n_buckets, has_degenerate;
};

constexpr double EPS = 100.0 * DBL_EPSILON;  // scipy.spatial.qhull: eps = 100 * DBL_EPSILON, eps_broad = sqrt(eps)

__device__ __forceinline__ bool inside(const double *T, double x, double y, double eps, double c[3]) {
    // _barycentric_inside: c_i = T[i][0] * (x - r_0) + T[i][1] * (y - r_1), c_2 = (1 - c_0) - c_1
    const double dx = __dsub_rn(x, T[4]), dy = __dsub_rn(y, T[5]);
    c[0] = __dadd_rn(__dmul_rn(T[0], dx), __dmul_rn(T[1], dy));
    if (!(-eps <= c[0] && c[0] <= 1.0 + eps)) return false;
    c[1] = __dadd_rn(__dmul_rn(T[2], dx), __dmul_rn(T[3], dy));
    if (!(-eps <= c[1] && c[1] <= 1.0 + eps)) return false;
    c[2] = __dsub_rn(__dsub_rn(1.0, c[0]), c[1]);
    return -eps <= c[2] && c[2] <= 1.0 + eps;
}

// piecewise-linear interpolant at (x, y); NaN outside the triangulation (and for NaN queries)
__device__ double table_lookup(const TableDev &t, double x, double y) {
    if (!(x == x) || !(y == y) || t.n_buckets < 1) return nan("");
    // interval k with bounds[k] <= x <= bounds[k+1]; queries beyond the ends test the end intervals (and fail there)
    int lo = 0, hi = t.n_buckets;  // upper_bound(bounds[1..n_buckets-1], x)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t.bounds[mid] <= x) lo = mid; else hi = mid;
    }
    double c[3];
    for (int pass = 0; pass < 1 + t.has_degenerate; pass++) {
        const int32_t *ptr = pass ? t.broad_ptr : t.bucket_ptr, *tri = pass ? t.broad_tri : t.bucket_tri;
        const double eps = pass ? sqrt(EPS) : EPS;
        for (int i = ptr[lo]; i < ptr[lo + 1]; i++) {
            const int s = tri[i];
            if (inside(t.transform + 6ll * s, x, y, eps, c)) {
                const int32_t *v = t.simplices + 3ll * s;
                double out = __dmul_rn(c[0], t.values[v[0]]);  // 0 + c_0 v_0
                out = __dadd_rn(out, __dmul_rn(c[1], t.values[v[1]]));
                out = __dadd_rn(out, __dmul_rn(c[2], t.values[v[2]]));
                return out;
            }
        }
    }
    return nan("");
}

__global__ void __launch_bounds__(256) jf_table_eval_kernel(TableDev t, long long n, const double *__restrict__ xq,
                                                            const double *__restrict__ yq, int log_x, int log_y, int exp_out,
                                                            double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = xq[i], y = yq[i];
    if (log_x) x = log(x);
    if (log_y) y = log(y);
    double r = table_lookup(t, x, y);
    if (exp_out) r = exp(r);
    out[i] = r;
}

// infer_pressure for the vectors of all frames (force.py:399-432).  in: x, y, u, v (concatenated frames);
// out: distance, displacement, angle, pressure per vector and keep = 1 where the reference keeps the vector
__global__ void __launch_bounds__(256) jf_infer_kernel(TableDev t, long long n, int n_frames, const long long *__restrict__ frame_ptr,
                                                       const double *__restrict__ fr /* (n_frames,3): cx, cy, r */,
                                                       const double *__restrict__ x, const double *__restrict__ y,
                                                       const double *__restrict__ u, const double *__restrict__ v, int use_filter,
                                                       double cos_min, double *__restrict__ o_dist, double *__restrict__ o_disp,
                                                       double *__restrict__ o_ang, double *__restrict__ o_p, int32_t *__restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = 0, hi = n_frames;  // frame of vector i
    while (hi - f > 1) {
        const int mid = (f + hi) >> 1;
        if (frame_ptr[mid] <= i) f = mid; else hi = mid;
    }
    const double cx = fr[3 * f], cy = fr[3 * f + 1], r = fr[3 * f + 2];
    const double ui = u[i], vi = v[i];
    const double dx = __ddiv_rn(__dsub_rn(x[i], cx), r), dy = __ddiv_rn(__dsub_rn(y[i], cy), r);
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    const double u2 = __ddiv_rn(ui, r), v2 = __ddiv_rn(vi, r);
    const double mx = __ddiv_rn(dx, dist), my = __ddiv_rn(dy, dist);
    const double disp = -__dadd_rn(__dmul_rn(u2, mx), __dmul_rn(v2, my));
    int k = ui == ui;  // ~isnan(u)
    if (k && use_filter) {
        const double ab = __dsqrt_rn(__dadd_rn(__dmul_rn(u2, u2), __dmul_rn(v2, v2)));
        k = __ddiv_rn(disp, ab) > cos_min;  // NaN compares false, as in numpy
    }
    keep[i] = k;
    o_dist[i] = dist;
    o_disp[i] = disp;
    o_ang[i] = atan2(dy, dx);
    o_p[i] = k ? exp(table_lookup(t, log(dist), disp)) : nan("");
}

// order-preserving compaction: one CTA scans the keep flags (n is a PIV field, not a mesh)
__global__ void __launch_bounds__(1024) jf_scan_kernel(long long n, const int32_t *__restrict__ keep, int32_t *__restrict__ pos,
                                                       int32_t *__restrict__ total) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const int k = i < n ? keep[i] : 0;
        int incl = k;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (warp ? warp_tot[warp - 1] : 0) + incl - k;
        if (i < n) pos[i] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) jf_compact_kernel(long long n, const int32_t *__restrict__ keep, const int32_t *__restrict__ pos,
                                                         const double *__restrict__ a0, const double *__restrict__ a1,
                                                         const double *__restrict__ a2, const double *__restrict__ a3,
                                                         double *__restrict__ out, long long stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const long long p = pos[i];
    out[p] = a0[i];
    out[stride + p] = a1[i];
    out[2 * stride + p] = a2[i];
    out[3 * stride + p] = a3[i];
}

// ---- per-frame statistics (force.py:95-121) --------------------------------------------------------------
constexpr int ST_THREADS = 512;
constexpr int N_SECTORS = 72;
constexpr int FRAME_OUT = 4 + N_SECTORS;  // mean, median, std(ddof=1), count of finite-or-inf (non-NaN) pressures, sector means

__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// fixed-order CTA sum of one double (and one count) per thread
__device__ double cta_sum(double v, double *sm) {
    __syncthreads();
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int w = ST_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
        __syncthreads();
    }
    return sm[0];
}

__global__ void __launch_bounds__(ST_THREADS) jf_frame_stats_kernel(const long long *__restrict__ frame_ptr, const double *__restrict__ dist,
                                                                    const double *__restrict__ ang, const double *__restrict__ pres,
                                                                    const int32_t *__restrict__ keep, int has_rmin, double r_min,
                                                                    int has_rmax, double r_max, const double *__restrict__ sector_lo,
                                                                    const double *__restrict__ sector_hi, double *__restrict__ out) {
    __shared__ double sm[ST_THREADS];
    __shared__ int hist[256];
    __shared__ int sel_bin, sel_k;
    const int f = blockIdx.x;
    const long long beg = frame_ptr[f], end = frame_ptr[f + 1];
    double *o = out + (long long)f * FRAME_OUT;
    // the reference's mask: distance > r_min (0 if None), and r_max > distance when r_max is given
    auto used = [&](long long i) -> bool {
        if (!keep[i]) return false;
        const double d = dist[i];
        bool m = d > (has_rmin ? r_min : 0.0);
        if (has_rmax) m = (r_max > d) && m;
        return m && pres[i] == pres[i];  // nan-aware statistics skip NaN pressures
    };
    double s = 0.0, cnt = 0.0;
    for (long long i = beg + threadIdx.x; i < end; i += ST_THREADS)
        if (used(i)) {
            s += pres[i];
            cnt += 1.0;
        }
    const double total = cta_sum(s, sm);
    const double count = cta_sum(cnt, sm);
    const double mean = count > 0.0 ? total / count : nan("");
    double ss = 0.0;
    for (long long i = beg + threadIdx.x; i < end; i += ST_THREADS)
        if (used(i)) {
            const double d = pres[i] - mean;
            ss += d * d;
        }
    const double ssum = cta_sum(ss, sm);
    // 72 sectors of +-5 degrees around alpha = -180, -175, ... 175 (each vector lies in two of them)
    for (int a = 0; a < N_SECTORS; a++) {
        const double lo = sector_lo[a], hi = sector_hi[a];
        double s2 = 0.0, c2 = 0.0;
        for (long long i = beg + threadIdx.x; i < end; i += ST_THREADS)
            if (used(i) && ang[i] >= lo && ang[i] < hi) {
                s2 += pres[i];
                c2 += 1.0;
            }
        const double st = cta_sum(s2, sm), ct = cta_sum(c2, sm);
        if (threadIdx.x == 0) o[4 + a] = ct > 0.0 ? st / ct : nan("");
    }
    // exact median: the middle order statistic(s) by radix selection on order-preserving keys
    const long long m = (long long)count;
    double med = nan("");
    if (m > 0) {
        double mid[2];
        for (int which = 0; which < 2; which++) {
            long long k = which ? m / 2 : (m - 1) / 2;
            if (which && k == (m - 1) / 2) {
                mid[1] = mid[0];
                break;
            }
            unsigned long long prefix = 0, mask = 0;
            for (int shift = 56; shift >= 0; shift -= 8) {
                for (int b = threadIdx.x; b < 256; b += ST_THREADS) hist[b] = 0;
                __syncthreads();
                for (long long i = beg + threadIdx.x; i < end; i += ST_THREADS)
                    if (used(i)) {
                        const unsigned long long key = order_key(pres[i]);
                        if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1);
                    }
                __syncthreads();
                if (threadIdx.x == 0) {
                    long long cum = 0;
                    int b = 0;
                    for (; b < 255; b++) {
                        if (k < cum + hist[b]) break;
                        cum += hist[b];
                    }
                    sel_bin = b;
                    sel_k = (int)(k - cum);
                }
                __syncthreads();
                prefix |= (unsigned long long)sel_bin << shift;
                mask |= 255ull << shift;
                k = sel_k;
                __syncthreads();
            }
            mid[which] = key_value(prefix);
        }
        med = mid[0] == mid[1] ? mid[0] : __dmul_rn(__dadd_rn(mid[0], mid[1]), 0.5);
    }
    if (threadIdx.x == 0) {
        o[0] = mean;
        o[1] = med;
        o[2] = count > 1.0 ? sqrt(ssum / (count - 1.0)) : nan("");
        o[3] = count;
    }
}

template <typename T>
int grow(jf_table *t, T **p, size_t *cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t n = need + need / 4 + 64;
    if (cudaMalloc((void **)p, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        jf_set_error("jf_table: out of device memory (%zu bytes)", n * sizeof(T));
        return JF_ERR_NOMEM;
    }
    *cap = n;
    return 0;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src) {
    const size_t n = src.size() ? src.size() : 1;
    if (cudaMalloc((void **)dst, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        jf_set_error("jf_table: out of device memory");
        return JF_ERR_NOMEM;
    }
    if (src.size()) JF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

TableDev dev_view(const jf_table *t) {
    TableDev d;
    d.values = t->d_values;
    d.transform = t->d_transform;
    d.bounds = t->d_bounds;
    d.simplices = t->d_simplices;
    d.bucket_ptr = t->d_bucket_ptr;
    d.bucket_tri = t->d_bucket_tri;
    d.broad_ptr = t->d_broad_ptr;
    d.broad_tri = t->d_broad_tri;
    d.n_buckets = t->n_buckets;
    d.has_degenerate = t->has_degenerate;
    return d;
}

}  // namespace

extern "C" {

void jf_table_destroy(jf_table *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    void *ptrs[] = {t->d_values, t->d_transform, t->d_bounds, t->d_simplices, t->d_bucket_ptr, t->d_bucket_tri, t->d_broad_ptr,
                    t->d_broad_tri, t->d_in, t->d_out, t->d_flag, t->d_frame};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

int jf_table_create(jf_table **out, int device, int64_t n_points, const double *points, const double *values, int64_t n_simplices,
                    const int32_t *simplices, const double *transform, const int32_t *neighbors) {
    if (!out || !points || !values || !simplices || !transform || !neighbors || n_points < 3 || n_simplices < 1 ||
        n_simplices > 0x3fffffff) {
        jf_set_error("jf_table_create: bad arguments");
        return JF_ERR_ARG;
    }
    *out = nullptr;
    for (int64_t s = 0; s < 3 * n_simplices; s++)
        if (simplices[s] < 0 || simplices[s] >= n_points || neighbors[s] < -1 || neighbors[s] >= n_simplices) {
            jf_set_error("jf_table_create: simplex %lld refers outside the point set", (long long)(s / 3));
            return JF_ERR_ARG;
        }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) {
        cudaGetLastError();
        jf_set_error("jf_table_create: no CUDA device (there is no CPU fallback)");
        return JF_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) {
        jf_set_error("jf_table_create: device %d out of range", device);
        return JF_ERR_ARG;
    }
    JF_CUDA(cudaSetDevice(device));
    // intervals between distinct abscissae (at most 1024 of them, by rank)
    std::vector<double> xs((size_t)n_points);
    for (int64_t i = 0; i < n_points; i++) xs[i] = points[2 * i];
    std::sort(xs.begin(), xs.end());
    xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
    if (!(xs.front() == xs.front()) || !(xs.back() == xs.back()) || xs.size() < 2) {
        jf_set_error("jf_table_create: points must be finite and not all on one abscissa");
        return JF_ERR_ARG;
    }
    std::vector<double> bounds;
    const size_t max_b = 1024;
    if (xs.size() - 1 <= max_b) bounds = xs;
    else
        for (size_t k = 0; k <= max_b; k++) bounds.push_back(xs[k * (xs.size() - 1) / max_b]);
    const int nb = (int)bounds.size() - 1;
    const double tol = 1e-9 * (bounds.back() - bounds.front());
    std::vector<uint8_t> degenerate((size_t)n_simplices, 0), near_degenerate((size_t)n_simplices, 0);
    int has_degenerate = 0;
    for (int64_t s = 0; s < n_simplices; s++)
        if (!(transform[6 * s] == transform[6 * s])) degenerate[s] = 1, has_degenerate = 1;
    if (has_degenerate)
        for (int64_t s = 0; s < n_simplices; s++)
            if (degenerate[s])
                for (int k = 0; k < 3; k++) {
                    const int32_t nbr = neighbors[3 * s + k];
                    if (nbr >= 0 && !degenerate[nbr]) near_degenerate[nbr] = 1;
                }
    std::vector<int32_t> ptr((size_t)nb + 1, 0), bptr((size_t)nb + 1, 0), tri, btri;
    std::vector<std::vector<int32_t>> lists((size_t)nb), blists((size_t)nb);
    for (int64_t s = 0; s < n_simplices; s++) {
        if (degenerate[s]) continue;
        double lo = INFINITY, hi = -INFINITY;
        for (int k = 0; k < 3; k++) {
            const double x = points[2 * (int64_t)simplices[3 * s + k]];
            lo = std::min(lo, x);
            hi = std::max(hi, x);
        }
        // near-degenerate triangles are tested with the broad tolerance, which reaches further out
        const double reach = near_degenerate[s] ? std::max(tol, 2.0 * sqrt(EPS) * (hi - lo)) : tol;
        int k0 = (int)(std::upper_bound(bounds.begin(), bounds.end(), lo - reach) - bounds.begin()) - 1;
        int k1 = (int)(std::lower_bound(bounds.begin(), bounds.end(), hi + reach) - bounds.begin());
        k0 = std::max(k0, 0);
        k1 = std::min(k1, nb);
        for (int k = k0; k < k1; k++) {
            lists[k].push_back((int32_t)s);
            if (near_degenerate[s]) blists[k].push_back((int32_t)s);
        }
    }
    for (int k = 0; k < nb; k++) {
        ptr[k + 1] = ptr[k] + (int32_t)lists[k].size();
        tri.insert(tri.end(), lists[k].begin(), lists[k].end());
        bptr[k + 1] = bptr[k] + (int32_t)blists[k].size();
        btri.insert(btri.end(), blists[k].begin(), blists[k].end());
    }
    jf_table *t = new jf_table();
    t->device = device;
    t->n_points = n_points;
    t->n_simplices = n_simplices;
    t->n_buckets = nb;
    t->has_degenerate = has_degenerate;
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "stream", __FILE__, __LINE__); break; }
        rc = upload(&t->d_values, std::vector<double>(values, values + n_points));
        if (rc) break;
        rc = upload(&t->d_transform, std::vector<double>(transform, transform + 6 * n_simplices));
        if (rc) break;
        rc = upload(&t->d_simplices, std::vector<int32_t>(simplices, simplices + 3 * n_simplices));
        if (rc) break;
        rc = upload(&t->d_bounds, bounds);
        if (rc) break;
        rc = upload(&t->d_bucket_ptr, ptr);
        if (rc) break;
        rc = upload(&t->d_bucket_tri, tri);
        if (rc) break;
        rc = upload(&t->d_broad_ptr, bptr);
        if (rc) break;
        rc = upload(&t->d_broad_tri, btri);
        if (rc) break;
    } while (0);
    if (rc) {
        jf_table_destroy(t);
        return rc;
    }
    *out = t;
    return JF_OK;
}

// out[i] = f(x_i, y_i), optionally on (log x, log y) and/or exponentiated: get_displacement is (log, log, -),
// get_pressure (log, -, exp) -- simulation.py:612-619
int jf_table_eval(jf_table *t, int64_t n, const double *xq, const double *yq, int log_x, int log_y, int exp_out, double *out) {
    if (!t || n < 0 || (n > 0 && (!xq || !yq || !out))) {
        jf_set_error("jf_table_eval: bad arguments");
        return JF_ERR_ARG;
    }
    if (n == 0) return JF_OK;
    JF_CUDA(cudaSetDevice(t->device));
    int rc;
    if ((rc = grow(t, &t->d_in, &t->cap_in, 2 * (size_t)n))) return rc;
    if ((rc = grow(t, &t->d_out, &t->cap_out, (size_t)n))) return rc;
    JF_CUDA(cudaMemcpyAsync(t->d_in, xq, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, t->stream));
    JF_CUDA(cudaMemcpyAsync(t->d_in + n, yq, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, t->stream));
    jf_table_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(dev_view(t), n, t->d_in, t->d_in + n, log_x, log_y, exp_out, t->d_out);
    JF_CUDA(cudaGetLastError());
    t->launches++;
    JF_CUDA(cudaMemcpyAsync(out, t->d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    JF_CUDA(cudaStreamSynchronize(t->stream));
    return JF_OK;
}

static int run_infer(jf_table *t, int n_frames, const int64_t *frame_ptr, const double *frames, const double *x, const double *y,
                     const double *u, const double *v, int use_filter, double cos_min) {
    const int64_t n = frame_ptr[n_frames];
    int rc;
    if ((rc = grow(t, &t->d_in, &t->cap_in, 4 * (size_t)n))) return rc;
    if ((rc = grow(t, &t->d_out, &t->cap_out, 8 * (size_t)n))) return rc;  // 4 per-vector results + 4 compacted
    if ((rc = grow(t, &t->d_flag, &t->cap_flag, 2 * (size_t)n + 2))) return rc;
    if ((rc = grow(t, &t->d_frame, &t->cap_frame, (size_t)n_frames * (3 + FRAME_OUT) + 2 * N_SECTORS + (size_t)n_frames + 1))) return rc;
    const double *src[4] = {x, y, u, v};
    for (int k = 0; k < 4; k++) JF_CUDA(cudaMemcpyAsync(t->d_in + k * n, src[k], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, t->stream));
    JF_CUDA(cudaMemcpyAsync(t->d_frame, frames, (size_t)n_frames * 3 * sizeof(double), cudaMemcpyHostToDevice, t->stream));
    // frame_ptr as 64-bit integers in the tail of the frame buffer
    long long *d_fp = (long long *)(t->d_frame + (size_t)n_frames * (3 + FRAME_OUT) + 2 * N_SECTORS);
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    JF_CUDA(cudaMemcpyAsync(d_fp, frame_ptr, ((size_t)n_frames + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, t->stream));
    jf_infer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(dev_view(t), n, n_frames, d_fp, t->d_frame, t->d_in, t->d_in + n,
                                                                        t->d_in + 2 * n, t->d_in + 3 * n, use_filter, cos_min, t->d_out,
                                                                        t->d_out + n, t->d_out + 2 * n, t->d_out + 3 * n, t->d_flag);
    JF_CUDA(cudaGetLastError());
    t->launches++;
    return JF_OK;
}

// infer_pressure (force.py:399-432) for one frame: the kept vectors' distance, displacement, angle, pressure, in
// input order (each output holds up to n values; *n_out says how many).  cos_min = cos(2 pi angle_filter / 360) as
// the caller computes it; use_filter = 0 is angle_filter=None.
int jf_infer_pressure(jf_table *t, int64_t n, const double *x, const double *y, const double *u, const double *v, double cx, double cy,
                      double r_sph, int use_filter, double cos_min, double *distance, double *displacement, double *angle,
                      double *pressure, int64_t *n_out) {
    if (!t || n < 0 || !n_out || (n > 0 && (!x || !y || !u || !v || !distance || !displacement || !angle || !pressure)) || n > 0x7fffffff) {
        jf_set_error("jf_infer_pressure: bad arguments");
        return JF_ERR_ARG;
    }
    *n_out = 0;
    if (n == 0) return JF_OK;
    JF_CUDA(cudaSetDevice(t->device));
    const int64_t fp[2] = {0, n};
    const double fr[3] = {cx, cy, r_sph};
    int rc = run_infer(t, 1, fp, fr, x, y, u, v, use_filter, cos_min);
    if (rc) return rc;
    int32_t *pos = t->d_flag + n, *total = t->d_flag + 2 * n;
    jf_scan_kernel<<<1, 1024, 0, t->stream>>>(n, t->d_flag, pos, total);
    jf_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(n, t->d_flag, pos, t->d_out, t->d_out + n, t->d_out + 2 * n,
                                                                          t->d_out + 3 * n, t->d_out + 4 * n, n);
    JF_CUDA(cudaGetLastError());
    t->launches += 2;
    int32_t kept = 0;
    JF_CUDA(cudaMemcpyAsync(&kept, total, sizeof(int32_t), cudaMemcpyDeviceToHost, t->stream));
    JF_CUDA(cudaStreamSynchronize(t->stream));
    double *dst[4] = {distance, displacement, angle, pressure};
    for (int k = 0; k < 4 && kept > 0; k++)
        JF_CUDA(cudaMemcpyAsync(dst[k], t->d_out + (4 + k) * n, (size_t)kept * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    JF_CUDA(cudaStreamSynchronize(t->stream));
    *n_out = kept;
    return JF_OK;
}

// The per-frame loop of reconstruct (force.py:70-121) for n_frames frames in one go: frame f owns the vectors
// frame_ptr[f] .. frame_ptr[f+1] of x, y, u, v and frames[f] = (cx, cy, r).  r_min / r_max: NaN stands for None.
// sector_bounds (72,2): the (alpha-5)*pi/180, (alpha+5)*pi/180 the caller computes (same doubles as the reference).
// stats (n_frames, 76): mean, median, std (ddof 1), number of pressures used, 72 sector means.
int jf_reconstruct_frames(jf_table *t, int32_t n_frames, const int64_t *frame_ptr, const double *frames, const double *x, const double *y,
                          const double *u, const double *v, int use_filter, double cos_min, double r_min, double r_max,
                          const double *sector_bounds, double *stats) {
    if (!t || n_frames < 1 || !frame_ptr || !frames || !sector_bounds || !stats || frame_ptr[0] != 0) {
        jf_set_error("jf_reconstruct_frames: bad arguments");
        return JF_ERR_ARG;
    }
    for (int f = 0; f < n_frames; f++)
        if (frame_ptr[f + 1] < frame_ptr[f]) {
            jf_set_error("jf_reconstruct_frames: frame_ptr must be non-decreasing");
            return JF_ERR_ARG;
        }
    const int64_t n = frame_ptr[n_frames];
    if (n > 0 && (!x || !y || !u || !v)) {
        jf_set_error("jf_reconstruct_frames: null vector arrays");
        return JF_ERR_ARG;
    }
    JF_CUDA(cudaSetDevice(t->device));
    int rc = run_infer(t, n_frames, frame_ptr, frames, x, y, u, v, use_filter, cos_min);
    if (rc) return rc;
    double *d_stats = t->d_frame + (size_t)n_frames * 3, *d_sec = d_stats + (size_t)n_frames * FRAME_OUT;
    long long *d_fp = (long long *)(d_sec + 2 * N_SECTORS);
    double lo_hi[2 * N_SECTORS];
    for (int a = 0; a < N_SECTORS; a++) {
        lo_hi[a] = sector_bounds[2 * a];
        lo_hi[N_SECTORS + a] = sector_bounds[2 * a + 1];
    }
    JF_CUDA(cudaMemcpyAsync(d_sec, lo_hi, sizeof(lo_hi), cudaMemcpyHostToDevice, t->stream));
    JF_CUDA(cudaStreamSynchronize(t->stream));  // lo_hi is a stack buffer
    jf_frame_stats_kernel<<<n_frames, ST_THREADS, 0, t->stream>>>(d_fp, t->d_out, t->d_out + 2 * n, t->d_out + 3 * n, t->d_flag,
                                                                  r_min == r_min, r_min, r_max == r_max, r_max, d_sec, d_sec + N_SECTORS, d_stats);
    JF_CUDA(cudaGetLastError());
    t->launches++;
    JF_CUDA(cudaMemcpyAsync(stats, d_stats, (size_t)n_frames * FRAME_OUT * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    JF_CUDA(cudaStreamSynchronize(t->stream));
    return JF_OK;
}

int64_t jf_table_launches(jf_table *t) { return t ? t->launches : 0; }

}  // extern "C"
