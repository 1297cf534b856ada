// One-off structure build for a context: locality reordering, shape tensors,
// SELL-32 block-matrix pattern with its deterministic gather lists (replaces saenopy's
// _compute_connections + per-iteration COO->CSR), and the folded direction table
// (SURVEY A2).  Host code; runs once per mesh, not on the hot path.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <queue>

#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <atomic>
#include <thread>

#include "jf_internal.h"

// ------------------------------------------------------------------------------------------------
// direction sphere, saenopy build_beams(300)
// ------------------------------------------------------------------------------------------------
static std::vector<double> default_beams(int n_req) {
    int n = (int)floor(sqrt(n_req * M_PI + 0.5));
    std::vector<double> b;
    for (int i = 0; i < n; i++) {
        double theta = (2 * M_PI / n) * i;
        int jmax = (int)floor(n * sin(theta) + 0.5);
        for (int j = 0; j < jmax; j++) {
            double rho = (2 * M_PI / jmax) * j;
            b.push_back(sin(theta) * cos(rho));
            b.push_back(sin(theta) * sin(rho));
            b.push_back(cos(theta));
        }
    }
    return b;
}

// index of the quartic monomial x^ex y^ey z^ez (ex+ey+ez = 4) in the 15-entry table
static inline int mono_id(int ex, int ey) { return (4 - ex) * (5 - ex) / 2 + (4 - ex - ey); }

static int build_dirtab(jf_ctx *c, const double *beams, int n_beams) {
    std::vector<double> own;
    if (!beams) {
        own = default_beams(300);
        beams = own.data();
        n_beams = (int)(own.size() / 3);
    }
    if (n_beams <= 0) {
        jf_set_error("n_beams must be positive");
        return JF_ERR_ARG;
    }
    c->n_beams = n_beams;
    // fold antipodal pairs: every term of A5-A8 is even in s, so s and -s contribute identically
    std::vector<int> rep;
    std::vector<double> wt;
    std::vector<char> used(n_beams, 0);
    for (int i = 0; i < n_beams; i++) {
        if (used[i]) continue;
        used[i] = 1;
        double w = 1.0;
        if (!(c->flags & JF_FLAG_NO_FOLD)) {
            for (int j = i + 1; j < n_beams; j++) {
                if (used[j]) continue;
                double d = 0;
                for (int k = 0; k < 3; k++) {
                    double e = beams[3 * i + k] + beams[3 * j + k];
                    d += e * e;
                }
                if (d < 1e-24) {
                    used[j] = 1;
                    w = 2.0;
                    break;
                }
            }
        }
        rep.push_back(i);
        wt.push_back(w);
    }
    c->n_dirs = (int)rep.size();
    int quantum = 16;  // lanes (1, 2, 4 or 8) x unroll 2
    c->n_dirs_pad = (c->n_dirs + quantum - 1) / quantum * quantum;
    // layout: (JF_NTAB/2) planes of n_dirs_pad double2
    std::vector<double> tab((size_t)JF_NTAB * c->n_dirs_pad, 0.0);
    auto put = [&](int comp, int b, double v) { tab[((size_t)(comp >> 1) * c->n_dirs_pad + b) * 2 + (comp & 1)] = v; };
    for (int b = 0; b < c->n_dirs_pad; b++) {
        double s[3] = {1.0, 0.0, 0.0}, w = 0.0;  // padding: harmless direction with zero weight
        if (b < c->n_dirs) {
            for (int k = 0; k < 3; k++) s[k] = beams[3 * rep[b] + k];
            w = wt[b];
        }
        put(0, b, s[0] * s[0]);
        put(1, b, s[1] * s[1]);
        put(2, b, s[2] * s[2]);
        put(3, b, s[0] * s[1]);
        put(4, b, s[0] * s[2]);
        put(5, b, s[1] * s[2]);
        for (int ex = 4; ex >= 0; ex--)
            for (int ey = 4 - ex; ey >= 0; ey--) {
                int ez = 4 - ex - ey;
                double v = 1.0;
                for (int q = 0; q < ex; q++) v *= s[0];
                for (int q = 0; q < ey; q++) v *= s[1];
                for (int q = 0; q < ez; q++) v *= s[2];
                put(6 + mono_id(ex, ey), b, v);
            }
        put(21, b, w);
    }
    if (jf_dev_alloc(c, &c->d_dirtab, tab.size())) return JF_ERR_CUDA;
    JF_CUDA(cudaMemcpy(c->d_dirtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(tab.size() * sizeof(double));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// node adjacency (sorted, unique, including self)
// ------------------------------------------------------------------------------------------------
// node -> incident (tet, corner) pairs, encoded t*4+m, ascending.  Threads own node ranges and scan all corners
// for them: no scatter conflicts, same result for any thread count.
static void build_incidence(int64_t N, int64_t T, const int32_t *tets, std::vector<int32_t> &iptr, std::vector<int32_t> &isrc) {
    iptr.assign(N + 1, 0);
    jf_parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = 0; i < 4 * T; i++) {
            const int32_t n = tets[i];
            if (n >= lo && n < hi) iptr[n + 1]++;
        }
    });
    for (int64_t i = 0; i < N; i++) iptr[i + 1] += iptr[i];
    isrc.resize((size_t)4 * T);
    jf_parallel_for(N, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> fill(iptr.begin() + lo, iptr.begin() + hi);
        for (int64_t i = 0; i < 4 * T; i++) {
            const int32_t n = tets[i];
            if (n >= lo && n < hi) isrc[fill[n - lo]++] = (int32_t)i;
        }
    });
}

static void build_adjacency(int64_t N, int64_t T, const int32_t *tets, std::vector<int64_t> &ptr,
                            std::vector<int32_t> &adj) {
    std::vector<int32_t> iptr, isrc;
    build_incidence(N, T, tets, iptr, isrc);
    // neighbours of a node (itself included): the corners of its incident tets, sorted, distinct
    std::vector<int32_t> raw((size_t)4 * iptr[N]);
    ptr.assign(N + 1, 0);
    jf_parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            int32_t *a = &raw[4 * (size_t)iptr[i]], *w = a;
            for (int32_t k = iptr[i]; k < iptr[i + 1]; k++) {
                const int32_t *tt = &tets[4 * (int64_t)(isrc[k] >> 2)];
                for (int r = 0; r < 4; r++) *w++ = tt[r];
            }
            std::sort(a, w);
            ptr[i + 1] = std::unique(a, w) - a;
        }
    });
    for (int64_t i = 0; i < N; i++) ptr[i + 1] += ptr[i];
    adj.resize((size_t)ptr[N]);
    jf_parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++)
            std::copy(raw.begin() + 4 * (size_t)iptr[i], raw.begin() + 4 * (size_t)iptr[i] + (ptr[i + 1] - ptr[i]), adj.begin() + ptr[i]);
    });
}

// Cost of the SpMV's vector gathers for a candidate ordering of the free nodes: the number of
// distinct 128-byte lines touched by each 32-lane gather (one per slice and block column), summed.
// On B200 that is what bounds the CG kernel for these matrices (L1tex wavefronts; DESIGN.md
// "CG"), so the ordering is chosen by it rather than by bandwidth alone.  Also returns the padded
// slot count.
static double gather_cost(const std::vector<int32_t> &order, int64_t N, const std::vector<int64_t> &ptr,
                          const std::vector<int32_t> &adj, int64_t *slots_out) {
    const int64_t nf = (int64_t)order.size();
    std::vector<int32_t> o2n(N, -1);
    for (int64_t i = 0; i < nf; i++) o2n[order[i]] = (int32_t)i;
    const int64_t n_slices = (nf + JF_SLICE - 1) / JF_SLICE;
    std::vector<double> part_lines(64, 0.0);
    std::vector<int64_t> part_slots(64, 0);
    std::atomic<int> part_id{0};
    jf_parallel_for(n_slices, [&](int64_t sl_lo, int64_t sl_hi) {
        std::vector<int32_t> cols[JF_SLICE];
        int32_t tmp[JF_SLICE];
        double lines = 0;
        int64_t slots = 0;
        for (int64_t sl = sl_lo; sl < sl_hi; sl++) {
            const int64_t s0 = sl * JF_SLICE;
            size_t w = 0;
            const int nrow = (int)std::min<int64_t>(JF_SLICE, nf - s0);
            for (int r = 0; r < nrow; r++) {
                cols[r].clear();
                const int32_t old = order[s0 + r];
                for (int64_t k = ptr[old]; k < ptr[old + 1]; k++)
                    if (o2n[adj[k]] >= 0) cols[r].push_back(o2n[adj[k]]);
                std::sort(cols[r].begin(), cols[r].end());
                w = std::max(w, cols[r].size());
            }
            slots += (int64_t)w * JF_SLICE;
            for (size_t j = 0; j < w; j++) {
                int n = 0;
                for (int r = 0; r < nrow; r++)
                    if (j < cols[r].size()) tmp[n++] = cols[r][j] >> 2;  // four 32-byte entries per line
                std::sort(tmp, tmp + n);
                lines += (double)(std::unique(tmp, tmp + n) - tmp);
            }
        }
        const int id = part_id.fetch_add(1);
        part_lines[id] = lines;
        part_slots[id] = slots;
    });
    double lines = 0;
    int64_t slots = 0;
    for (int i = 0; i < 64; i++) {
        lines += part_lines[i];  // integer-valued: exact in any order
        slots += part_slots[i];
    }
    if (slots_out) *slots_out = slots;
    return lines;
}

// ------------------------------------------------------------------------------------------------
// adjacency and ordering scores on the device
// ------------------------------------------------------------------------------------------------
constexpr int ADJ_CAP = 128;  // distinct neighbours per node the device path handles (else: host path)

__global__ void __launch_bounds__(256) jf_inc_count_kernel(long long n4, const int32_t *__restrict__ tets, int32_t *__restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) atomicAdd(&cnt[tets[i] + 1], 1);
}

__global__ void __launch_bounds__(256) jf_inc_fill_kernel(long long n4, const int32_t *__restrict__ tets, int32_t *__restrict__ cur,
                                                          int32_t *__restrict__ isrc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) isrc[atomicAdd(&cur[tets[i]], 1)] = (int32_t)i;  // order within a node is irrelevant: the consumer sorts
}

// neighbours of node i (itself included): corners of its incident tets, sorted, distinct.
// FILL = false: nd[i + 1] = count; FILL = true: adj[aptr[i] ..] = list
template <bool FILL>
__global__ void __launch_bounds__(128) jf_adj_kernel(long long N, const int32_t *__restrict__ tets, const int32_t *__restrict__ iptr,
                                                     const int32_t *__restrict__ isrc, int32_t *__restrict__ nd_or_aptr,
                                                     int32_t *__restrict__ adj, int *__restrict__ overflow) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int32_t list[ADJ_CAP];
    int n = 0;
    for (int32_t k = iptr[i]; k < iptr[i + 1]; k++) {
        const int32_t *tt = tets + 4ll * (isrc[k] >> 2);
        for (int r = 0; r < 4; r++) {
            const int32_t v = tt[r];
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (list[mid] < v) lo = mid + 1; else hi = mid;
            }
            if (lo < n && list[lo] == v) continue;
            if (n == ADJ_CAP) {
                *overflow = 1;
                continue;
            }
            for (int q = n; q > lo; q--) list[q] = list[q - 1];
            list[lo] = v;
            n++;
        }
    }
    if (FILL) {
        int32_t *out = adj + nd_or_aptr[i];
        for (int q = 0; q < n; q++) out[q] = list[q];
    } else {
        nd_or_aptr[i + 1] = n;
    }
}

// gather_cost on the device: one warp per slice of 32 rows of a candidate ordering; out[0] += distinct 128-byte
// lines over all block columns, out[1] += padded slots.  Integer sums: the same numbers as the host function.
__global__ void __launch_bounds__(128) jf_score_kernel(long long nf, const int32_t *__restrict__ order, const int32_t *__restrict__ o2n,
                                                       const int32_t *__restrict__ aptr, const int32_t *__restrict__ adj,
                                                       unsigned long long *__restrict__ out) {
    const long long sl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (sl * JF_SLICE >= nf) return;
    const long long row = sl * JF_SLICE + lane;
    int32_t cols[ADJ_CAP];
    int n = 0;
    if (row < nf) {
        const int32_t old = order[row];
        for (int32_t k = aptr[old]; k < aptr[old + 1]; k++) {
            const int32_t v = o2n[adj[k]];
            if (v < 0) continue;
            int q = n++;
            for (; q > 0 && cols[q - 1] > v; q--) cols[q] = cols[q - 1];
            cols[q] = v;
        }
    }
    int w = n;
    for (int d = 16; d > 0; d >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, d));
    unsigned long long lines = 0;
    for (int j = 0; j < w; j++) {
        const bool valid = j < n;
        const int key = valid ? (cols[j] >> 2) : -1 - lane;  // four 32-byte entries per line
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const bool leader = lane == __ffs(m) - 1;
        lines += __popc(__ballot_sync(0xffffffffu, leader && valid));
    }
    if (lane == 0) {
        atomicAdd(&out[0], lines);
        atomicAdd(&out[1], (unsigned long long)w * JF_SLICE);
    }
}

// in-place inclusive prefix sum of n int32 values by one CTA (setup only)
__global__ void __launch_bounds__(1024) jf_prefix_kernel(long long n, int32_t *__restrict__ v) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        int incl = i < n ? v[i] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int x = warp_tot[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += o;
            }
            warp_tot[lane] = x;
        }
        __syncthreads();
        if (i < n) v[i] = carry + (warp ? warp_tot[warp - 1] : 0) + incl;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
}

// temporaries of the device-side setup, released by the caller
struct SetupDev {
    int32_t *tets_in = nullptr, *aptr = nullptr, *adj = nullptr;
    unsigned long long *score = nullptr;  // [4]: lines, slots of the caller's order; lines, slots of Cuthill-McKee
    bool have = false;
};

// adjacency on the device; host copies in ptr/adj.  Returns 1 if done, 0 if the host path must be used, < 0 on error.
static int build_adjacency_device(jf_ctx *c, int64_t N, int64_t T, const int32_t *tets, std::vector<int64_t> &ptr, std::vector<int32_t> &adj,
                                  SetupDev &d) {
    if (T < 1 || getenv("JF_HOST_SETUP")) return 0;
    const long long n4 = 4 * T;
    int32_t *iptr = nullptr, *cur = nullptr, *isrc = nullptr;
    int *overflow = nullptr;
    int rc;
    if ((rc = jf_dev_alloc(c, &d.tets_in, (size_t)n4))) return rc;
    if ((rc = jf_dev_alloc(c, &iptr, (size_t)N + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &cur, (size_t)N + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &isrc, (size_t)n4))) return rc;
    if ((rc = jf_dev_alloc(c, &d.aptr, (size_t)N + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &overflow, 1))) return rc;
    if ((rc = jf_dev_alloc(c, &d.score, 4))) return rc;
    JF_CUDA(cudaMemcpyAsync(d.tets_in, tets, (size_t)n4 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    c->stats.bytes_h2d += (int64_t)((size_t)n4 * sizeof(int32_t));
    JF_CUDA(cudaMemsetAsync(iptr, 0, ((size_t)N + 1) * sizeof(int32_t), c->stream));
    JF_CUDA(cudaMemsetAsync(d.aptr, 0, ((size_t)N + 1) * sizeof(int32_t), c->stream));
    JF_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), c->stream));
    JF_CUDA(cudaMemsetAsync(d.score, 0, 4 * sizeof(unsigned long long), c->stream));
    const unsigned g4 = (unsigned)((n4 + 255) / 256), gn = (unsigned)((N + 127) / 128);
    jf_inc_count_kernel<<<g4, 256, 0, c->stream>>>(n4, d.tets_in, iptr);
    jf_prefix_kernel<<<1, 1024, 0, c->stream>>>(N + 1, iptr);
    JF_CUDA(cudaMemcpyAsync(cur, iptr, ((size_t)N + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    jf_inc_fill_kernel<<<g4, 256, 0, c->stream>>>(n4, d.tets_in, cur, isrc);
    jf_adj_kernel<false><<<gn, 128, 0, c->stream>>>(N, d.tets_in, iptr, isrc, d.aptr, nullptr, overflow);
    jf_prefix_kernel<<<1, 1024, 0, c->stream>>>(N + 1, d.aptr);
    JF_CUDA(cudaGetLastError());
    std::vector<int32_t> aptr32((size_t)N + 1);
    int over = 0;
    JF_CUDA(cudaMemcpyAsync(aptr32.data(), d.aptr, ((size_t)N + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaMemcpyAsync(&over, overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    bool ok = !over;
    if (ok) {
        adj.resize((size_t)aptr32[N]);
        if ((rc = jf_dev_alloc(c, &d.adj, adj.size()))) return rc;
        jf_adj_kernel<true><<<gn, 128, 0, c->stream>>>(N, d.tets_in, iptr, isrc, d.aptr, d.adj, overflow);
        JF_CUDA(cudaGetLastError());
        JF_CUDA(cudaMemcpyAsync(adj.data(), d.adj, adj.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        JF_CUDA(cudaStreamSynchronize(c->stream));
        c->stats.bytes_d2h += (int64_t)((adj.size() + aptr32.size()) * sizeof(int32_t));
        ptr.assign(aptr32.begin(), aptr32.end());
        d.have = true;
    }
    for (void *tmp_buf : {(void *)iptr, (void *)cur, (void *)isrc, (void *)overflow}) cudaFreeAsync(tmp_buf, c->stream);
    return ok ? 1 : 0;
}

static void release_setup_dev(jf_ctx *c, SetupDev &d) {
    for (void *tmp_buf : {(void *)d.tets_in, (void *)d.aptr, (void *)d.adj, (void *)d.score})
        if (tmp_buf) cudaFreeAsync(tmp_buf, c->stream);
    d = SetupDev();
}

// launch the scoring of one candidate ordering (slot 0: the caller's order, 1: Cuthill-McKee); temporaries are
// stream-ordered, so they can be released right after the launch
static int score_on_device(jf_ctx *c, SetupDev &d, int which, const std::vector<int32_t> &order, int64_t N) {
    std::vector<int32_t> o2n((size_t)N, -1);
    for (size_t i = 0; i < order.size(); i++) o2n[order[i]] = (int32_t)i;
    int32_t *d_order = nullptr, *d_o2n = nullptr;
    int rc;
    if ((rc = jf_dev_alloc(c, &d_order, order.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &d_o2n, (size_t)N))) return rc;
    JF_CUDA(cudaMemcpyAsync(d_order, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    JF_CUDA(cudaMemcpyAsync(d_o2n, o2n.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    const long long nf = (long long)order.size();
    const long long n_slices = (nf + JF_SLICE - 1) / JF_SLICE;
    if (n_slices > 0) {
        jf_score_kernel<<<(unsigned)((n_slices + 3) / 4), 128, 0, c->stream>>>(nf, d_order, d_o2n, d.aptr, d.adj, d.score + 2 * which);
        JF_CUDA(cudaGetLastError());
    }
    cudaFreeAsync(d_order, c->stream);
    cudaFreeAsync(d_o2n, c->stream);
    return 0;
}

// Candidate orderings of the free nodes: the caller's order, and Cuthill-McKee (breadth-first,
// neighbours by increasing degree); the one with the cheapest gathers wins.
static int order_nodes(jf_ctx *c, const uint8_t *movable, const std::vector<int64_t> &ptr,
                       const std::vector<int32_t> &adj, SetupDev &dev, StageTimer *tm = nullptr) {
    int64_t N = c->N;
    std::vector<int32_t> order;
    order.reserve(N);
    for (int64_t i = 0; i < N; i++)
        if (movable[i]) order.push_back((int32_t)i);
    if (!(c->flags & JF_FLAG_NO_REORDER)) {
        int64_t slots_id = 0, slots_cm = 0;
        double cost_id = 0, cost_cm = 0;
        // the caller's order is scored (device, or a host thread) while Cuthill-McKee is being built
        std::thread score_id;
        if (dev.have) {
            int rc = score_on_device(c, dev, 0, order, N);
            if (rc) return rc;
        } else {
            score_id = std::thread([&] { cost_id = gather_cost(order, N, ptr, adj, &slots_id); });
        }
        std::vector<int32_t> cm;
        cm.reserve(order.size());
        std::vector<int32_t> deg(N, 0);
        for (int64_t i = 0; i < N; i++)
            if (movable[i])
                for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) deg[i] += movable[adj[k]] ? 1 : 0;
        std::vector<char> seen(N, 0);
        // seeds in order of increasing degree (first unvisited wins)
        std::vector<int32_t> seeds(order);
        std::stable_sort(seeds.begin(), seeds.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
        std::vector<int32_t> nb;
        for (int32_t seed : seeds) {
            if (seen[seed]) continue;
            seen[seed] = 1;
            size_t head = cm.size();
            cm.push_back(seed);
            while (head < cm.size()) {
                int32_t u = cm[head++];
                nb.clear();
                for (int64_t k = ptr[u]; k < ptr[u + 1]; k++) {
                    int32_t v = adj[k];
                    if (movable[v] && !seen[v]) {
                        seen[v] = 1;
                        nb.push_back(v);
                    }
                }
                std::stable_sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
                cm.insert(cm.end(), nb.begin(), nb.end());
            }
        }
        if (tm) tm->mark("ordering: cuthill-mckee");
        if (dev.have) {
            int rc = score_on_device(c, dev, 1, cm, N);
            if (rc) return rc;
            unsigned long long sc[4];
            JF_CUDA(cudaMemcpyAsync(sc, dev.score, sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
            JF_CUDA(cudaStreamSynchronize(c->stream));
            cost_id = (double)sc[0];
            slots_id = (int64_t)sc[1];
            cost_cm = (double)sc[2];
            slots_cm = (int64_t)sc[3];
        } else {
            score_id.join();
            cost_cm = gather_cost(cm, N, ptr, adj, &slots_cm);
        }
        if (tm) tm->mark("ordering: scores");
        // a gathered line and ~2 padded block slots cost about the same L1tex time
        if (cost_cm + 0.5 * (double)slots_cm / JF_SLICE < cost_id + 0.5 * (double)slots_id / JF_SLICE) order.swap(cm);
    }
    c->Nf = (int64_t)order.size();
    for (int64_t i = 0; i < N; i++)
        if (!movable[i]) order.push_back((int32_t)i);
    c->node_new2old = order;
    c->node_old2new.assign(N, 0);
    for (int64_t i = 0; i < N; i++) c->node_old2new[order[i]] = (int32_t)i;
    return 0;
}

template <typename Tp>
static int upload(jf_ctx *c, Tp **dst, const std::vector<Tp> &src) {
    if (jf_dev_alloc(c, dst, src.size())) return JF_ERR_CUDA;
    if (!src.empty()) JF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(src.size() * sizeof(Tp));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// structure pieces built on the device (they are per-tet / per-row independent and were the longest host stages)
// ------------------------------------------------------------------------------------------------
// shape tensors Phi = Chi * B^-1, V = |det B|/6  (SURVEY A3); plain IEEE operations in the order of the CPU
// restatement (no FMA contraction), first degenerate tet reported through *bad (old tet index, -1 if none)
__global__ void __launch_bounds__(128) jf_shape_kernel(long long T, const int32_t *__restrict__ tets, const double *__restrict__ xyz,
                                                       const int32_t *__restrict__ tet_new2old, double *__restrict__ phi,
                                                       double *__restrict__ vol, int *__restrict__ bad) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int32_t *n = tets + 4 * t;
    double B[3][3];
    for (int col = 0; col < 3; col++)
        for (int i = 0; i < 3; i++) B[i][col] = __dsub_rn(xyz[3ll * n[col + 1] + i], xyz[3ll * n[0] + i]);
#define MUL(a, b) __dmul_rn(a, b)
#define SUB(a, b) __dsub_rn(a, b)
#define MINOR(a, b, c, d) SUB(MUL(a, b), MUL(c, d))
    const double det = __dadd_rn(SUB(MUL(B[0][0], MINOR(B[1][1], B[2][2], B[1][2], B[2][1])), MUL(B[0][1], MINOR(B[1][0], B[2][2], B[1][2], B[2][0]))),
                                 MUL(B[0][2], MINOR(B[1][0], B[2][1], B[1][1], B[2][0])));
    if (det == 0.0 || !isfinite(det)) {
        atomicMax(bad, (int)tet_new2old[t]);  // any one of them will do for the message
        return;
    }
    double inv[3][3];
    inv[0][0] = __ddiv_rn(MINOR(B[1][1], B[2][2], B[1][2], B[2][1]), det);
    inv[0][1] = __ddiv_rn(MINOR(B[0][2], B[2][1], B[0][1], B[2][2]), det);
    inv[0][2] = __ddiv_rn(MINOR(B[0][1], B[1][2], B[0][2], B[1][1]), det);
    inv[1][0] = __ddiv_rn(MINOR(B[1][2], B[2][0], B[1][0], B[2][2]), det);
    inv[1][1] = __ddiv_rn(MINOR(B[0][0], B[2][2], B[0][2], B[2][0]), det);
    inv[1][2] = __ddiv_rn(MINOR(B[0][2], B[1][0], B[0][0], B[1][2]), det);
    inv[2][0] = __ddiv_rn(MINOR(B[1][0], B[2][1], B[1][1], B[2][0]), det);
    inv[2][1] = __ddiv_rn(MINOR(B[0][1], B[2][0], B[0][0], B[2][1]), det);
    inv[2][2] = __ddiv_rn(MINOR(B[0][0], B[1][1], B[0][1], B[1][0]), det);
#undef MINOR
#undef SUB
#undef MUL
    vol[t] = __ddiv_rn(fabs(det), 6.0);
    double *P = phi + 12 * t;
    for (int j = 0; j < 3; j++) {
        P[j] = -__dadd_rn(__dadd_rn(inv[0][j], inv[1][j]), inv[2][j]);
        P[3 + j] = inv[0][j];
        P[6 + j] = inv[1][j];
        P[9 + j] = inv[2][j];
    }
}

// Gather lists of the stiffness: slot <- ((t*10+pair)*2+transpose), ascending t (deterministic summation order).
// One thread per free row a: its contributions come from its incident tets (the force gather list, ascending t*4+m);
// the j-th block of the row lives in slot slice_ptr[a/32] + 32 j + a%32, touched by this thread only.
// FILL = false counts into cnt[slot + 1]; FILL = true writes the sources at cur[slot]++ (cur starts as the prefix sum).
template <bool FILL>
__global__ void __launch_bounds__(128) jf_contrib_kernel(long long Nf, const int32_t *__restrict__ tets, const int32_t *__restrict__ fptr,
                                                         const int32_t *__restrict__ fsrc, const int32_t *__restrict__ rptr,
                                                         const int32_t *__restrict__ rcol, const int32_t *__restrict__ slice_ptr,
                                                         int32_t *__restrict__ cnt_or_cur, int32_t *__restrict__ csrc) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= Nf) return;
    const int pair_of[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
    const int32_t *cb = rcol + rptr[a];
    const int w = rptr[a + 1] - rptr[a];
    const long long base = (long long)slice_ptr[a / JF_SLICE] + a % JF_SLICE;
    for (int32_t k = fptr[a]; k < fptr[a + 1]; k++) {
        const long long t = fsrc[k] >> 2;
        const int m = fsrc[k] & 3;
        for (int r = 0; r < 4; r++) {
            const int32_t b = tets[4 * t + r];
            if (b >= Nf) continue;
            int lo = 0, hi = w;  // lower_bound in the row's sorted block columns
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cb[mid] < b) lo = mid + 1; else hi = mid;
            }
            const long long slot = base + (long long)lo * JF_SLICE;
            if (FILL) csrc[cnt_or_cur[slot]++] = (int32_t)(((t * 10 + pair_of[m][r]) << 1) | (m > r));
            else cnt_or_cur[slot + 1]++;
        }
    }
}

// stream-ordered upload (pageable source: staged before the call returns, ordered before later work on the stream)
template <typename Tp>
static int upload_async(jf_ctx *c, Tp **dst, const std::vector<Tp> &src) {
    if (jf_dev_alloc(c, dst, src.size())) return JF_ERR_CUDA;
    if (!src.empty()) JF_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice, c->stream));
    c->stats.bytes_h2d += (int64_t)(src.size() * sizeof(Tp));
    return 0;
}

int jf_build_structure(jf_ctx *c, const double *nodes, const int32_t *tets_in, const uint8_t *movable,
                       const double *beams, int n_beams) {
    const int64_t N = c->N, T = c->T;
    for (int64_t i = 0; i < 4 * T; i++)
        if (tets_in[i] < 0 || tets_in[i] >= N) {
            jf_set_error("tetrahedron %lld references node %d outside [0,%lld)", (long long)(i / 4), tets_in[i], (long long)N);
            return JF_ERR_ARG;
        }
    StageTimer tm;
    int rc = build_dirtab(c, beams, n_beams);
    if (rc) return rc;

    std::vector<int64_t> aptr;
    std::vector<int32_t> adj;
    tm.mark("validate+dirtab");
    SetupDev dev;
    rc = build_adjacency_device(c, N, T, tets_in, aptr, adj, dev);
    if (rc < 0) return rc;
    if (rc == 0) build_adjacency(N, T, tets_in, aptr, adj);
    tm.mark(dev.have ? "adjacency (device)" : "adjacency (host)");
    rc = order_nodes(c, movable, aptr, adj, dev, &tm);
    release_setup_dev(c, dev);
    if (rc) return rc;
    tm.mark("ordering");
    const int64_t Nf = c->Nf;
    const std::vector<int32_t> &o2n = c->node_old2new;

    // tets: renumber, order by smallest new node id (keeps U gathers and K_el reads local)
    std::vector<int32_t> torder(T);
    if (!(c->flags & JF_FLAG_NO_REORDER)) {
        std::vector<int32_t> key(T);
        jf_parallel_for(T, [&](int64_t lo, int64_t hi) {
            for (int64_t t = lo; t < hi; t++) {
                int32_t k = o2n[tets_in[4 * t]];
                for (int m = 1; m < 4; m++) k = std::min(k, o2n[tets_in[4 * t + m]]);
                key[t] = k;
            }
        });
        // stable counting sort (keys are node ids)
        std::vector<int32_t> start(N + 1, 0);
        for (int64_t t = 0; t < T; t++) start[key[t] + 1]++;
        for (int64_t i = 0; i < N; i++) start[i + 1] += start[i];
        for (int64_t t = 0; t < T; t++) torder[start[key[t]]++] = (int32_t)t;
    } else {
        std::iota(torder.begin(), torder.end(), 0);
    }
    c->tet_new2old = torder;
    std::vector<int32_t> tets((size_t)4 * T);
    jf_parallel_for(T, [&](int64_t lo, int64_t hi) {
        for (int64_t t = lo; t < hi; t++)
            for (int m = 0; m < 4; m++) tets[4 * t + m] = o2n[tets_in[4 * (int64_t)torder[t] + m]];
    });
    std::vector<double> xyz((size_t)3 * N);
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < 3; k++) xyz[3 * i + k] = nodes[3 * (int64_t)c->node_new2old[i] + k];

    tm.mark("renumber tets");
    // shape tensors on the device while the host goes on with the sparsity pattern
    int *d_bad = nullptr;
    int32_t *d_torder = nullptr;
    if ((rc = upload_async(c, &c->d_tets, tets))) return rc;
    if ((rc = upload_async(c, &c->d_nodes, xyz))) return rc;
    if ((rc = upload_async(c, &d_torder, torder))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_phi, (size_t)12 * T))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_vol, (size_t)T))) return rc;
    if ((rc = jf_dev_alloc(c, &d_bad, 1))) return rc;
    JF_CUDA(cudaMemsetAsync(d_bad, 0xff, sizeof(int), c->stream));  // -1
    if (T > 0) {
        jf_shape_kernel<<<(unsigned)((T + 127) / 128), 128, 0, c->stream>>>(T, c->d_tets, c->d_nodes, d_torder, c->d_phi, c->d_vol, d_bad);
        JF_CUDA(cudaGetLastError());
    }
    tm.mark("shape tensors (device)");
    // SELL-32 pattern over free rows / free columns, in new numbering
    c->n_slices = (Nf + JF_SLICE - 1) / JF_SLICE;
    c->Nf_pad = c->n_slices * JF_SLICE;
    std::vector<int64_t> rptr(Nf + 1, 0);   // free-neighbour lists in new ids, sorted
    std::vector<int32_t> rcol;
    rcol.reserve((size_t)aptr[N]);
    {
        std::vector<int32_t> tmp;
        for (int64_t i = 0; i < Nf; i++) {
            int32_t old = c->node_new2old[i];
            tmp.clear();
            for (int64_t k = aptr[old]; k < aptr[old + 1]; k++) {
                int32_t v = o2n[adj[k]];
                if (v < Nf) tmp.push_back(v);
            }
            std::sort(tmp.begin(), tmp.end());
            rcol.insert(rcol.end(), tmp.begin(), tmp.end());
            rptr[i + 1] = (int64_t)rcol.size();
        }
    }
    c->n_blocks = (int64_t)rcol.size();
    std::vector<int32_t> slice_ptr(c->n_slices + 1, 0);
    for (int64_t k = 0; k < c->n_slices; k++) {
        int64_t w = 0;
        for (int64_t i = k * JF_SLICE; i < std::min<int64_t>(Nf, (k + 1) * JF_SLICE); i++) w = std::max(w, rptr[i + 1] - rptr[i]);
        int64_t next = (int64_t)slice_ptr[k] + w * JF_SLICE;
        if (next * 9 > 2000000000LL) {
            jf_set_error("mesh too large for 32-bit slot offsets");
            return JF_ERR_ARG;
        }
        slice_ptr[k + 1] = (int32_t)next;
    }
    c->n_slots = slice_ptr[c->n_slices];
    std::vector<int32_t> sell_col((size_t)c->n_slots);
    for (int64_t k = 0; k < c->n_slices; k++) {
        int64_t w = (slice_ptr[k + 1] - slice_ptr[k]) / JF_SLICE;
        for (int64_t j = 0; j < w; j++)
            for (int lane = 0; lane < JF_SLICE; lane++) {
                int64_t row = k * JF_SLICE + lane;
                int32_t col = (int32_t)std::min<int64_t>(row, Nf - 1);  // padding: zero block on a valid column
                if (row < Nf && j < rptr[row + 1] - rptr[row]) col = rcol[rptr[row] + j];
                sell_col[slice_ptr[k] + j * JF_SLICE + lane] = col;
            }
    }
    tm.mark("SELL pattern");
    // gather lists for nodal forces: node <- (t*4+m), ascending
    std::vector<int32_t> fptr, fsrc;
    build_incidence(N, T, tets.data(), fptr, fsrc);
    tm.mark("force gather lists");
    // gather lists for the stiffness on the device (jf_contrib_kernel): count, prefix sum, fill
    if ((int64_t)T * 20 > 2000000000LL) {
        jf_set_error("mesh too large for 32-bit element offsets");
        return JF_ERR_ARG;
    }
    std::vector<int32_t> rptr32(rptr.begin(), rptr.end());
    int32_t *d_rptr = nullptr, *d_rcol = nullptr, *d_cur = nullptr;
    if ((rc = upload_async(c, &c->d_slice_ptr, slice_ptr))) return rc;
    if ((rc = upload_async(c, &c->d_sell_col, sell_col))) return rc;
    if ((rc = upload_async(c, &c->d_fn_ptr, fptr))) return rc;
    if ((rc = upload_async(c, &c->d_fn_src, fsrc))) return rc;
    if ((rc = upload_async(c, &d_rptr, rptr32))) return rc;
    if ((rc = upload_async(c, &d_rcol, rcol))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_contrib_ptr, (size_t)c->n_slots + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &d_cur, (size_t)c->n_slots + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_contrib_src, (size_t)16 * T))) return rc;  // upper bound: every (t, m, r) between free nodes
    JF_CUDA(cudaMemsetAsync(c->d_contrib_ptr, 0, ((size_t)c->n_slots + 1) * sizeof(int32_t), c->stream));
    if (Nf > 0) {
        const unsigned grid_rows = (unsigned)((Nf + 127) / 128);
        jf_contrib_kernel<false><<<grid_rows, 128, 0, c->stream>>>(Nf, c->d_tets, c->d_fn_ptr, c->d_fn_src, d_rptr, d_rcol, c->d_slice_ptr,
                                                                   c->d_contrib_ptr, nullptr);
        jf_prefix_kernel<<<1, 1024, 0, c->stream>>>((long long)c->n_slots + 1, c->d_contrib_ptr);
        JF_CUDA(cudaMemcpyAsync(d_cur, c->d_contrib_ptr, ((size_t)c->n_slots + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        jf_contrib_kernel<true><<<grid_rows, 128, 0, c->stream>>>(Nf, c->d_tets, c->d_fn_ptr, c->d_fn_src, d_rptr, d_rcol, c->d_slice_ptr, d_cur,
                                                                  c->d_contrib_src);
        JF_CUDA(cudaGetLastError());
    }
    int bad_tet = -1;
    JF_CUDA(cudaMemcpyAsync(&bad_tet, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    for (void *tmp_buf : {(void *)d_rptr, (void *)d_rcol, (void *)d_cur, (void *)d_bad, (void *)d_torder}) cudaFreeAsync(tmp_buf, c->stream);
    if (bad_tet >= 0) {
        jf_set_error("tetrahedron %d is degenerate (zero volume)", bad_tet);
        return JF_ERR_ARG;
    }
    tm.mark("stiffness lists (device)");
    c->h_rowlen.resize(Nf);
    for (int64_t i = 0; i < Nf; i++) c->h_rowlen[i] = (int32_t)(rptr[i + 1] - rptr[i]);
    c->h_sell_col.swap(sell_col);
    c->h_slice_ptr.swap(slice_ptr);
    return 0;
}
