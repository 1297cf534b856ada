// Resident CG kernel: see jf_cg.cu for the overview of the CG design.
//
// One CTA owns `W` consecutive 32-row slices of ONE system for the whole solve:
//   * their 3x3 blocks (W * wmax * 288 doubles) are staged once into shared memory,
//   * x, r, p, Ap of a row live in the registers of the thread that owns the row,
//   * each iteration the CTA first stages its HALO -- every distinct block column its slices
//     touch, ~4x fewer than the references to them -- computing p_j = r_j + beta*pold_j once per
//     node with coalesced-as-possible 256-bit loads, into shared memory; the block-row products
//     then read p_j from shared memory through a 16-bit local column index.
// The global gathers were what bounded the previous version (one L1tex wavefront per distinct line
// per gather instruction: DESIGN.md K3); the halo cuts their number by the reuse factor.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "jf_cg_common.cuh"

namespace {
constexpr int HALO_BATCH = 4;

__global__ void __launch_bounds__(RES_WARPS * 32, 1) jf_cg_resident_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ CgShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int s = blockIdx.x / a.ctas_per_sys;
    const int cb = blockIdx.x - s * a.ctas_per_sys;
    const long long sl = (long long)cb * W + warp;
    const bool have = sl < a.n_slices;
    const long long row = sl * 32 + lane;
    const long long vstride = a.Nf_pad * 4;
    // shared-memory carve-up
    double *mv_all = (double *)dyn_smem;                                   // W * wmax * 288
    double *hv = mv_all + (size_t)W * a.wmax * 288;                        // hmax * 3
    int *hn = (int *)(hv + (size_t)a.hmax * 3);                            // hmax
    unsigned short *lc_all = (unsigned short *)(hn + a.hmax);              // W * wmax * 32
    double *mv = mv_all + (size_t)warp * a.wmax * 288;
    unsigned short *lc = lc_all + (size_t)warp * a.wmax * 32;
    const int h0 = a.halo_ptr[cb], nh = a.halo_ptr[cb + 1] - h0;
    unsigned epoch = 0, nsum = 0;

    cg_init_flags(a, sh);
    const bool sys_on = !sh.done[s];

    // stage once: matrix slice in SELL order (lane-contiguous), local column indices, halo node ids
    int w = 0;
    if (have && sys_on) {
        const int q0 = a.slice_ptr[sl];
        w = (a.slice_ptr[sl + 1] - q0) >> 5;
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288;
        for (int i = lane; i < w * 288; i += 32) mv[i] = v[i];
        for (int j = 0; j < w; j++) lc[j * 32 + lane] = a.lcol[q0 + j * 32 + lane];
    }
    for (int h = threadIdx.x; h < nh; h += blockDim.x) hn[h] = a.halo_nodes[h0 + h];
    __syncthreads();

    double4 *const R4 = (double4 *)(a.r + s * vstride);
    double4 *const Pa = (double4 *)(a.p0 + s * vstride);
    double4 *const Pb = (double4 *)(a.p1 + s * vstride);
    double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, y0 = 0, y1 = 0, y2 = 0;
    double acc = 0.0;
    if (have && sys_on) {
        const double4 bv = ((const double4 *)(a.b + s * vstride))[row];
        r0 = bv.x; r1 = bv.y; r2 = bv.z;
        R4[row] = bv;
        Pb[row] = make_double4(0.0, 0.0, 0.0, 0.0);
        acc = r0 * r0 + r1 * r1 + r2 * r2;
    }
    acc = warp_sum(acc);
    if (lane == 0) sh.warp[s * MAXW + warp] = acc;
    grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    cg_after_normb(a, sh);

    int cur = 0;
#ifdef JF_CG_TIMING
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.alldone) break;
        TMARK(0);
        const bool on = !sh.done[s];  // uniform over the CTA: all its slices belong to system s
        acc = 0.0;
        if (on) {
            const double beta = sh.beta[s];
            const double4 *Po = cur ? Pa : Pb;
            double4 *Pn = cur ? Pb : Pa;
            // ---- halo: p_j = r_j + beta * pold_j, once per distinct block column of this CTA ----------
            for (int hb = threadIdx.x; hb < nh; hb += HALO_BATCH * blockDim.x) {
                int node[HALO_BATCH];
                double4 rv[HALO_BATCH], pv[HALO_BATCH];
#pragma unroll
                for (int k = 0; k < HALO_BATCH; k++) {
                    const int h = hb + k * blockDim.x;
                    node[k] = hn[h < nh ? h : hb];
                }
#pragma unroll
                for (int k = 0; k < HALO_BATCH; k++) {
                    rv[k] = ld_node(R4 + node[k]);
                    pv[k] = ld_node(Po + node[k]);
                }
#pragma unroll
                for (int k = 0; k < HALO_BATCH; k++) {
                    const int h = hb + k * blockDim.x;
                    if (h < nh) {
                        hv[3 * h] = fma(beta, pv[k].x, rv[k].x);
                        hv[3 * h + 1] = fma(beta, pv[k].y, rv[k].y);
                        hv[3 * h + 2] = fma(beta, pv[k].z, rv[k].z);
                    }
                }
            }
            __syncthreads();
            // ---- block row: Ap_i = sum_j K_ij p_j, matrix and p_j from shared memory -------------------
            if (have) {
                y0 = y1 = y2 = 0.0;
#pragma unroll 3
                for (int j = 0; j < w; j++) {
                    const double *pj = hv + 3 * (int)lc[j * 32 + lane];
                    const double px = pj[0], py = pj[1], pz = pj[2];
                    const double *vj = mv + j * 288 + lane;
                    JF_BLOCK_FMA(vj, 32, px, py, pz)
                }
                p0 = fma(beta, p0, r0);
                p1 = fma(beta, p1, r1);
                p2 = fma(beta, p2, r2);
                Pn[row] = make_double4(p0, p1, p2, 0.0);
                acc = p0 * y0 + p1 * y1 + p2 * y2;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        TMARK(1);
        if (a.sums_only) grid_allreduce_sums_only(a, sh.warp, sh.tot, nsum);
        else grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(2);

        acc = 0.0;
        if (on && have) {
            const double alpha = sh.resid[s] / sh.tot[s];
            x0 = fma(alpha, p0, x0);
            x1 = fma(alpha, p1, x1);
            x2 = fma(alpha, p2, x2);
            r0 = fma(-alpha, y0, r0);
            r1 = fma(-alpha, y1, r1);
            r2 = fma(-alpha, y2, r2);
            R4[row] = make_double4(r0, r1, r2, 0.0);
            acc = r0 * r0 + r1 * r1 + r2 * r2;
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        TMARK(3);
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(4);
        cg_after_rr(a, sh, it);
        cur ^= 1;
        TMARK(5);
    }
#ifdef JF_CG_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("cg-res iters %d cycles/iter: top %.0f phaseA %.0f arA %.0f phaseB %.0f arB %.0f tail %.0f  (nh %d)\n", sh.iters[0],
               (double)tacc[0] / sh.iters[0], (double)tacc[1] / sh.iters[0], (double)tacc[2] / sh.iters[0],
               (double)tacc[3] / sh.iters[0], (double)tacc[4] / sh.iters[0], (double)tacc[5] / sh.iters[0], nh);
#endif

    // (A12) U[free] += step * x ; du = step^2 * sum x^2 ; x kept for jf_get_cg_solution
    acc = 0.0;
    if (have && a.state[s].active) {
        ((double4 *)(a.x + s * vstride))[row] = make_double4(x0, x1, x2, 0.0);
        if (a.apply_update && row < a.Nf) {
            double *U = a.U + ((long long)s * a.N + row) * 3;
            U[0] += a.step * x0;
            U[1] += a.step * x1;
            U[2] += a.step * x2;
            acc = x0 * x0 + x1 * x1 + x2 * x2;
        }
    }
    if (a.apply_update) {
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    }
    cg_write_state(a, sh);
}

// -------------------------------------------------------------------------------------------------
// Register-resident variant (rows up to 18 blocks wide, <= 4 slices per CTA): two lanes share a row,
// lane half h keeps the blocks j = h, h+2, ... (nine 3x3 blocks = 81 doubles) and their 16-bit halo
// indices in REGISTERS for the whole solve, so an iteration touches shared memory only for the
// halo values.  The shared-memory version above moved 213 KB per iteration per SM through the
// 128 B/clk shared-memory pipe (matrix re-read every iteration); this one moves ~30 KB.
// -------------------------------------------------------------------------------------------------
constexpr int RR_J = 9;
constexpr int RR_HALO_BATCH = 2;  // 256 threads stage ~330 halo entries: two per thread suffice

__global__ void __launch_bounds__(256, 1) jf_cg_regres_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ CgShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = blockDim.x >> 6;  // slices per CTA (two warps per 32-row slice)
    const int s = blockIdx.x / a.ctas_per_sys;
    const int cb = blockIdx.x - s * a.ctas_per_sys;
    const int half = lane >> 4;
    const int rloc = warp * 16 + (lane & 15);  // row within the CTA
    const long long sl = (long long)cb * W + (rloc >> 5);
    const bool have = sl < a.n_slices;
    const bool owner = have && half == 0;  // the lane that writes the row's vectors and counts its dots
    const long long row = sl * 32 + (rloc & 31);
    const long long vstride = a.Nf_pad * 4;
    double *hv = (double *)dyn_smem;               // hmax * 3
    int *hn = (int *)(hv + (size_t)a.hmax * 3);    // hmax
    const int h0 = a.halo_ptr[cb], nh = a.halo_ptr[cb + 1] - h0;
    unsigned epoch = 0, nsum = 0;

    cg_init_flags(a, sh);
    const bool sys_on = !sh.done[s];

    // the row's blocks and local column indices, once, into registers
    double m[RR_J][9];
    int lc[RR_J];
    int w = 0;
#pragma unroll
    for (int jj = 0; jj < RR_J; jj++) {
        lc[jj] = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) m[jj][q] = 0.0;
    }
    if (have && sys_on) {
        const int q0 = a.slice_ptr[sl];
        w = (a.slice_ptr[sl + 1] - q0) >> 5;
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288 + (rloc & 31);
#pragma unroll
        for (int jj = 0; jj < RR_J; jj++) {
            const int j = 2 * jj + half;
            if (j < w) {
                lc[jj] = a.lcol[q0 + j * 32 + (rloc & 31)];
#pragma unroll
                for (int q = 0; q < 9; q++) m[jj][q] = v[288 * j + 32 * q];
            }
        }
    }
    for (int h = threadIdx.x; h < nh; h += blockDim.x) hn[h] = a.halo_nodes[h0 + h];
    __syncthreads();

    double4 *const R4 = (double4 *)(a.r + s * vstride);
    double4 *const Pa = (double4 *)(a.p0 + s * vstride);
    double4 *const Pb = (double4 *)(a.p1 + s * vstride);
    double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, y0 = 0, y1 = 0, y2 = 0;
    double acc = 0.0;
    if (have && sys_on) {
        const double4 bv = ((const double4 *)(a.b + s * vstride))[row];
        r0 = bv.x; r1 = bv.y; r2 = bv.z;
        if (owner) {
            R4[row] = bv;
            Pb[row] = make_double4(0.0, 0.0, 0.0, 0.0);
            acc = r0 * r0 + r1 * r1 + r2 * r2;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) sh.warp[s * MAXW + warp] = acc;
    grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    cg_after_normb(a, sh);

    int cur = 0;
#ifdef JF_CG_TIMING
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.alldone) break;
        TMARK(0);
        const bool on = !sh.done[s];
        acc = 0.0;
        if (on) {
            const double beta = sh.beta[s];
            const double4 *Po = cur ? Pa : Pb;
            double4 *Pn = cur ? Pb : Pa;
            for (int hb = threadIdx.x; hb < nh; hb += RR_HALO_BATCH * blockDim.x) {
                int node[RR_HALO_BATCH];
                double4 rv[RR_HALO_BATCH], pv[RR_HALO_BATCH];
#pragma unroll
                for (int k = 0; k < RR_HALO_BATCH; k++) {
                    const int h = hb + k * blockDim.x;
                    node[k] = hn[h < nh ? h : hb];
                }
#pragma unroll
                for (int k = 0; k < RR_HALO_BATCH; k++) {
                    rv[k] = ld_node(R4 + node[k]);
                    pv[k] = ld_node(Po + node[k]);
                }
#pragma unroll
                for (int k = 0; k < RR_HALO_BATCH; k++) {
                    const int h = hb + k * blockDim.x;
                    if (h < nh) {
                        hv[3 * h] = fma(beta, pv[k].x, rv[k].x);
                        hv[3 * h + 1] = fma(beta, pv[k].y, rv[k].y);
                        hv[3 * h + 2] = fma(beta, pv[k].z, rv[k].z);
                    }
                }
            }
            __syncthreads();
            if (have) {
                y0 = y1 = y2 = 0.0;
#pragma unroll
                for (int jj = 0; jj < RR_J; jj++) {
                    if (2 * jj < w) {  // warp-uniform; the odd leftover multiplies a zero block
                        const double *pj = hv + 3 * lc[jj];
                        const double px = pj[0], py = pj[1], pz = pj[2];
                        JF_BLOCK_FMA(m[jj], 1, px, py, pz)
                    }
                }
                y0 += __shfl_xor_sync(0xffffffffu, y0, 16);
                y1 += __shfl_xor_sync(0xffffffffu, y1, 16);
                y2 += __shfl_xor_sync(0xffffffffu, y2, 16);
                p0 = fma(beta, p0, r0);
                p1 = fma(beta, p1, r1);
                p2 = fma(beta, p2, r2);
                if (owner) {
                    Pn[row] = make_double4(p0, p1, p2, 0.0);
                    acc = p0 * y0 + p1 * y1 + p2 * y2;
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        TMARK(1);
        if (a.sums_only) grid_allreduce_sums_only(a, sh.warp, sh.tot, nsum);
        else grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(2);

        acc = 0.0;
        if (on && have) {
            const double alpha = sh.resid[s] / sh.tot[s];
            x0 = fma(alpha, p0, x0);
            x1 = fma(alpha, p1, x1);
            x2 = fma(alpha, p2, x2);
            r0 = fma(-alpha, y0, r0);
            r1 = fma(-alpha, y1, r1);
            r2 = fma(-alpha, y2, r2);
            if (owner) {
                R4[row] = make_double4(r0, r1, r2, 0.0);
                acc = r0 * r0 + r1 * r1 + r2 * r2;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        TMARK(3);
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(4);
        cg_after_rr(a, sh, it);
        cur ^= 1;
        TMARK(5);
    }
#ifdef JF_CG_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("cg-regres iters %d cycles/iter: top %.0f phaseA %.0f arA %.0f phaseB %.0f arB %.0f tail %.0f  (nh %d)\n", sh.iters[0],
               (double)tacc[0] / sh.iters[0], (double)tacc[1] / sh.iters[0], (double)tacc[2] / sh.iters[0],
               (double)tacc[3] / sh.iters[0], (double)tacc[4] / sh.iters[0], (double)tacc[5] / sh.iters[0], nh);
#endif

    acc = 0.0;
    if (owner && a.state[s].active) {
        ((double4 *)(a.x + s * vstride))[row] = make_double4(x0, x1, x2, 0.0);
        if (a.apply_update && row < a.Nf) {
            double *U = a.U + ((long long)s * a.N + row) * 3;
            U[0] += a.step * x0;
            U[1] += a.step * x1;
            U[2] += a.step * x2;
            acc = x0 * x0 + x1 * x1 + x2 * x2;
        }
    }
    if (a.apply_update) {
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    }
    cg_write_state(a, sh);
}

size_t resident_smem(int W, int wmax, int hmax) {
    return (size_t)W * wmax * 288 * sizeof(double) + (size_t)hmax * (3 * sizeof(double) + sizeof(int)) +
           (size_t)W * wmax * 32 * sizeof(unsigned short);
}
}  // namespace

// Decide whether the context's systems fit the resident kernel; if so build the per-CTA halo lists.
// Returns 1 (resident configured), 0 (does not fit), or a negative jf_status.
int jf_cg_resident_configure(jf_ctx *c) {
    c->cg_resident = 0;
    if (c->flags & JF_FLAG_NO_RESIDENT) return 0;
    if (c->n_slices < 1 || c->n_sys > c->sm_count) return 0;
    cudaFuncAttributes fa;
    JF_CUDA(cudaFuncGetAttributes(&fa, jf_cg_resident_kernel));
    int max_optin = 0;
    JF_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const long long avail = (long long)max_optin - (long long)fa.sharedSizeBytes - 512;
    // fewest warps per CTA such that all systems fit with one CTA per SM
    const int ctas_budget = c->sm_count / c->n_sys;
    if (ctas_budget < 1) return 0;
    int W = (int)((c->n_slices + ctas_budget - 1) / ctas_budget);
    if (W < 1) W = 1;
    if (W > RES_WARPS) return 0;
    const int ctas = (int)((c->n_slices + W - 1) / W);
    // halo lists: sorted distinct block columns (incl. padding columns, which repeat the row) per CTA
    std::vector<int32_t> hptr(ctas + 1, 0), hnodes;
    std::vector<uint16_t> lcol((size_t)c->n_slots, 0);
    int hmax = 1;
    std::vector<std::vector<int32_t>> lists((size_t)ctas);
    jf_parallel_for(ctas, [&](int64_t lo, int64_t hi) {
        for (int64_t cb = lo; cb < hi; cb++) {
            const int64_t s0 = cb * W, s1 = std::min<int64_t>(c->n_slices, s0 + W);
            const int64_t q0 = c->h_slice_ptr[s0], q1 = c->h_slice_ptr[s1];
            // the CTA's own rows first, in row order (entry i = local row i; jf_cg_fused_kernel keeps their p in
            // registers and only mirrors it here), then the foreign block columns, sorted
            std::vector<int32_t> &tmp = lists[cb];
            const int32_t own0 = (int32_t)(s0 * JF_SLICE), own1 = (int32_t)(s1 * JF_SLICE);
            std::vector<int32_t> foreign;
            for (int64_t q = q0; q < q1; q++) {
                const int32_t col = c->h_sell_col[q];
                if (col < own0 || col >= own1) foreign.push_back(col);
            }
            std::sort(foreign.begin(), foreign.end());
            foreign.erase(std::unique(foreign.begin(), foreign.end()), foreign.end());
            tmp.resize((size_t)(own1 - own0));
            for (int32_t i = own0; i < own1; i++) tmp[i - own0] = i;
            tmp.insert(tmp.end(), foreign.begin(), foreign.end());
            if (tmp.size() > 65535) continue;
            const int32_t n_own = own1 - own0;
            for (int64_t q = q0; q < q1; q++) {
                const int32_t col = c->h_sell_col[q];
                lcol[q] = (col >= own0 && col < own1)
                              ? (uint16_t)(col - own0)
                              : (uint16_t)(n_own + (std::lower_bound(foreign.begin(), foreign.end(), col) - foreign.begin()));
            }
        }
    }, 8);
    c->cg_foreign_max = 0;
    for (int cb = 0; cb < ctas; cb++) {
        if (lists[cb].size() > 65535) return 0;
        const int64_t own = (std::min<int64_t>(c->n_slices, (int64_t)(cb + 1) * W) - (int64_t)cb * W) * JF_SLICE;
        c->cg_foreign_max = std::max(c->cg_foreign_max, (int)((int64_t)lists[cb].size() - own));
        hnodes.insert(hnodes.end(), lists[cb].begin(), lists[cb].end());
        hptr[cb + 1] = (int32_t)hnodes.size();
        hmax = std::max(hmax, (int)lists[cb].size());
    }
    size_t smem = resident_smem(W, c->cg_wmax, hmax);
    int per_sm = 0, variant = 1;
    if (c->cg_wmax <= 2 * RR_J && W <= 4 && !getenv("JF_CG_NO_REGRES")) {
        // matrix in registers: shared memory holds only the halo
        const size_t smem_rr = (size_t)hmax * (3 * sizeof(double) + sizeof(int));
        JF_CUDA(cudaFuncSetAttribute(jf_cg_regres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rr));
        JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_regres_kernel, W * 64, smem_rr));
        if (per_sm >= 1) {
            variant = 2;
            smem = smem_rr;
        }
    }
    if (variant == 1) {
        if ((long long)smem > avail) return 0;
        JF_CUDA(cudaFuncSetAttribute(jf_cg_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_resident_kernel, W * 32, smem));
        if (per_sm < 1) return 0;
    }
    int rc;
    if ((rc = jf_dev_alloc(c, &c->d_halo_ptr, hptr.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_halo_nodes, hnodes.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_lcol, lcol.size()))) return rc;
    JF_CUDA(cudaMemcpy(c->d_halo_ptr, hptr.data(), hptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_halo_nodes, hnodes.data(), hnodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_lcol, lcol.data(), lcol.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(hptr.size() * 4 + hnodes.size() * 4 + lcol.size() * 2);
    c->cg_resident = variant;
    c->cg_res_warps = W;
    c->cg_res_smem = smem;
    c->cg_hmax = hmax;
    c->cg_ctas_per_sys = ctas;
    c->cg_grid = ctas * c->n_sys;
    return 1;
}

int jf_cg_resident_launch(jf_ctx *c, CgArgs &a) {
    a.halo_ptr = c->d_halo_ptr;
    a.halo_nodes = c->d_halo_nodes;
    a.lcol = c->d_lcol;
    a.hmax = c->cg_hmax;
    a.ctas_per_sys = c->cg_ctas_per_sys;
    void *args[] = {&a};
    if (c->cg_resident == 2)
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_regres_kernel, dim3(c->cg_grid), dim3(c->cg_res_warps * 64), args,
                                            c->cg_res_smem, c->stream));
    else
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_resident_kernel, dim3(c->cg_grid), dim3(c->cg_res_warps * 32), args,
                                            c->cg_res_smem, c->stream));
    return 0;
}
