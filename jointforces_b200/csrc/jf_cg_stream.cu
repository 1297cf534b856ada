// Streaming CG kernel: see jf_cg.cu for the overview.  Matrix streamed from L2/HBM in SELL-32 order
// (every warp load is a 256-byte line), vectors in global memory, warps loop over (system, slice).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "jf_cg_common.cuh"

namespace {

// =================================================================================================
// streaming: matrix from L2/HBM, vectors in global memory, warps loop over (system, slice) items
// =================================================================================================
// CACHE (one system): the kernel is persistent and warp `wg` always works on slices wg, wg + nwg, ...; the first ST_TM_BLOCKS
// block columns of ONE of them live in the thread's share of TENSOR MEMORY (one CTA per SM: all 512 columns, the four warps of
// a lane quarter take 128 each = 7 blocks per thread) and the next `fs_sm_blocks` in shared memory for the whole CG solve:
// on an HBM-bound matrix that is 148 SMs x (229 KB + 110 KB) = 50 MB that no longer come from DRAM in every iteration
// (BASELINE configs[4]: 8 % of 643 MB).  Which slice of a warp is cached rotates with the warp, so that at any moment most
// warps stream.  Same summation order, same bits.
constexpr int ST_TM_COLS = 512;
constexpr int ST_TM_BLOCKS = (ST_TM_COLS / 4) / 18;

template <bool CACHE>
__global__ void __launch_bounds__(JF_CG_BLOCK, 1) jf_cg_stream_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double st_cache[];  // CACHE: fs_sm_blocks x 9 x JF_CG_BLOCK doubles, [column][component][thread]
    __shared__ CgShared sh;
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long wg = (long long)warp * gridDim.x + blockIdx.x;
    const long long nwg = (long long)gridDim.x * nwarps;
    const long long vstride = a.Nf_pad * 4;
    unsigned epoch = 0;

    if (CACHE) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((unsigned long long)__cvta_generic_to_shared(&s_tmem)),
                         "n"(ST_TM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    cg_init_flags(a, sh);
    int tmc = 0, smc = 0;  // block columns of the cached slice in tensor memory / in shared memory
    long long csl = -1;    // the cached slice of this warp
    unsigned tm = 0;
    const double *msm = st_cache + threadIdx.x;
    if (CACHE) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm = s_tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(warp >> 2) * (ST_TM_COLS / 4);
        // which of the warp's slices is cached rotates with the warp: at any moment only 1/K of the warps work on a cached
        // slice, the others keep DRAM busy
        const long long K = wg < a.n_slices ? (a.n_slices - wg + nwg - 1) / nwg : 0;
        csl = K > 0 ? wg + ((long long)(warp + blockIdx.x) % K) * nwg : -1;
        if (csl >= 0 && !sh.done[0]) {
            const int q0 = a.slice_ptr[csl];
            const int w0 = (a.slice_ptr[csl + 1] - q0) >> 5;
            tmc = min(w0, ST_TM_BLOCKS);
            smc = min(w0 - tmc, a.fs_sm_blocks);
            const double *v = a.val + (long long)(q0 >> 5) * 288 + lane;
            for (int t = 0; t < tmc; t++) {
#pragma unroll
                for (int q = 0; q < 9; q++) tmem_st_double(tm + t * 18 + 2 * q, v[288 * t + 32 * q]);
            }
            for (int t = 0; t < smc; t++) {
#pragma unroll
                for (int q = 0; q < 9; q++) st_cache[(t * 9 + q) * JF_CG_BLOCK + threadIdx.x] = v[288 * (tmc + t) + 32 * q];
            }
        }
        tmem_wait_st();
    }

    // ---- init: x = 0, r = b, pold = 0, normb = b.b --------------------------------------------
    for (int s = 0; s < a.n_sys; s++) {
        double acc = 0.0;
        if (!sh.done[s]) {
            const double4 *b4 = (const double4 *)(a.b + s * vstride);
            double4 *x4 = (double4 *)(a.x + s * vstride), *r4 = (double4 *)(a.r + s * vstride), *q4 = (double4 *)(a.p1 + s * vstride);
            for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                const long long row = sl * 32 + lane;
                const double4 bv = b4[row];
                x4[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                q4[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                r4[row] = bv;
                acc += bv.x * bv.x + bv.y * bv.y + bv.z * bv.z;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
    }
    grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    cg_after_normb(a, sh);

    int cur = 0;  // phase A writes p into buffer `cur`, reads pold from the other one
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.alldone) break;
        double *pnew_base = cur ? a.p1 : a.p0;
        const double *pold_base = cur ? a.p0 : a.p1;

        // ---- phase A: Ap = K (r + beta pold), p.Ap ------------------------------------------------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!sh.done[s]) {
                const double beta = sh.beta[s];
                const double4 *r4 = (const double4 *)(a.r + s * vstride);
                const double4 *po4 = (const double4 *)(pold_base + s * vstride);
                double4 *pn4 = (double4 *)(pnew_base + s * vstride);
                double4 *Ap4 = (double4 *)(a.Ap + s * vstride);
                const double *vals = a.val + (long long)s * a.n_slots * 9;
                for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                    const long long row = sl * 32 + lane;
                    const int q0 = a.slice_ptr[sl];
                    const int w = (a.slice_ptr[sl + 1] - q0) >> 5;
                    const double *v = vals + (long long)(q0 >> 5) * 288 + lane;
                    const int32_t *ci = a.sell_col + q0 + lane;
                    const double4 ri = r4[row], pi = po4[row];
                    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                    int jfirst = 0;
                    if (CACHE && sl == csl) {  // the cached columns of this slice: gathers in batches, blocks one at a time, same order
                        for (int t0 = 0; t0 < tmc + smc; t0 += STR_BATCH) {
                            double4 rv[STR_BATCH], pv[STR_BATCH];
#pragma unroll
                            for (int k = 0; k < STR_BATCH; k++) {
                                const int cc = (t0 + k < tmc + smc) ? ci[32 * (t0 + k)] : (int)row;
                                rv[k] = ld_node(r4 + cc);
                                pv[k] = ld_node(po4 + cc);
                            }
#pragma unroll
                            for (int k = 0; k < STR_BATCH; k++) {
                                const int t = t0 + k;
                                if (t < tmc + smc) {
                                    double md[9];
                                    if (t < tmc) {
                                        unsigned mw[18];
                                        tmem_ld_x16(tm + t * 18, mw);
                                        tmem_ld_x2(tm + t * 18 + 16, mw + 16);
                                        tmem_wait_ld();
#pragma unroll
                                        for (int q = 0; q < 9; q++) md[q] = __hiloint2double((int)mw[2 * q + 1], (int)mw[2 * q]);
                                    } else {
#pragma unroll
                                        for (int q = 0; q < 9; q++) md[q] = msm[((t - tmc) * 9 + q) * JF_CG_BLOCK];
                                    }
                                    const double px = fma(beta, pv[k].x, rv[k].x), py = fma(beta, pv[k].y, rv[k].y), pz = fma(beta, pv[k].z, rv[k].z);
                                    JF_BLOCK_FMA(md, 1, px, py, pz)
                                }
                            }
                        }
                        jfirst = tmc + smc;
                    }
                    for (int j0 = jfirst; j0 < w; j0 += STR_BATCH) {
                        int c[STR_BATCH];
                        double4 rj[STR_BATCH], pj[STR_BATCH];
                        double m[STR_BATCH][9];
#pragma unroll
                        for (int k = 0; k < STR_BATCH; k++) c[k] = (j0 + k < w) ? ci[32 * (j0 + k)] : (int)row;
#pragma unroll
                        for (int k = 0; k < STR_BATCH; k++) {
                            const int jj = (j0 + k < w) ? j0 + k : j0;  // clamped: never read past the slice
#pragma unroll
                            for (int q = 0; q < 9; q++) m[k][q] = v[288 * jj + 32 * q];
                            rj[k] = ld_node(r4 + c[k]);
                            pj[k] = ld_node(po4 + c[k]);
                        }
#pragma unroll
                        for (int k = 0; k < STR_BATCH; k++) {
                            if (j0 + k < w) {
                                const double px = fma(beta, pj[k].x, rj[k].x), py = fma(beta, pj[k].y, rj[k].y), pz = fma(beta, pj[k].z, rj[k].z);
                                JF_BLOCK_FMA(m[k], 1, px, py, pz)
                            }
                        }
                    }
                    const double px = fma(beta, pi.x, ri.x), py = fma(beta, pi.y, ri.y), pz = fma(beta, pi.z, ri.z);
                    pn4[row] = make_double4(px, py, pz, 0.0);
                    Ap4[row] = make_double4(y0, y1, y2, 0.0);
                    acc += px * y0 + py * y1 + pz * y2;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        cg_after_pAp(a, sh);

        // ---- phase B: x += alpha p, r -= alpha Ap, r.r ----------------------------------------------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!sh.done[s]) {
                const double alpha = sh.alpha[s];
                double4 *x4 = (double4 *)(a.x + s * vstride), *r4 = (double4 *)(a.r + s * vstride);
                const double4 *pn4 = (const double4 *)(pnew_base + s * vstride);
                const double4 *Ap4 = (const double4 *)(a.Ap + s * vstride);
                for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                    const long long row = sl * 32 + lane;
                    const double4 p = pn4[row], q = Ap4[row];
                    double4 xv = x4[row], rv = r4[row];
                    xv.x = fma(alpha, p.x, xv.x);
                    xv.y = fma(alpha, p.y, xv.y);
                    xv.z = fma(alpha, p.z, xv.z);
                    rv.x = fma(-alpha, q.x, rv.x);
                    rv.y = fma(-alpha, q.y, rv.y);
                    rv.z = fma(-alpha, q.z, rv.z);
                    x4[row] = xv;
                    r4[row] = rv;
                    acc += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        cg_after_rr(a, sh, it);
        cur ^= 1;
    }

    // ---- (A12) U[free] += step * x ; du = step^2 * sum x^2 -------------------------------------------
    if (a.apply_update) {
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (a.state[s].active) {
                const double4 *x4 = (const double4 *)(a.x + s * vstride);
                double *U = a.U + (long long)s * a.N * 3;
                for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                    const long long row = sl * 32 + lane;
                    if (row < a.Nf) {
                        const double4 xv = x4[row];
                        U[3 * row] += a.step * xv.x;
                        U[3 * row + 1] += a.step * xv.y;
                        U[3 * row + 2] += a.step * xv.z;
                        acc += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z;
                    }
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    }
    cg_write_state(a, sh);
    if (CACHE) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(ST_TM_COLS) : "memory");
    }
}
constexpr int SH_SLICES = 8;               // slices per halo group = warps per CTA
constexpr int SH_THREADS = SH_SLICES * 32;

// =================================================================================================
// streaming with per-group halo staging: the neighbour gathers of a 256-row group are done once per
// distinct node (coalesced-as-possible 256-bit loads into shared memory) instead of once per
// reference; the matrix stream stays as it was.  Two CTAs per SM overlap one group's staging with
// another group's products.
// =================================================================================================
__global__ void __launch_bounds__(SH_THREADS, 2) jf_cg_streamhalo_kernel(CgArgs a) {
    extern __shared__ double hv[];  // hmax * 3: p_j = r_j + beta*pold_j of the current group's halo
    cg::grid_group grid = cg::this_grid();
    __shared__ CgShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long wg = (long long)warp * gridDim.x + blockIdx.x;
    const long long nwg = (long long)gridDim.x * nwarps;
    const long long vstride = a.Nf_pad * 4;
    unsigned epoch = 0;

    cg_init_flags(a, sh);

    // ---- init: x = 0, r = b, pold = 0, normb = b.b --------------------------------------------
    for (int s = 0; s < a.n_sys; s++) {
        double acc = 0.0;
        if (!sh.done[s]) {
            const double4 *b4 = (const double4 *)(a.b + s * vstride);
            double4 *x4 = (double4 *)(a.x + s * vstride), *r4 = (double4 *)(a.r + s * vstride), *q4 = (double4 *)(a.p1 + s * vstride);
            for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                const long long row = sl * 32 + lane;
                const double4 bv = b4[row];
                x4[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                q4[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                r4[row] = bv;
                acc += bv.x * bv.x + bv.y * bv.y + bv.z * bv.z;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[s * MAXW + warp] = acc;
    }
    grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    cg_after_normb(a, sh);

    int cur = 0;  // phase A writes p into buffer `cur`, reads pold from the other one
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.alldone) break;
        double *pnew_base = cur ? a.p1 : a.p0;
        const double *pold_base = cur ? a.p0 : a.p1;

        // ---- phase A: Ap = K (r + beta pold), p.Ap; groups of SH_SLICES slices, halo staged per group ------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!sh.done[s]) {
                const double beta = sh.beta[s];
                const double4 *r4 = (const double4 *)(a.r + s * vstride);
                const double4 *po4 = (const double4 *)(pold_base + s * vstride);
                double4 *pn4 = (double4 *)(pnew_base + s * vstride);
                double4 *Ap4 = (double4 *)(a.Ap + s * vstride);
                const double *vals = a.val + (long long)s * a.n_slots * 9;
                const long long n_groups = (a.n_slices + SH_SLICES - 1) / SH_SLICES;
                for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
                    const int h0 = a.halo_ptr[g], nh = a.halo_ptr[g + 1] - h0;
                    for (int hb = threadIdx.x; hb < nh; hb += 4 * SH_THREADS) {
                        int node[4];
                        double4 rv[4], pv[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int h = hb + k * SH_THREADS;
                            node[k] = a.halo_nodes[h0 + (h < nh ? h : hb)];
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            rv[k] = ld_node(r4 + node[k]);
                            pv[k] = ld_node(po4 + node[k]);
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int h = hb + k * SH_THREADS;
                            if (h < nh) {
                                hv[3 * h] = fma(beta, pv[k].x, rv[k].x);
                                hv[3 * h + 1] = fma(beta, pv[k].y, rv[k].y);
                                hv[3 * h + 2] = fma(beta, pv[k].z, rv[k].z);
                            }
                        }
                    }
                    __syncthreads();
                    const long long sl = g * SH_SLICES + warp;
                    if (sl < a.n_slices) {
                        const long long row = sl * 32 + lane;
                        const int q0 = a.slice_ptr[sl];
                        const int w = (a.slice_ptr[sl + 1] - q0) >> 5;
                        const double *v = vals + (long long)(q0 >> 5) * 288 + lane;
                        const unsigned short *lc = a.lcol + q0 + lane;
                        const double4 ri = r4[row], pi = po4[row];
                        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                        for (int j0 = 0; j0 < w; j0 += STR_BATCH) {
                            int c[STR_BATCH];
                            double m[STR_BATCH][9];
#pragma unroll
                            for (int k = 0; k < STR_BATCH; k++) {
                                const int jj = (j0 + k < w) ? j0 + k : j0;  // clamped: never read past the slice
                                c[k] = lc[32 * jj];
#pragma unroll
                                for (int q = 0; q < 9; q++) m[k][q] = v[288 * jj + 32 * q];
                            }
#pragma unroll
                            for (int k = 0; k < STR_BATCH; k++) {
                                if (j0 + k < w) {
                                    const double *pj = hv + 3 * c[k];
                                    const double px = pj[0], py = pj[1], pz = pj[2];
                                    JF_BLOCK_FMA(m[k], 1, px, py, pz)
                                }
                            }
                        }
                        const double px = fma(beta, pi.x, ri.x), py = fma(beta, pi.y, ri.y), pz = fma(beta, pi.z, ri.z);
                        pn4[row] = make_double4(px, py, pz, 0.0);
                        Ap4[row] = make_double4(y0, y1, y2, 0.0);
                        acc += px * y0 + py * y1 + pz * y2;
                    }
                    __syncthreads();
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        cg_after_pAp(a, sh);

        // ---- phase B: x += alpha p, r -= alpha Ap, r.r ----------------------------------------------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!sh.done[s]) {
                const double alpha = sh.alpha[s];
                double4 *x4 = (double4 *)(a.x + s * vstride), *r4 = (double4 *)(a.r + s * vstride);
                const double4 *pn4 = (const double4 *)(pnew_base + s * vstride);
                const double4 *Ap4 = (const double4 *)(a.Ap + s * vstride);
                for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                    const long long row = sl * 32 + lane;
                    const double4 p = pn4[row], q = Ap4[row];
                    double4 xv = x4[row], rv = r4[row];
                    xv.x = fma(alpha, p.x, xv.x);
                    xv.y = fma(alpha, p.y, xv.y);
                    xv.z = fma(alpha, p.z, xv.z);
                    rv.x = fma(-alpha, q.x, rv.x);
                    rv.y = fma(-alpha, q.y, rv.y);
                    rv.z = fma(-alpha, q.z, rv.z);
                    x4[row] = xv;
                    r4[row] = rv;
                    acc += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        cg_after_rr(a, sh, it);
        cur ^= 1;
    }

    // ---- (A12) U[free] += step * x ; du = step^2 * sum x^2 -------------------------------------------
    if (a.apply_update) {
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (a.state[s].active) {
                const double4 *x4 = (const double4 *)(a.x + s * vstride);
                double *U = a.U + (long long)s * a.N * 3;
                for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                    const long long row = sl * 32 + lane;
                    if (row < a.Nf) {
                        const double4 xv = x4[row];
                        U[3 * row] += a.step * xv.x;
                        U[3 * row + 1] += a.step * xv.y;
                        U[3 * row + 2] += a.step * xv.z;
                        acc += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z;
                    }
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sh.warp[s * MAXW + warp] = acc;
        }
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    }
    cg_write_state(a, sh);
}
}  // namespace

// group halo lists for the staged variant: sorted distinct block columns of each group of SH_SLICES slices
static int build_group_halos(jf_ctx *c) {
    const int64_t n_groups = (c->n_slices + SH_SLICES - 1) / SH_SLICES;
    std::vector<int32_t> hptr((size_t)n_groups + 1, 0), hnodes;
    std::vector<uint16_t> lcol((size_t)c->n_slots, 0);
    int hmax = 1;
    std::vector<std::vector<int32_t>> lists((size_t)n_groups);
    jf_parallel_for(n_groups, [&](int64_t lo, int64_t hi) {
        for (int64_t g = lo; g < hi; g++) {
            const int64_t s0 = g * SH_SLICES, s1 = std::min<int64_t>(c->n_slices, s0 + SH_SLICES);
            const int64_t q0 = c->h_slice_ptr[s0], q1 = c->h_slice_ptr[s1];
            std::vector<int32_t> &tmp = lists[g];
            tmp.assign(c->h_sell_col.begin() + q0, c->h_sell_col.begin() + q1);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            if (tmp.size() > 65535) continue;
            for (int64_t q = q0; q < q1; q++)
                lcol[q] = (uint16_t)(std::lower_bound(tmp.begin(), tmp.end(), c->h_sell_col[q]) - tmp.begin());
        }
    }, 8);
    for (int64_t g = 0; g < n_groups; g++) {
        if (lists[g].size() > 65535) return 0;
        hnodes.insert(hnodes.end(), lists[g].begin(), lists[g].end());
        if (hnodes.size() > 2000000000ull) return 0;
        hptr[g + 1] = (int32_t)hnodes.size();
        hmax = std::max(hmax, (int)lists[g].size());
    }
    int rc;
    if ((rc = jf_dev_alloc(c, &c->d_halo_ptr, hptr.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_halo_nodes, hnodes.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_lcol, lcol.size()))) return rc;
    JF_CUDA(cudaMemcpy(c->d_halo_ptr, hptr.data(), hptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_halo_nodes, hnodes.data(), hnodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_lcol, lcol.data(), lcol.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(hptr.size() * 4 + hnodes.size() * 4 + lcol.size() * 2);
    c->cg_hmax = hmax;
    return 1;
}

int jf_cg_stream_configure(jf_ctx *c) {
    c->cg_stream_halo = 0;
    // halo staging pays where the matrix is (mostly) L2 resident and the gathers are what is left (config 4,
    // lockstep batches: 3-11 %); on HBM-bound matrices the plain stream is 2.5 % faster (121.4 vs 124.5 us on 503 k nodes)
    const double matrix_mb = (double)c->n_sys * (double)c->n_slots * 72.0 / 1e6;
    if (!getenv("JF_CG_NO_STREAM_HALO") && c->n_slices > 0 && (matrix_mb < 256.0 || getenv("JF_CG_FORCE_STREAM_HALO"))) {
        int ok = build_group_halos(c);
        if (ok < 0) return ok;
        if (ok == 1) {
            const size_t smem = (size_t)c->cg_hmax * 3 * sizeof(double);
            int per_sm = 0;
            JF_CUDA(cudaFuncSetAttribute(jf_cg_streamhalo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_streamhalo_kernel, SH_THREADS, smem));
            if (per_sm >= 1) {
                const long long n_groups = (c->n_slices + SH_SLICES - 1) / SH_SLICES;
                long long grid = std::min<long long>((long long)c->sm_count * per_sm, std::max<long long>(n_groups, c->sm_count));
                c->cg_grid = (int)grid;
                c->cg_res_smem = smem;
                c->cg_stream_halo = 1;
                return 0;
            }
        }
    }
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_stream_kernel<false>, JF_CG_BLOCK, 0));
    if (per_sm < 1) {
        jf_set_error("CG kernel does not fit on an SM");
        return JF_ERR_CUDA;
    }
    // one system on one CTA per SM: cache the first columns of every warp's first slice in tensor + shared memory
    // (one CTA per SM is also what the occupancy calculator grants a kernel that allocates tensor memory, so the
    //  cooperative launch stays)
    c->cg_stream_cache = 0;
    c->cg_stream_smem = 0;
    // Measured on the HBM-bound 503 k-node mesh (profiles/README.md r3n, r3r, r3s), us per iteration: no cache 122.0;
    // first version (every warp's FIRST slice, one column at a time, 5 columns in shared memory) 168.8 -- nothing streams
    // while all warps sit on their cached slices, and 184 KB of shared memory leave the vector gathers 50 KB of L1;
    // rotated cached slice + batched gathers, tensor memory only 119.2; plus 1 / 2 / 3 / 4 / 5 columns in shared
    // memory 118.7 / 118.0 / 117.6 / 121.0 / 169.2.  Default: tensor memory + 3 columns.  JF_CG_STREAM_NO_CACHE=1: off.
    if (c->n_sys == 1 && per_sm == 1 && !getenv("JF_CG_STREAM_NO_CACHE")) {
        cudaFuncAttributes fa;
        int max_optin = 0;
        if (cudaFuncGetAttributes(&fa, jf_cg_stream_kernel<true>) == cudaSuccess &&
            cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) == cudaSuccess) {
            const size_t per_block = (size_t)9 * JF_CG_BLOCK * sizeof(double);
            const long long avail = (long long)max_optin - (long long)fa.sharedSizeBytes - 1024;
            const char *eb = getenv("JF_CG_STREAM_CACHE_SMEM");   // block columns in shared memory on top of tensor memory (default
            const int want = eb ? atoi(eb) : 3;                   //  three: beyond that the vector gathers lose their L1)
            const int blocks = avail > 0 ? (int)std::min<long long>(std::max(0, std::min(want, 8)), avail / (long long)per_block) : 0;
            const size_t smem = blocks * per_block;
            int p2 = 0;
            if (cudaFuncSetAttribute(jf_cg_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, jf_cg_stream_kernel<true>, JF_CG_BLOCK, smem) == cudaSuccess && p2 >= 1) {
                c->cg_stream_cache = 1;
                c->cg_stream_smem = smem;
                c->cg_fs_sm_blocks = blocks;
            }
        }
        cudaGetLastError();
    }
    // enough CTAs to give every slice of every system its own warp, at most what can be co-resident
    const long long items = (long long)c->n_sys * c->n_slices;
    const int wpb = JF_CG_BLOCK / 32;
    long long want = (items + wpb - 1) / wpb;
    long long cap = (long long)c->sm_count * per_sm;
    long long grid = want < c->sm_count ? c->sm_count : (want > cap ? cap : want);
    grid = (grid + c->sm_count - 1) / c->sm_count * c->sm_count;
    if (grid > cap) grid = cap;
    c->cg_grid = (int)grid;
    return 0;
}

int jf_cg_stream_launch(jf_ctx *c, CgArgs &a) {
    void *args[] = {&a};
    if (c->cg_stream_halo) {
        a.halo_ptr = c->d_halo_ptr;
        a.halo_nodes = c->d_halo_nodes;
        a.lcol = c->d_lcol;
        a.hmax = c->cg_hmax;
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_streamhalo_kernel, dim3(c->cg_grid), dim3(SH_THREADS), args, c->cg_res_smem,
                                            c->stream));
        return 0;
    }
    if (c->cg_stream_cache && c->cg_grid <= c->sm_count) {
        a.fs_sm_blocks = (int)(c->cg_stream_smem / ((size_t)9 * JF_CG_BLOCK * sizeof(double)));
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_stream_kernel<true>, dim3(c->cg_grid), dim3(JF_CG_BLOCK), args, c->cg_stream_smem,
                                            c->stream));
        return 0;
    }
    JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_stream_kernel<false>, dim3(c->cg_grid), dim3(JF_CG_BLOCK), args, 0, c->stream));
    return 0;
}
