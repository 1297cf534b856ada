// K3/K4 -- the whole conjugate-gradient solve of one relaxation step as ONE persistent
// cooperative kernel (saenopy's `cg` is a Python loop around scipy matvecs,
// called from _solve_CG under /root/reference/jointforces/simulation.py:167).
//
// Arithmetic is plain Hestenes-Stiefel in saenopy's order -- alpha = resid/(p.Ap); x += alpha p;
// r -= alpha Ap; rsnew = r.r; stop when rsnew < tol*normb; beta = rsnew/resid; p = r + beta p --
// for all systems of the context in lockstep.  Two grid-wide synchronisations per iteration:
//   phase A  Ap = K p, with p_j = r_j + beta*pold_j formed on the fly while gathering (so the
//            "p = r + beta p" pass and its barrier disappear; own-row p goes to the other p
//            buffer), fused p.Ap partials
//   phase B  x += alpha p, r -= alpha Ap, fused r.r partials
//
// Dot products: lane -> warp shuffle tree -> one slot per CTA -> grid barrier -> every CTA adds
// the slots in the same fixed order.  No atomics on data, so results are bitwise reproducible
// for a given grid.  (A flag-in-data slot protocol that fuses barrier and reduction was measured
// slower than the cooperative-groups barrier on 134-148 CTAs -- tools/micro/barrier_bench.cu,
// DESIGN.md "Barrier" -- because every CTA polls every slot.)
//
// Two kernels share that skeleton:
//   jf_cg_resident_kernel  every (system, 32-row slice) has its own warp for the whole solve: the
//                          slice's matrix blocks and column indices are staged ONCE in shared
//                          memory, x/r/p/Ap of the row live in registers, and the only global
//                          traffic per iteration is the neighbour gather of r and p.  Used when
//                          n_sys * n_slices <= #SM * warps-that-fit (config 2: 536 slices, 18 MB
//                          of blocks spread over the 148 SMs' shared memory).
//   jf_cg_stream_kernel    larger meshes / wider batches: matrix streamed from L2/HBM in SELL-32
//                          order (every warp load is a 256-byte line), vectors in global memory.
// In both, the gathers of one batch of block columns are issued before any of them is consumed
// (explicit register staging), because a row's 15-17 dependent L2 round trips were what bounded
// the first version.
//
// Matrix: SELL-32 3x3 blocks (jf_assembly.cu); one thread owns one block row.  Vectors are
// (row,4) doubles so a gathered node is one 32-byte sector.
#include <cooperative_groups.h>

#include "jf_internal.h"

namespace cg = cooperative_groups;

#ifdef JF_CG_TIMING
#include <cstdio>
#define TMARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } } while (0)
#else
#define TMARK(i) do { } while (0)
#endif

namespace {

constexpr int MAXW = 32;      // warps per CTA upper bound (shared partial table stride)
constexpr int RES_WARPS = 8;  // resident kernel: at most 8 warps per CTA -> 255 registers per thread
constexpr int RES_BATCH = 9;  // block columns gathered per batch (resident)
constexpr int STR_BATCH = 3;  // block columns gathered per batch (streaming)

struct CgArgs {
    const int32_t *slice_ptr;
    const int32_t *sell_col;
    const double *val;
    const double *b;
    double *x, *r, *p0, *p1, *Ap;
    double *partial;  // (2, n_sys, grid) per-CTA partial sums, double-buffered by barrier parity
    jf_sys_state *state;
    double *U;
    long long N, Nf, Nf_pad, n_slices, n_slots;
    int n_sys;
    int wmax;  // widest slice (resident kernel: shared-memory stride)
    double tol;
    long long maxiter;
    double step;
    int apply_update;
};

// one 32-byte vector entry with a single 256-bit load (LDG.E.256 on sm_100a): a gathered node costs
// one L1tex wavefront per distinct line instead of two
__device__ __forceinline__ double4 ld_node(const double4 *p) {
    double4 v;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Barrier + all-reduce.  In: s_warp[s*MAXW + warp] partial of every warp of this CTA for every
// system.  Out: s_tot[s] = sum over the whole grid, identical in every CTA.  All global writes
// made by this CTA before the call are visible to every CTA after it (grid.sync semantics).
__device__ __forceinline__ void grid_allreduce(const CgArgs &a, cg::grid_group &grid, double *s_warp, double *s_tot,
                                               unsigned &epoch) {
    epoch++;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double *part = a.partial + (size_t)(epoch & 1u) * a.n_sys * gridDim.x;
    __syncthreads();
    for (int s = warp; s < a.n_sys; s += nwarps) {
        double v = lane < nwarps ? s_warp[s * MAXW + lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) part[(size_t)s * gridDim.x + blockIdx.x] = v;
    }
    grid.sync();
    for (int s = warp; s < a.n_sys; s += nwarps) {
        const double *p = part + (size_t)s * gridDim.x;
        double t = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) t += __ldcg(p + i);
        t = warp_sum(t);
        if (lane == 0) s_tot[s] = t;
    }
    __syncthreads();
}

struct CgShared {
    double warp[JF_MAX_SYS * MAXW];
    double tot[JF_MAX_SYS];
    double normb[JF_MAX_SYS], resid[JF_MAX_SYS], alpha[JF_MAX_SYS], beta[JF_MAX_SYS];
    int done[JF_MAX_SYS], iters[JF_MAX_SYS];
    int alldone;
};

__device__ __forceinline__ void cg_init_flags(const CgArgs &a, CgShared &sh) {
    if (threadIdx.x < a.n_sys) {
        sh.done[threadIdx.x] = a.state[threadIdx.x].active ? 0 : 1;
        sh.iters[threadIdx.x] = 0;
        sh.beta[threadIdx.x] = 0.0;
    }
    for (int i = threadIdx.x; i < JF_MAX_SYS * MAXW; i += blockDim.x) sh.warp[i] = 0.0;
    __syncthreads();
}

__device__ __forceinline__ void cg_after_normb(const CgArgs &a, CgShared &sh) {
    if (threadIdx.x < a.n_sys) {
        const int s = threadIdx.x;
        sh.normb[s] = sh.tot[s];
        sh.resid[s] = sh.tot[s];
        if (!sh.done[s] && sh.tot[s] == 0.0) sh.done[s] = 1;  // cg returns 0 when normb == 0
    }
    if (threadIdx.x == 0) {
        int all = 1;
        for (int s = 0; s < a.n_sys; s++) all &= (sh.done[s] || sh.tot[s] == 0.0);
        sh.alldone = all;
    }
    __syncthreads();
}

__device__ __forceinline__ void cg_after_pAp(const CgArgs &a, CgShared &sh) {
    if (threadIdx.x < a.n_sys && !sh.done[threadIdx.x]) sh.alpha[threadIdx.x] = sh.resid[threadIdx.x] / sh.tot[threadIdx.x];
    __syncthreads();
}

// convergence test, beta, and the all-done flag for the next iteration (one barrier for all three)
__device__ __forceinline__ void cg_after_rr(const CgArgs &a, CgShared &sh, long long it) {
    if (threadIdx.x == 0) {
        int all = 1;
        for (int s = 0; s < a.n_sys; s++) {
            if (!sh.done[s]) {
                const double rsnew = sh.tot[s];
                sh.iters[s] = (int)it;
                if (rsnew < a.tol * sh.normb[s] || it == a.maxiter) {
                    sh.done[s] = 1;
                } else {
                    sh.beta[s] = rsnew / sh.resid[s];
                    sh.resid[s] = rsnew;
                }
            }
            all &= sh.done[s];
        }
        sh.alldone = all;
    }
    __syncthreads();
}

__device__ __forceinline__ void cg_write_state(const CgArgs &a, CgShared &sh) {
    if (blockIdx.x == 0 && threadIdx.x < a.n_sys && a.state[threadIdx.x].active) {
        const int s = threadIdx.x;
        a.state[s].cg_iters = sh.iters[s];
        a.state[s].normb = sh.normb[s];
        a.state[s].resid = sh.resid[s];
        if (a.apply_update) a.state[s].du = a.step * a.step * sh.tot[s];
    }
}

#define JF_BLOCK_FMA(vj, STRIDE, px, py, pz)  \
    y0 = fma((vj)[0 * (STRIDE)], px, y0);     \
    y0 = fma((vj)[1 * (STRIDE)], py, y0);     \
    y0 = fma((vj)[2 * (STRIDE)], pz, y0);     \
    y1 = fma((vj)[3 * (STRIDE)], px, y1);     \
    y1 = fma((vj)[4 * (STRIDE)], py, y1);     \
    y1 = fma((vj)[5 * (STRIDE)], pz, y1);     \
    y2 = fma((vj)[6 * (STRIDE)], px, y2);     \
    y2 = fma((vj)[7 * (STRIDE)], py, y2);     \
    y2 = fma((vj)[8 * (STRIDE)], pz, y2);

// =================================================================================================
// resident: one warp per (system, slice); matrix slice in shared memory, vectors in registers
// =================================================================================================
__global__ void __launch_bounds__(RES_WARPS * 32, 1) jf_cg_resident_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ CgShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)warp * gridDim.x + blockIdx.x;  // consecutive slices on different SMs
    const bool have = item < (long long)a.n_sys * a.n_slices;
    const int s = have ? (int)(item / a.n_slices) : 0;
    const long long sl = have ? item - (long long)s * a.n_slices : 0;
    const long long row = sl * 32 + lane;
    const long long vstride = a.Nf_pad * 4;
    const size_t per_warp = (size_t)a.wmax * (288 * sizeof(double) + 32 * sizeof(int));
    double *mv = (double *)(dyn_smem + warp * per_warp);
    int *mc = (int *)(mv + (size_t)a.wmax * 288);
    unsigned epoch = 0;

    cg_init_flags(a, sh);
    const bool mine = have && !sh.done[s];

    // stage the slice once: values in SELL order (lane-contiguous), column indices
    int w = 0;
    if (mine) {
        const int q0 = a.slice_ptr[sl];
        w = (a.slice_ptr[sl + 1] - q0) >> 5;
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288;
        for (int i = lane; i < w * 288; i += 32) mv[i] = v[i];
        for (int j = 0; j < w; j++) mc[j * 32 + lane] = a.sell_col[q0 + j * 32 + lane];
    }
    __syncwarp();

    double4 *const R4 = (double4 *)(a.r + s * vstride);
    double4 *const Pa = (double4 *)(a.p0 + s * vstride);
    double4 *const Pb = (double4 *)(a.p1 + s * vstride);
    double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, y0 = 0, y1 = 0, y2 = 0;
    double acc = 0.0;
    if (mine) {
        const double4 bv = ((const double4 *)(a.b + s * vstride))[row];
        r0 = bv.x; r1 = bv.y; r2 = bv.z;
        R4[row] = bv;
        Pb[row] = make_double4(0.0, 0.0, 0.0, 0.0);
        acc = r0 * r0 + r1 * r1 + r2 * r2;
    }
    acc = warp_sum(acc);
    if (lane == 0 && have) sh.warp[s * MAXW + warp] = acc;
    grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    cg_after_normb(a, sh);

    int cur = 0;
#ifdef JF_CG_TIMING
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.alldone) break;
        TMARK(0);
        const bool act = have && !sh.done[s];
        acc = 0.0;
        if (act) {
            const double beta = sh.beta[s];
            const double4 *Po = cur ? Pa : Pb;
            double4 *Pn = cur ? Pb : Pa;
            y0 = y1 = y2 = 0.0;
            for (int j0 = 0; j0 < w; j0 += RES_BATCH) {
                double4 rj[RES_BATCH], pj[RES_BATCH];
#pragma unroll
                int c[RES_BATCH];
#pragma unroll
                for (int k = 0; k < RES_BATCH; k++) c[k] = (j0 + k < w) ? mc[(j0 + k) * 32 + lane] : (int)row;
#pragma unroll
                for (int k = 0; k < RES_BATCH; k++) {
                    rj[k] = ld_node(R4 + c[k]);
                    pj[k] = ld_node(Po + c[k]);
                }
#pragma unroll
                for (int k = 0; k < RES_BATCH; k++) {
                    if (j0 + k < w) {
                        const double px = fma(beta, pj[k].x, rj[k].x), py = fma(beta, pj[k].y, rj[k].y), pz = fma(beta, pj[k].z, rj[k].z);
                        const double *vj = mv + (j0 + k) * 288 + lane;
                        JF_BLOCK_FMA(vj, 32, px, py, pz)
                    }
                }
            }
            p0 = fma(beta, p0, r0);
            p1 = fma(beta, p1, r1);
            p2 = fma(beta, p2, r2);
            Pn[row] = make_double4(p0, p1, p2, 0.0);
            acc = p0 * y0 + p1 * y1 + p2 * y2;
        }
        acc = warp_sum(acc);
        if (lane == 0 && have) sh.warp[s * MAXW + warp] = acc;
        TMARK(1);
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(2);
        cg_after_pAp(a, sh);

        acc = 0.0;
        if (act) {
            const double alpha = sh.alpha[s];
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
        if (lane == 0 && have) sh.warp[s * MAXW + warp] = acc;
        TMARK(3);
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
        TMARK(4);
        cg_after_rr(a, sh, it);
        cur ^= 1;
        TMARK(5);
    }
#ifdef JF_CG_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("cg-res iters %d cycles/iter: top %.0f phaseA %.0f arA %.0f phaseB %.0f arB %.0f tail %.0f\n", sh.iters[0],
               (double)tacc[0] / sh.iters[0], (double)tacc[1] / sh.iters[0], (double)tacc[2] / sh.iters[0],
               (double)tacc[3] / sh.iters[0], (double)tacc[4] / sh.iters[0], (double)tacc[5] / sh.iters[0]);
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
        if (lane == 0 && have) sh.warp[s * MAXW + warp] = acc;
        grid_allreduce(a, grid, sh.warp, sh.tot, epoch);
    }
    cg_write_state(a, sh);
}

// =================================================================================================
// streaming: matrix from L2/HBM, vectors in global memory, warps loop over (system, slice) items
// =================================================================================================
__global__ void __launch_bounds__(JF_CG_BLOCK, 1) jf_cg_stream_kernel(CgArgs a) {
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
                    for (int j0 = 0; j0 < w; j0 += STR_BATCH) {
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
}

}  // namespace

int jf_cg_configure(jf_ctx *c) {
    // widest slice
    int wmax = 1;
    for (int64_t k = 0; k < c->n_slices; k++) wmax = std::max(wmax, (c->h_slice_ptr[k + 1] - c->h_slice_ptr[k]) / JF_SLICE);
    c->cg_wmax = wmax;
    cudaFuncAttributes fa;
    JF_CUDA(cudaFuncGetAttributes(&fa, jf_cg_resident_kernel));
    int max_optin = 0;
    JF_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const size_t per_warp = (size_t)wmax * (288 * sizeof(double) + 32 * sizeof(int));
    long long avail = (long long)max_optin - (long long)fa.sharedSizeBytes - 1024;
    int warps = (int)std::min<long long>(RES_WARPS, avail > 0 ? avail / (long long)per_warp : 0);
    const long long items = (long long)c->n_sys * c->n_slices;
    c->cg_resident = 0;
    if (warps >= 1 && items <= (long long)warps * c->sm_count && !(c->flags & JF_FLAG_NO_RESIDENT)) {
        // fewest warps per CTA that still cover all items with one CTA per SM
        int need = (int)((items + c->sm_count - 1) / c->sm_count);
        warps = std::max(1, std::min(warps, need));
        size_t smem = per_warp * warps;
        JF_CUDA(cudaFuncSetAttribute(jf_cg_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_resident_kernel, warps * 32, smem));
        if (per_sm >= 1) {
            c->cg_resident = 1;
            c->cg_res_warps = warps;
            c->cg_res_smem = smem;
            c->cg_grid = (int)std::min<long long>(c->sm_count, (items + warps - 1) / warps);
            if (c->cg_grid < 1) c->cg_grid = 1;
            return 0;
        }
    }
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_stream_kernel, JF_CG_BLOCK, 0));
    if (per_sm < 1) {
        jf_set_error("CG kernel does not fit on an SM");
        return JF_ERR_CUDA;
    }
    // enough CTAs to give every slice of every system its own warp, at most what can be co-resident
    const int wpb = JF_CG_BLOCK / 32;
    long long want = (items + wpb - 1) / wpb;
    long long cap = (long long)c->sm_count * per_sm;
    long long grid = want < c->sm_count ? c->sm_count : (want > cap ? cap : want);
    grid = (grid + c->sm_count - 1) / c->sm_count * c->sm_count;
    if (grid > cap) grid = cap;
    c->cg_grid = (int)grid;
    return 0;
}

int jf_launch_cg(jf_ctx *c, double tol, int64_t maxiter, double step, int apply_update) {
    CgArgs a;
    a.slice_ptr = c->d_slice_ptr;
    a.sell_col = c->d_sell_col;
    a.val = c->d_val;
    a.b = c->d_b;
    a.x = c->d_x;
    a.r = c->d_r;
    a.p0 = c->d_p0;
    a.p1 = c->d_p1;
    a.Ap = c->d_Ap;
    a.partial = c->d_partial;
    a.state = c->d_state;
    a.U = c->d_U;
    a.N = c->N;
    a.Nf = c->Nf;
    a.Nf_pad = c->Nf_pad;
    a.n_slices = c->n_slices;
    a.n_slots = c->n_slots;
    a.n_sys = c->n_sys;
    a.wmax = c->cg_wmax;
    a.tol = tol;
    a.maxiter = maxiter;
    a.step = step;
    a.apply_update = apply_update;
    void *args[] = {&a};
    if (c->cg_resident)
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_resident_kernel, dim3(c->cg_grid), dim3(c->cg_res_warps * 32), args,
                                            c->cg_res_smem, c->stream));
    else
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_stream_kernel, dim3(c->cg_grid), dim3(JF_CG_BLOCK), args, 0, c->stream));
    c->stats.kernel_launches++;
    c->stats.cg_launches++;
    return 0;
}
