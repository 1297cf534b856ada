// K3/K4 -- the whole conjugate-gradient solve of one relaxation step as ONE persistent
// cooperative kernel (saenopy's `cg` is a Python loop around scipy matvecs,
// called from _solve_CG under /root/reference/jointforces/simulation.py:167).
//
// Arithmetic is plain Hestenes-Stiefel in saenopy's order -- alpha = resid/(p.Ap); x += alpha p;
// r -= alpha Ap; rsnew = r.r; stop when rsnew < tol*normb; beta = rsnew/resid; p = r + beta p --
// for all systems of the context in lockstep.  Two grid barriers per iteration:
//   phase A  Ap = K p, with p_j = r_j + beta*pold_j formed on the fly while gathering (so the
//            "p = r + beta p" pass and its barrier disappear; own-row p is written to the other
//            p buffer), fused p.Ap partials
//   phase B  x += alpha p, r -= alpha Ap, fused r.r partials
// Dot products are reduced in a fixed order (lane -> warp shuffle tree -> per-CTA slot -> every
// CTA sums the slots in the same order), so results are bitwise reproducible for a given grid.
// The final pass applies U += step*x on the free nodes and reduces du (A12).
//
// Matrix: SELL-32 3x3 blocks (jf_assembly.cu); one thread owns one block row, so every matrix
// load of a warp is a 256-byte line.  Vectors are (row,4) doubles so a gathered node is one
// 32-byte sector.  HBM-bound for matrices beyond L2 (126 MB), barrier/latency-bound below.
#include <cooperative_groups.h>

#include "jf_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int WARPS = JF_CG_BLOCK / 32;

struct CgArgs {
    const int32_t *slice_ptr;
    const int32_t *sell_col;
    const double *val;
    const double *b;
    double *x, *r, *p0, *p1, *Ap;
    double *partial;  // (2, n_sys, grid)
    jf_sys_state *state;
    double *U;
    long long N, Nf, Nf_pad, n_slices, n_slots;
    int n_sys;
    double tol;
    long long maxiter;
    double step;
    int apply_update;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// s_warp[s][warp] -> partial[s][block]; every warp handles systems warp, warp+WARPS, ...
__device__ __forceinline__ void block_publish(double (*s_warp)[WARPS], double *partial, int n_sys) {
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < n_sys; s += WARPS) {
        double v = lane < WARPS ? s_warp[s][lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) partial[(long long)s * gridDim.x + blockIdx.x] = v;
    }
}

// after a grid barrier: s_tot[s] = sum over blocks, same order in every CTA
__device__ __forceinline__ void block_collect(const double *partial, double *s_tot, int n_sys) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < n_sys; s += WARPS) {
        const double *p = partial + (long long)s * gridDim.x;
        double v = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(p + i);
        v = warp_sum(v);
        if (lane == 0) s_tot[s] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(JF_CG_BLOCK) jf_cg_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_warp[JF_MAX_SYS][WARPS];
    __shared__ double s_tot[JF_MAX_SYS];
    __shared__ double s_normb[JF_MAX_SYS], s_resid[JF_MAX_SYS], s_alpha[JF_MAX_SYS], s_beta[JF_MAX_SYS];
    __shared__ int s_done[JF_MAX_SYS], s_iters[JF_MAX_SYS];
    __shared__ int s_alldone;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // consecutive slices go to different CTAs first, so small systems spread over all SMs
    const long long wg = (long long)warp * gridDim.x + blockIdx.x;
    const long long nwg = (long long)gridDim.x * WARPS;
    const long long vstride = a.Nf_pad * 4;
    double *partA = a.partial;
    double *partB = a.partial + (long long)a.n_sys * gridDim.x;

    if (threadIdx.x < a.n_sys) {
        s_done[threadIdx.x] = a.state[threadIdx.x].active ? 0 : 1;
        s_iters[threadIdx.x] = 0;
        s_beta[threadIdx.x] = 0.0;
    }
    __syncthreads();

    // ---- init: x = 0, r = b, pold = 0, normb = b.b --------------------------------------------
    for (int s = 0; s < a.n_sys; s++) {
        double acc = 0.0;
        if (!s_done[s]) {
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
        if (lane == 0) s_warp[s][warp] = acc;
    }
    block_publish(s_warp, partA, a.n_sys);
    grid.sync();
    block_collect(partA, s_tot, a.n_sys);
    if (threadIdx.x < a.n_sys) {
        const int s = threadIdx.x;
        s_normb[s] = s_tot[s];
        s_resid[s] = s_tot[s];
        if (!s_done[s] && s_tot[s] == 0.0) s_done[s] = 1;  // cg returns 0 when normb == 0
    }
    __syncthreads();

    int cur = 0;  // phase A writes p into buffer `cur`, reads pold from the other one
    for (long long it = 1; it <= a.maxiter; it++) {
        if (threadIdx.x == 0) {
            int all = 1;
            for (int s = 0; s < a.n_sys; s++) all &= s_done[s];
            s_alldone = all;
        }
        __syncthreads();
        if (s_alldone) break;
        double *pnew_base = cur ? a.p1 : a.p0;
        const double *pold_base = cur ? a.p0 : a.p1;

        // ---- phase A: Ap = K (r + beta pold), p.Ap ------------------------------------------------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!s_done[s]) {
                const double beta = s_beta[s];
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
                    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll 4
                    for (int j = 0; j < w; j++) {
                        const int c = ci[32 * j];
                        const double4 rj = r4[c], pj = po4[c];
                        const double px = fma(beta, pj.x, rj.x), py = fma(beta, pj.y, rj.y), pz = fma(beta, pj.z, rj.z);
                        const double *vj = v + 288 * j;
                        y0 = fma(vj[0], px, y0);
                        y0 = fma(vj[32], py, y0);
                        y0 = fma(vj[64], pz, y0);
                        y1 = fma(vj[96], px, y1);
                        y1 = fma(vj[128], py, y1);
                        y1 = fma(vj[160], pz, y1);
                        y2 = fma(vj[192], px, y2);
                        y2 = fma(vj[224], py, y2);
                        y2 = fma(vj[256], pz, y2);
                    }
                    const double4 ri = r4[row], pi = po4[row];
                    const double px = fma(beta, pi.x, ri.x), py = fma(beta, pi.y, ri.y), pz = fma(beta, pi.z, ri.z);
                    pn4[row] = make_double4(px, py, pz, 0.0);
                    Ap4[row] = make_double4(y0, y1, y2, 0.0);
                    acc += px * y0 + py * y1 + pz * y2;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) s_warp[s][warp] = acc;
        }
        block_publish(s_warp, partA, a.n_sys);
        grid.sync();
        block_collect(partA, s_tot, a.n_sys);
        if (threadIdx.x < a.n_sys && !s_done[threadIdx.x]) s_alpha[threadIdx.x] = s_resid[threadIdx.x] / s_tot[threadIdx.x];
        __syncthreads();

        // ---- phase B: x += alpha p, r -= alpha Ap, r.r ----------------------------------------------
        for (int s = 0; s < a.n_sys; s++) {
            double acc = 0.0;
            if (!s_done[s]) {
                const double alpha = s_alpha[s];
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
            if (lane == 0) s_warp[s][warp] = acc;
        }
        block_publish(s_warp, partB, a.n_sys);
        grid.sync();
        block_collect(partB, s_tot, a.n_sys);
        if (threadIdx.x < a.n_sys && !s_done[threadIdx.x]) {
            const int s = threadIdx.x;
            const double rsnew = s_tot[s];
            s_iters[s] = (int)it;
            if (rsnew < a.tol * s_normb[s] || it == a.maxiter) {
                s_done[s] = 1;
            } else {
                s_beta[s] = rsnew / s_resid[s];
                s_resid[s] = rsnew;
            }
        }
        __syncthreads();
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
            if (lane == 0) s_warp[s][warp] = acc;
        }
        block_publish(s_warp, partA, a.n_sys);
        grid.sync();
        block_collect(partA, s_tot, a.n_sys);
    }
    if (blockIdx.x == 0 && threadIdx.x < a.n_sys && a.state[threadIdx.x].active) {
        const int s = threadIdx.x;
        a.state[s].cg_iters = s_iters[s];
        a.state[s].normb = s_normb[s];
        a.state[s].resid = s_resid[s];
        if (a.apply_update) a.state[s].du = a.step * a.step * s_tot[s];
    }
}

}  // namespace

int jf_cg_configure(jf_ctx *c) {
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_kernel, JF_CG_BLOCK, 0));
    if (per_sm < 1) {
        jf_set_error("CG kernel does not fit on an SM");
        return JF_ERR_CUDA;
    }
    // enough CTAs to give every slice of every system its own warp, at most what can be co-resident
    long long want = (c->n_slices + WARPS - 1) / WARPS;
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
    a.tol = tol;
    a.maxiter = maxiter;
    a.step = step;
    a.apply_update = apply_update;
    void *args[] = {&a};
    JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_kernel, dim3(c->cg_grid), dim3(JF_CG_BLOCK), args, 0, c->stream));
    c->stats.kernel_launches++;
    c->stats.cg_launches++;
    return 0;
}
