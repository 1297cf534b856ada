// Streaming CG kernel: see jf_cg.cu for the overview.  Matrix streamed from L2/HBM in SELL-32 order
// (every warp load is a 256-byte line), vectors in global memory, warps loop over (system, slice).
#include <algorithm>

#include "jf_cg_common.cuh"

namespace {

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

int jf_cg_stream_configure(jf_ctx *c) {
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_stream_kernel, JF_CG_BLOCK, 0));
    if (per_sm < 1) {
        jf_set_error("CG kernel does not fit on an SM");
        return JF_ERR_CUDA;
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
    JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_stream_kernel, dim3(c->cg_grid), dim3(JF_CG_BLOCK), args, 0, c->stream));
    return 0;
}
