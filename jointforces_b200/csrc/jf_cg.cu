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
//   jf_cg_resident_kernel  (jf_cg_resident.cu) every 32-row slice has its own warp for the whole
//                          solve: the slice's matrix blocks are staged ONCE in shared memory,
//                          x/r/p/Ap of the row live in registers, and per iteration the CTA stages
//                          its halo (each distinct neighbour once) into shared memory.  Used when
//                          all slices fit one CTA per SM (config 2: 536 slices, 18 MB of blocks
//                          spread over the SMs' shared memory).
//   jf_cg_stream_kernel    (jf_cg_stream.cu) larger meshes / wider batches: matrix streamed from L2/HBM in SELL-32
//                          order (every warp load is a 256-byte line), vectors in global memory.
// In both, the gathers of one batch of block columns are issued before any of them is consumed
// (explicit register staging), because a row's 15-17 dependent L2 round trips were what bounded
// the first version.
//
// Matrix: SELL-32 3x3 blocks (jf_assembly.cu); one thread owns one block row.  Vectors are
// (row,4) doubles so a gathered node is one 32-byte sector.
#include <stdlib.h>

#include "jf_cg_common.cuh"

int jf_cg_resident_configure(jf_ctx *c);
int jf_cg_resident_launch(jf_ctx *c, CgArgs &a);
int jf_cg_stream_configure(jf_ctx *c);
int jf_cg_stream_launch(jf_ctx *c, CgArgs &a);
int jf_cg_fused_configure(jf_ctx *c);
int jf_cg_fused_launch(jf_ctx *c, CgArgs &a);
int jf_cg_fused_stream_configure(jf_ctx *c);
int jf_cg_fused_stream_launch(jf_ctx *c, CgArgs &a);
int jf_cg_pair_configure(jf_ctx *c);
int jf_cg_pair_launch(jf_ctx *c, CgArgs &a);

// Streamed matrices that are not much larger than the L2 (BASELINE configs[3]: 92 MB of blocks against 126 MB) are
// read once per CG iteration and partly evicted by the vectors and by their own tail before the next iteration comes
// round (ncu, round 1: 66 % L2 hits, 33 MB per iteration from DRAM).  An access-policy window that marks the matrix as
// persisting in the set-aside part of L2 was tried for that: measured twice on config 4, it is SLOWER than leaving the
// replacement policy alone (19.65 vs 18.28 and 19.96 vs 19.42 us per iteration, profiles/README.md r2u/r2x -- the
// set-aside takes L2 away from the vectors and the gathered halo), so it is opt-in: JF_L2_WINDOW=1.
static int jf_cg_l2_window(jf_ctx *c) {
    if (c->cg_resident || !getenv("JF_L2_WINDOW")) return 0;
    int max_persist = 0, max_window = 0;
    if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device) != cudaSuccess || max_persist <= 0) {
        cudaGetLastError();
        return 0;
    }
    const size_t bytes = (size_t)c->n_sys * c->n_slots * 9 * sizeof(double);
    if (bytes < ((size_t)8 << 20) || bytes > (size_t)4 * max_persist) return 0;  // small: stays anyway; far larger: streams anyway
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.base_ptr = c->d_val;
    v.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)max_persist / (double)v.accessPolicyWindow.num_bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
    return 0;
}

int jf_cg_configure(jf_ctx *c) {
    int wmax = 1;
    for (int64_t k = 0; k < c->n_slices; k++) wmax = std::max(wmax, (c->h_slice_ptr[k + 1] - c->h_slice_ptr[k]) / JF_SLICE);
    c->cg_wmax = wmax;
    int rc = jf_cg_resident_configure(c);
    if (rc < 0) return rc;
    if (rc == 1) {
        rc = jf_cg_fused_configure(c);
        return rc < 0 ? rc : 0;
    }
    rc = jf_cg_stream_configure(c);
    if (rc) return rc;
    rc = jf_cg_fused_stream_configure(c);
    if (rc < 0) return rc;
    rc = jf_cg_pair_configure(c);  // two systems of a config-2-size mesh: side by side, matrices on chip
    if (rc < 0) return rc;
    if (rc == 1) return 0;
    return jf_cg_l2_window(c);
}

int jf_launch_cg(jf_ctx *c, double tol, int64_t maxiter, double step, int apply_update) {
    CgArgs a = {};
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
    // measured in the real kernel on config 2: 7.29 us/iteration with it, 6.78 us with the cooperative barrier (the
    // isolated micro-benchmark says the opposite: 1.6 vs 2.05 us) -- off unless JF_CG_SUMS_ONLY=1
    a.sums_only = getenv("JF_CG_SUMS_ONLY") ? atoi(getenv("JF_CG_SUMS_ONLY")) : 0;
    a.sync_ctr = c->d_sync;
    a.sum_slots = (ulonglong2 *)(c->d_sync + 64);
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
    // counter and flag words of the sums-only barrier restart at zero in every launch (a few KB, same stream)
    JF_CUDA(cudaMemsetAsync(c->d_sync, 0, 256 + (size_t)2 * c->n_sys * c->cg_grid * sizeof(ulonglong2), c->stream));
    // The single-reduction kernel carries r.r by recurrence; its absolute error (~4 eps times the sum of all earlier
    // r.r) is harmless at saenopy's tolerance of 1e-5 but makes the stop test meaningless below ~1e-12, so tight
    // tolerances go through the classic two-reduction kernels.
    const bool fused = c->cg_fused && tol >= 1e-9;
    const bool fused_stream = c->cg_fused_stream && tol >= 1e-9;
    const bool pair = c->cg_pair && tol >= 1e-9;
    int rc = pair ? jf_cg_pair_launch(c, a)
                  : (fused ? jf_cg_fused_launch(c, a)
                           : (c->cg_resident ? jf_cg_resident_launch(c, a)
                                             : (fused_stream ? jf_cg_fused_stream_launch(c, a) : jf_cg_stream_launch(c, a))));
    if (rc) return rc;
    c->stats.kernel_launches++;
    c->stats.cg_launches++;
    return 0;
}
