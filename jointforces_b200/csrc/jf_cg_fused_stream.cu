// Single-reduction CG for streamed matrices: see the comment at the kernel (split out of jf_cg_fused.cu).
#include "jf_cg_allreduce.cuh"

namespace {

// =================================================================================================
// Single-reduction CG for matrices that are streamed (mid-size meshes, lockstep batches): same algebra as
// above, vectors in global memory.  Groups of FS_SLICES slices; per group the halo p_j = (R_j - alpha AP_j) +
// beta P_j is formed once per distinct node from the previous generation's published (r, Ap, p) and kept in
// shared memory for the group's block-row products; the owner pass advances x, r, p of its rows and publishes
// the new generation.  One grid barrier per iteration instead of two and no separate vector-update pass.
// =================================================================================================
#ifndef FS_DEFAULT_TMEM
#define FS_DEFAULT_TMEM 1
#endif
constexpr int FS_SLICES = 8;
constexpr int FS_THREADS = FS_SLICES * 32;

// TMEM = a CTA keeps the first FS_TM_BLOCKS block columns of its FIRST group of slices in tensor memory for the whole solve
// (one system only): with one group per CTA -- BASELINE configs[3]: 300 groups on 296 CTAs -- that is 7 of the ~15 blocks of
// every row that no longer stream from L2 in every iteration (84 MB matrix: 39 MB in tensor memory, 45 MB streamed).
// Plain launch, group barriers at both ends and the spin watchdog as for jf_cg_pair_kernel.
constexpr int FS_TM_BLOCKS = (TM_COLS / 2) / 18;  // warps w and w + 4 share a lane quarter: 128 columns = 7 blocks per thread

template <bool TMEM>
__global__ void __launch_bounds__(FS_THREADS, 2) jf_cg_fused_stream_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double hv[];  // hmax * 3
    __shared__ FusedShared sh;
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long wg = (long long)warp * gridDim.x + blockIdx.x;
    const long long nwg = (long long)gridDim.x * nwarps;
    const long long vstride = a.Nf_pad * 4;
    const long long n_groups = (a.n_slices + FS_SLICES - 1) / FS_SLICES;
    unsigned epoch = 0;

    if (threadIdx.x < a.n_sys) {
        sh.done[threadIdx.x] = a.state[threadIdx.x].active ? 0 : 1;
        sh.iters[threadIdx.x] = 0;
        sh.alpha[threadIdx.x] = 0.0;
        sh.beta[threadIdx.x] = 0.0;
        sh.lastcur[threadIdx.x] = 1;
    }
    if (threadIdx.x == 0) sh.expo_ok[0] = sh.expo_ok[1] = sh.fallback = 0;
    for (int i = threadIdx.x; i < 3 * F_MAX_SYS * 8; i += blockDim.x) (&sh.warp[0][0][0])[i] = 0.0;
    if (TMEM) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((unsigned long long)__cvta_generic_to_shared(&s_tmem)),
                         "n"(TM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    int tmc = 0;  // block columns of this warp's slice of the CTA's first group that live in tensor memory
    int smc = 0;  // ... and the next ones in shared memory behind the halo: [column][component][thread], conflict-free
    double *msm = hv + (size_t)a.hmax * 3 + threadIdx.x;
    unsigned tm = 0;
    if (TMEM) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm = s_tmem + ((unsigned)((warp & 3) * 32) << 16) + (warp >= 4 ? TM_COLS / 2 : 0);
        const long long sl0 = (long long)blockIdx.x * FS_SLICES + warp;
        if (sl0 < a.n_slices && !sh.done[0]) {
            const int q0 = a.slice_ptr[sl0];
            tmc = min((a.slice_ptr[sl0 + 1] - q0) >> 5, FS_TM_BLOCKS);
            const double *v = a.val + (long long)(q0 >> 5) * 288 + lane;
            for (int t = 0; t < tmc; t++) {
#pragma unroll
                for (int q = 0; q < 9; q++) tmem_st_double(tm + t * 18 + 2 * q, v[288 * t + 32 * q]);
            }
        }
        tmem_wait_st();
    }
    if (a.fs_sm_blocks > 0 && a.n_sys == 1) {
        const long long sl0 = (long long)blockIdx.x * FS_SLICES + warp;
        if (sl0 < a.n_slices && !sh.done[0]) {
            const int q0 = a.slice_ptr[sl0];
            smc = min(((a.slice_ptr[sl0 + 1] - q0) >> 5) - tmc, a.fs_sm_blocks);
            const double *v = a.val + (long long)(q0 >> 5) * 288 + lane;
            for (int t = 0; t < smc; t++) {
#pragma unroll
                for (int q = 0; q < 9; q++) msm[(t * 9 + q) * FS_THREADS] = v[288 * (tmc + t) + 32 * q];
            }
        }
    }

    // generation "-1": r_0 = b, p = 0, Ap = 0; x = 0
    for (int s = 0; s < a.n_sys; s++) {
        double acc = 0.0;
        if (!sh.done[s]) {
            const double4 *b4 = (const double4 *)(a.b + s * vstride);
            double4 *x4 = (double4 *)(a.x + s * vstride), *R1 = (double4 *)(a.r1 + s * vstride), *P1 = (double4 *)(a.p1 + s * vstride),
                    *A1 = (double4 *)(a.Ap1 + s * vstride);
            for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                const long long row = sl * 32 + lane;
                const double4 bv = b4[row];
                x4[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                P1[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                A1[row] = make_double4(0.0, 0.0, 0.0, 0.0);
                R1[row] = bv;
                acc += bv.x * bv.x + bv.y * bv.y + bv.z * bv.z;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[0][s][warp] = acc;
    }
    unsigned light_target = 0;
    {
        auto first = [&](int ss, double t0, double, double) {
            sh.normb[ss] = t0;
            sh.rho[ss] = t0;
            sh.tot0[ss] = t0;
            if (!sh.done[ss] && t0 == 0.0) sh.done[ss] = 1;
        };
        if (TMEM) {  // plain launch: counter barrier instead of grid.sync
            const double *part = fused_publish_sync<1, true>(a, grid, sh, epoch, &light_target);
            fused_collect<1, 10>(a, sh, part, first);
        } else {
            fused_allreduce<1, 10>(a, grid, sh, epoch, first);
        }
    }
    unsigned xe = 0;
    FxPrev fx_prev;
    int cur = 0;
    for (long long it = 1; it <= a.maxiter; it++) {
        int all = 1;
        for (int q = 0; q < a.n_sys; q++) all &= sh.done[q];
        if (all) break;
        for (int s = 0; s < a.n_sys; s++) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (!sh.done[s]) {
                const double alpha = sh.alpha[s], beta = sh.beta[s];
                const double4 *Ro = (const double4 *)((cur ? a.r : a.r1) + s * vstride);
                const double4 *Po = (const double4 *)((cur ? a.p0 : a.p1) + s * vstride);
                const double4 *Ao = (const double4 *)((cur ? a.Ap : a.Ap1) + s * vstride);
                double4 *Rn = (double4 *)((cur ? a.r1 : a.r) + s * vstride);
                double4 *Pn = (double4 *)((cur ? a.p1 : a.p0) + s * vstride);
                double4 *An = (double4 *)((cur ? a.Ap1 : a.Ap) + s * vstride);
                double4 *X4 = (double4 *)(a.x + s * vstride);
                const double *vals = a.val + (long long)s * a.n_slots * 9;
                for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
                    const int h0 = a.halo_ptr[g], nh = a.halo_ptr[g + 1] - h0;
                    for (int hb = threadIdx.x; hb < nh; hb += 3 * FS_THREADS) {
                        int node[3];
                        double4 rv[3], pv[3], av[3];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int h = hb + k * FS_THREADS;
                            node[k] = a.halo_nodes[h0 + (h < nh ? h : hb)];
                        }
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            rv[k] = ld_node_l2(Ro + node[k]);  // other CTAs' rows: from L2 (the loop's barrier keeps L1)
                            av[k] = ld_node_l2(Ao + node[k]);
                            pv[k] = ld_node_l2(Po + node[k]);
                        }
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int h = hb + k * FS_THREADS;
                            if (h < nh) {
                                hv[3 * h] = fma(beta, pv[k].x, fma(-alpha, av[k].x, rv[k].x));
                                hv[3 * h + 1] = fma(beta, pv[k].y, fma(-alpha, av[k].y, rv[k].y));
                                hv[3 * h + 2] = fma(beta, pv[k].z, fma(-alpha, av[k].z, rv[k].z));
                            }
                        }
                    }
                    __syncthreads();
                    const long long sl = g * FS_SLICES + warp;
                    if (sl < a.n_slices) {
                        const long long row = sl * 32 + lane;
                        const int q0 = a.slice_ptr[sl];
                        const int w = (a.slice_ptr[sl + 1] - q0) >> 5;
                        const double *v = vals + (long long)(q0 >> 5) * 288 + lane;
                        const unsigned short *lc = a.lcol + q0 + lane;
                        const double4 ro = Ro[row], ao = Ao[row], po = Po[row];
                        double4 xv = X4[row];
                        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                        int jfirst = 0;
                        if (TMEM && g == (long long)blockIdx.x) {  // this group's first columns live in tensor memory
                            for (int t = 0; t < tmc; t++) {
                                unsigned mw[18];
                                tmem_ld_x16(tm + t * 18, mw);
                                tmem_ld_x2(tm + t * 18 + 16, mw + 16);
                                const double *pj = hv + 3 * lc[32 * t];
                                const double px = pj[0], py = pj[1], pz = pj[2];
                                tmem_wait_ld();
                                double md[9];
#pragma unroll
                                for (int q = 0; q < 9; q++) md[q] = __hiloint2double((int)mw[2 * q + 1], (int)mw[2 * q]);
                                JF_BLOCK_FMA(md, 1, px, py, pz)
                            }
                            jfirst = tmc;
                        }
                        if (g == (long long)blockIdx.x) {  // the next columns of the group from shared memory
                            for (int t = 0; t < smc; t++) {
                                const double *pj = hv + 3 * lc[32 * (jfirst + t)];
                                const double px = pj[0], py = pj[1], pz = pj[2];
                                const double *mm = msm + t * 9 * FS_THREADS;
                                JF_BLOCK_FMA(mm, FS_THREADS, px, py, pz)
                            }
                            jfirst += smc;
                        }
                        for (int j0 = jfirst; j0 < w; j0 += STR_BATCH) {
                            int c[STR_BATCH];
                            double m[STR_BATCH][9];
#pragma unroll
                            for (int k = 0; k < STR_BATCH; k++) {
                                const int jj = (j0 + k < w) ? j0 + k : j0;
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
                        // the deferred updates of the previous iteration, then this iteration's publication
                        xv.x = fma(alpha, po.x, xv.x);
                        xv.y = fma(alpha, po.y, xv.y);
                        xv.z = fma(alpha, po.z, xv.z);
                        const double r0 = fma(-alpha, ao.x, ro.x), r1 = fma(-alpha, ao.y, ro.y), r2 = fma(-alpha, ao.z, ro.z);
                        const double p0 = fma(beta, po.x, r0), p1 = fma(beta, po.y, r1), p2 = fma(beta, po.z, r2);
                        X4[row] = xv;
                        Rn[row] = make_double4(r0, r1, r2, 0.0);
                        Pn[row] = make_double4(p0, p1, p2, 0.0);
                        An[row] = make_double4(y0, y1, y2, 0.0);
                        a0 += p0 * y0 + p1 * y1 + p2 * y2;
                        a1 += r0 * y0 + r1 * y1 + r2 * y2;
                        a2 += y0 * y0 + y1 * y1 + y2 * y2;
                    }
                    __syncthreads();
                }
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            a2 = warp_sum(a2);
            if (lane == 0) {
                sh.warp[0][s][warp] = a0;
                sh.warp[1][s][warp] = a1;
                sh.warp[2][s][warp] = a2;
            }
        }
        auto update = [&](int ss, double pAp, double rAp, double ApAp, int ln) {
            fused_update_scalars(sh, ss, pAp, rAp, ApAp, ln, it, a.maxiter, a.tol, a.n_sys == 1, cur);
        };
        if (a.exact_sums && a.n_sys == 1 && sh.expo_ok[0] && sh.expo_ok[1]) {
#ifdef JF_CG_TIMING
            long long tacc[14], tlast = 0;
#endif
            // rows of r, p and Ap cross CTAs here and carry no tag: always the fenced ordering
            fused_fx_allreduce<10>(a, sh, epoch, xe, fx_prev, light_target, true, it, a.n_sys == 1, cur, update, [] {}, Grp() FT_PASS);
        } else {
            const SlotCounters sc = fused_slot_allreduce<10>(a, &sh, epoch, light_target, it, a.n_sys == 1, cur);
            epoch = sc.epoch;
            light_target = sc.light_target;
        }
        cur ^= 1;
    }

    // the last iteration's x += alpha p, then (A12) U[free] += step * x ; du = step^2 * sum x^2
    for (int s = 0; s < a.n_sys; s++) {
        double acc = 0.0;
        if (a.state[s].active) {
            const double alpha = sh.alpha[s];
            const double4 *P4 = (const double4 *)((sh.lastcur[s] ? a.p1 : a.p0) + s * vstride);
            double4 *X4 = (double4 *)(a.x + s * vstride);
            double *U = a.U + (long long)s * a.N * 3;
            for (long long sl = wg; sl < a.n_slices; sl += nwg) {
                const long long row = sl * 32 + lane;
                const double4 pv = P4[row];
                double4 xv = X4[row];
                xv.x = fma(alpha, pv.x, xv.x);
                xv.y = fma(alpha, pv.y, xv.y);
                xv.z = fma(alpha, pv.z, xv.z);
                X4[row] = xv;
                if (a.apply_update && row < a.Nf) {
                    U[3 * row] += a.step * xv.x;
                    U[3 * row + 1] += a.step * xv.y;
                    U[3 * row + 2] += a.step * xv.z;
                    acc += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z;
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[0][s][warp] = acc;
    }
    if (a.apply_update) {
        auto last = [&](int ss, double t0, double, double) { sh.tot0[ss] = t0; };
        if (TMEM) {
            const double *part = fused_publish_sync<1, true>(a, grid, sh, epoch, &light_target);
            fused_collect<1, 10>(a, sh, part, last);
        } else {
            fused_allreduce<1, 10>(a, grid, sh, epoch, last);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < a.n_sys && a.state[threadIdx.x].active) {
        const int q = threadIdx.x;
        a.state[q].cg_iters = sh.iters[q];
        a.state[q].normb = sh.normb[q];
        a.state[q].resid = sh.rho[q];
        if (a.apply_update) a.state[q].du = a.step * a.step * sh.tot0[q];
    }
    if (TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TM_COLS) : "memory");
    }
}

}  // namespace

// streaming counterpart: needs the group halo lists of the halo-staged streaming kernel (cg_stream_halo)
int jf_cg_fused_stream_configure(jf_ctx *c) {
    c->cg_fused_stream = 0;
    c->cg_fused_stream_tmem = 0;
    if (!c->cg_stream_halo || c->n_sys > F_MAX_SYS || getenv("JF_CG_CLASSIC")) return 0;
    size_t smem = (size_t)c->cg_hmax * 3 * sizeof(double);
    // one system: the shared memory two CTAs per SM leave free holds further block columns of each CTA's first group
    c->cg_fs_sm_blocks = 0;
    if (c->n_sys == 1 && !getenv("JF_CG_STREAM_NO_SMEM_CACHE")) {
        const size_t per_block = (size_t)9 * FS_THREADS * sizeof(double), budget = 106 * 1024;
        if (smem + per_block <= budget) c->cg_fs_sm_blocks = (int)std::min<size_t>(8, (budget - smem) / per_block);
        smem += c->cg_fs_sm_blocks * per_block;
    }
    JF_CUDA(cudaFuncSetAttribute(jf_cg_fused_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JF_CUDA(cudaFuncSetAttribute(jf_cg_fused_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_fused_stream_kernel<false>, FS_THREADS, smem));
    if (per_sm < 2 && c->cg_fs_sm_blocks) {  // never trade the second CTA per SM for the cache
        smem -= c->cg_fs_sm_blocks * (size_t)9 * FS_THREADS * sizeof(double);
        c->cg_fs_sm_blocks = 0;
        JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_fused_stream_kernel<false>, FS_THREADS, smem));
    }
    if (per_sm < 1) return 0;
    const long long n_groups = (c->n_slices + FS_SLICES - 1) / FS_SLICES;
    long long grid = std::min<long long>((long long)c->sm_count * per_sm, std::max<long long>(n_groups, c->sm_count));
    if (grid > 320) grid = 320;  // the collect keeps ten partial sums per lane
    c->cg_fused_stream_grid = (int)grid;
    c->cg_fused_smem = smem;
    c->cg_fused_stream = 1;
    // tensor-memory cache of each CTA's first group (one system, default switches, two CTAs per SM for 256 columns each).
    // The occupancy calculator answers 1 for kernels that allocate tensor memory (see jf_cg_pair_configure): the twin's
    // answer above stands in, and the kernel is launched plainly.
    const char *e = getenv("JF_CG_STREAM_TMEM");
    if (c->n_sys == 1 && per_sm <= 2 && (e ? atoi(e) != 0 : FS_DEFAULT_TMEM != 0) && !getenv("JF_CG_CGSYNC") && !getenv("JF_CG_SLOT_SUMS"))
        c->cg_fused_stream_tmem = 1;
    return 1;
}

int jf_cg_fused_stream_launch(jf_ctx *c, CgArgs &a) {
    a.light_barrier = getenv("JF_CG_CGSYNC") == nullptr;
    a.exact_sums = a.light_barrier && getenv("JF_CG_SLOT_SUMS") == nullptr;
    a.fx_test = getenv("JF_FX_TEST_OVERFLOW") ? atoi(getenv("JF_FX_TEST_OVERFLOW")) : 0;
    a.acquire = getenv("JF_CG_ACQUIRE") != nullptr;
    fused_place_accumulators(c, a);
    a.halo_ptr = c->d_halo_ptr;
    a.halo_nodes = c->d_halo_nodes;
    a.lcol = c->d_lcol;
    a.hmax = c->cg_hmax;
    a.r1 = c->d_r1;
    a.Ap1 = c->d_Ap1;
    void *args[] = {&a};
    a.fs_sm_blocks = c->cg_fs_sm_blocks;
    if (c->cg_fused_stream_tmem && a.light_barrier && a.exact_sums)
        JF_CUDA(cudaLaunchKernel((void *)jf_cg_fused_stream_kernel<true>, dim3(c->cg_fused_stream_grid), dim3(FS_THREADS), args, c->cg_fused_smem,
                                 c->stream));
    else
        JF_CUDA(cudaLaunchCooperativeKernel((void *)jf_cg_fused_stream_kernel<false>, dim3(c->cg_fused_stream_grid), dim3(FS_THREADS), args,
                                            c->cg_fused_smem, c->stream));
    return 0;
}
