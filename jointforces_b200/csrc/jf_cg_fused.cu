// Register-resident CG with ONE grid barrier per iteration (see jf_cg.cu for the CG overview and
// jf_cg_resident.cu for the data placement this kernel shares with jf_cg_regres_kernel).
//
// Classic Hestenes-Stiefel needs two global reductions per iteration: p.Ap (for alpha) and r.r of the
// updated residual (for beta and the stop test).  The second one is algebraically available at the
// time of the first:
//       r' = r - alpha Ap      =>      r'.r' = r.r - 2 alpha (r.Ap) + alpha^2 (Ap.Ap)
// so one all-reduce of (p.Ap, r.Ap, Ap.Ap) yields alpha, the new r.r, the stop decision and beta.
// The iterates are the same as saenopy's cg in exact arithmetic (same alpha, same beta, same stop
// quantity); in floating point r'.r' carries a relative error of ~3 eps * (r.r)/(r'.r'), i.e. 1e-15
// for CG's usual 0.5-1.0 reduction per step -- the size of the summation-order differences between
// any two implementations of the dot product (DESIGN.md "Trajectory sensitivity").
//
// The vector updates that depended on the second barrier are deferred into the next iteration and
// formed on the fly where they are needed:
//   own row (registers):  x += alpha_prev p;  r -= alpha_prev Ap;  p = r + beta_prev p
//   halo node j (once per distinct neighbour; every CTA keeps r_j and p_j of ITS halo nodes in shared
//   memory for the whole solve and advances them itself):
//                          r_j -= alpha_prev Ap_j;  p_j = r_j + beta_prev p_j
// with the same FMA forms on the same inputs, so an owner's registers and its neighbours' halo copies
// agree bit for bit.  The only vector that crosses CTAs is therefore Ap: one 32-byte store per row and
// one 256-bit load per halo node per iteration; two generations alternate by parity.
//
// Measured on config 2: 6.4-6.8 us per iteration with two barriers -> see DESIGN.md / profiles.
#include "jf_cg_allreduce.cuh"

namespace {

// FAST = the default single-solve path as a specialisation of its own: one system, halo requests ahead, fixed-point sums,
// tagged rows (no fence), every CTA's foreign halo within one round of loads.  The iteration is bound by the issue
// latency of a lone warp's dependent instructions (ncu: 2 warps per scheduler, ~6 cycles per instruction, every branch
// ~20), so what the general kernel decides at run time -- loops over systems, rounds of halo loads, mode switches --
// is compiled away here.  Same arithmetic, same bits (test_resident_and_streaming_cg_agree, JF_CG_GENERAL=1 selects the
// general instantiation).
template <bool FAST>
__global__ void __launch_bounds__(256, 1) jf_cg_fused_kernel(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FusedShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = blockDim.x >> 6;  // slices per CTA (two warps per 32-row slice)
    const int s = FAST ? 0 : blockIdx.x / a.ctas_per_sys;
    const int cb = FAST ? blockIdx.x : blockIdx.x - s * a.ctas_per_sys;
    const int half = lane >> 4;
    const int rloc = warp * 16 + (lane & 15);
    const long long sl = (long long)cb * W + (rloc >> 5);
    const bool have = sl < a.n_slices;
    const bool owner = have && half == 0;
    const long long row = sl * 32 + (rloc & 31);
    const long long vstride = a.Nf_pad * 4;
    // halo = every block column this CTA's rows touch: its own rows first (entry i = local row i), then the foreign nodes.
    // r and p of the foreign nodes are advanced locally each iteration; p of the own rows is mirrored from registers.
    // Component-major (x[hmax], y[hmax], z[hmax]): consecutive lanes touch consecutive doubles, no bank conflicts.
    const int hs = a.hmax;
    double *hr = (double *)dyn_smem;        // 3 * hmax: r of the halo nodes
    double *hp = hr + (size_t)hs * 3;       // 3 * hmax: p of the halo nodes
    double *m8 = hp + (size_t)hs * 3;       // 9 * blockDim: every thread's ninth block, component-major
    int *lc8 = (int *)(m8 + 9 * blockDim.x);  // blockDim: its local column
    int *hn = lc8 + blockDim.x;             // hmax: node ids
    const int n_own = min(W, (int)max(0ll, a.n_slices - (long long)cb * W)) * 32;
    const int h0 = a.halo_ptr[cb], nh = a.halo_ptr[cb + 1] - h0;
    unsigned epoch = 0;

    if (threadIdx.x < a.n_sys) {
        sh.done[threadIdx.x] = a.state[threadIdx.x].active ? 0 : 1;
        sh.iters[threadIdx.x] = 0;
        sh.alpha[threadIdx.x] = 0.0;
        sh.beta[threadIdx.x] = 0.0;
    }
    if (threadIdx.x == 0) sh.expo_ok[0] = sh.expo_ok[1] = sh.fallback = 0;
#ifdef JF_FX_CHECK
    if (threadIdx.x == 0) { sh.dbg_max = 0.0; sh.dbg_n = 0; }
#endif
    for (int i = threadIdx.x; i < 3 * F_MAX_SYS * 8; i += blockDim.x) (&sh.warp[0][0][0])[i] = 0.0;
    __syncthreads();
    const bool sys_on = !sh.done[s];

    // the row's blocks and local column indices, once, into registers
    double m[F_JR][9];
    int lc[F_JR];
    int w = 0;
#pragma unroll
    for (int jj = 0; jj < F_JR; jj++) {
        lc[jj] = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) m[jj][q] = 0.0;
    }
    lc8[threadIdx.x] = 0;
#pragma unroll
    for (int q = 0; q < 9; q++) m8[q * blockDim.x + threadIdx.x] = 0.0;
    if (have && sys_on) {
        const int q0 = a.slice_ptr[sl];
        w = (a.slice_ptr[sl + 1] - q0) >> 5;
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288 + (rloc & 31);
#pragma unroll
        for (int jj = 0; jj < F_JR; jj++) {
            const int j = 2 * jj + half;
            if (j < w) {
                lc[jj] = a.lcol[q0 + j * 32 + (rloc & 31)];
#pragma unroll
                for (int q = 0; q < 9; q++) m[jj][q] = v[288 * j + 32 * q];
            }
        }
        const int j8 = 2 * F_JR + half;
        if (j8 < w) {
            lc8[threadIdx.x] = a.lcol[q0 + j8 * 32 + (rloc & 31)];
#pragma unroll
            for (int q = 0; q < 9; q++) m8[q * blockDim.x + threadIdx.x] = v[288 * j8 + 32 * q];
        }
    }
    for (int h = threadIdx.x; h < nh; h += blockDim.x) hn[h] = a.halo_nodes[h0 + h];
    __syncthreads();

    // the only vector that travels between CTAs: Ap of the iteration, two generations.  The fourth word of a row is its
    // epoch tag -- (launch number << 32) | iteration -- written and read with the row as ONE 32-byte access (STG.256 /
    // LDG.256 on one L2 sector), so a reader recognises a row that has not arrived yet and reads it again: nothing has
    // to order the rows against the all-reduce of the iteration (see fused_fx_allreduce).  The launch number keeps a row
    // left by an earlier CG solve at the same iteration from passing.
    double4 *const A_0 = (double4 *)(a.Ap + s * vstride), *const A_1 = (double4 *)(a.Ap1 + s * vstride);
    const long long tag_base = (long long)a.launch_seq << 32;
    auto tag_of = [&](long long iter) { return __longlong_as_double(tag_base | (long long)(unsigned)iter); };
    // r_0 = b and p = 0 for this CTA's halo nodes
    if (sys_on) {
        const double4 *b4 = (const double4 *)(a.b + s * vstride);
        for (int h = threadIdx.x; h < nh; h += blockDim.x) {
            const double4 bv = b4[hn[h]];
            hr[h] = bv.x; hr[hs + h] = bv.y; hr[2 * hs + h] = bv.z;
            hp[h] = 0.0; hp[hs + h] = 0.0; hp[2 * hs + h] = 0.0;
        }
    }
    double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, y0 = 0, y1 = 0, y2 = 0;
    double acc = 0.0;
    if (have && sys_on) {
        const double4 bv = ((const double4 *)(a.b + s * vstride))[row];
        r0 = bv.x; r1 = bv.y; r2 = bv.z;
        if (owner) {
            st_node(A_1 + row, 0.0, 0.0, 0.0, tag_of(0));  // generation "-1": Ap = 0 makes iteration 0 the general case (alpha = beta = 0)
            acc = r0 * r0 + r1 * r1 + r2 * r2;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) sh.warp[0][s][warp] = acc;
    fused_allreduce<1, 5>(a, grid, sh, epoch, [&](int ss, double t0, double, double) {
        sh.normb[ss] = t0;
        sh.rho[ss] = t0;
        sh.tot0[ss] = t0;
        if (!sh.done[ss] && t0 == 0.0) sh.done[ss] = 1;  // cg returns 0 when normb == 0
    });

#ifdef JF_CG_TIMING
    long long tacc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    int cur = 0;
    unsigned light_target = 0;  // arrivals expected at the light barrier so far (thread 0)
    unsigned xe = 0;  // iterations reduced through the fixed-point accumulators so far
    FxPrev fx_prev;
    double4 pre[F_HALO_BATCH];  // Ap of the previous iteration at this thread's first-round halo nodes (generation "-1": 0)
#pragma unroll
    for (int k = 0; k < F_HALO_BATCH; k++) pre[k] = make_double4(0.0, 0.0, 0.0, tag_of(0));
    for (long long it = 1; it <= a.maxiter; it++) {
        if (FAST) {
            if (sh.done[0]) break;
        } else {
            int all = 1;
            for (int q = 0; q < a.n_sys; q++) all &= sh.done[q];
            if (all) break;
        }
        FMARK(0);
        const bool on = FAST || !sh.done[s];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (on) {
            const double alpha = sh.alpha[s], beta = sh.beta[s];  // of the previous iteration (0, 0 for the first)
            const double4 *Ao = cur ? A_0 : A_1;  // Ap of the previous iteration
            double4 *An = cur ? A_1 : A_0;        // this iteration's
            // ---- halo: r_j -= alpha Ap_j ; p_j = r_j + beta p_j, once per distinct block column, r and p kept here.
            // The first round's Ap_j were requested right after the barrier (below), while the sums were collected.
            const long long expect = tag_base | (long long)(unsigned)(it - 1);
            // ---- own row: the deferred updates of the previous iteration (first: it ends the life of y) ------
            if (have) {
                x0 = fma(alpha, p0, x0);
                x1 = fma(alpha, p1, x1);
                x2 = fma(alpha, p2, x2);
                r0 = fma(-alpha, y0, r0);
                r1 = fma(-alpha, y1, r1);
                r2 = fma(-alpha, y2, r2);
                p0 = fma(beta, p0, r0);
                p1 = fma(beta, p1, r1);
                p2 = fma(beta, p2, r2);
                if (half == 0) {  // the row's p for the CTA's own block columns: straight from the registers
                    hp[rloc] = p0;
                    hp[hs + rloc] = p1;
                    hp[2 * hs + rloc] = p2;
                }
            }
            // (the shared-memory loads of a stage are issued before its first store: hr and hp could alias for all the
            //  compiler knows, and a load placed behind a store would wait for it)
            for (int hb = n_own + threadIdx.x; hb < nh; hb += FAST ? (1 << 30) : F_HALO_BATCH * blockDim.x) {
                double4 av[F_HALO_BATCH];
                double rn[F_HALO_BATCH][3], po[F_HALO_BATCH][3];
                int hh[F_HALO_BATCH];
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++) {
                    const int h = hb + k * blockDim.x;
                    hh[k] = h < nh ? h : hb;
                    av[k] = (FAST || (hb == n_own + (int)threadIdx.x && a.prefetch)) ? pre[k] : ld_node_l2(Ao + hn[hh[k]]);
                }
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++)
#pragma unroll
                    for (int q = 0; q < 3; q++) rn[k][q] = hr[q * hs + hh[k]];
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++) {
                    while (__double_as_longlong(av[k].w) != expect) {  // not there yet (rare): read again
                        av[k] = ld_node_l2(Ao + hn[hh[k]]);
#ifdef JF_CG_TIMING
                        if (threadIdx.x == 32) tacc[12]++;
#endif
                    }
                    rn[k][0] = fma(-alpha, av[k].x, rn[k][0]);
                    rn[k][1] = fma(-alpha, av[k].y, rn[k][1]);
                    rn[k][2] = fma(-alpha, av[k].z, rn[k][2]);
                }
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++)
#pragma unroll
                    for (int q = 0; q < 3; q++) po[k][q] = hp[q * hs + hh[k]];
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++) {
                    if (hb + k * (int)blockDim.x < nh) {
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            hr[q * hs + hh[k]] = rn[k][q];
                            hp[q * hs + hh[k]] = fma(beta, po[k][q], rn[k][q]);
                        }
                    }
                }
            }
            FMARK(1);
            __syncthreads();
            // ---- block row: Ap_i = sum_j K_ij p_j ----------------------------------------------------------------
            if (have) {
                y0 = y1 = y2 = 0.0;
                // no test on the slice width here: absent blocks are zero with column 0, and straight-line code lets the
                // loads of the next blocks travel under the multiply-adds of the current one
#pragma unroll
                for (int jj = 0; jj < F_JR; jj++) {
                    const double *pj = hp + lc[jj];
                    const double px = pj[0], py = pj[hs], pz = pj[2 * hs];
                    JF_BLOCK_FMA(m[jj], 1, px, py, pz)
                }
                if (2 * F_JR < w) {  // the slice is 17 or 18 blocks wide: the ninth block from shared memory
                    const double *pj = hp + lc8[threadIdx.x];
                    const double px = pj[0], py = pj[hs], pz = pj[2 * hs];
                    const double *m8t = m8 + threadIdx.x;
                    JF_BLOCK_FMA(m8t, blockDim.x, px, py, pz)
                }
                y0 += __shfl_xor_sync(0xffffffffu, y0, 16);
                y1 += __shfl_xor_sync(0xffffffffu, y1, 16);
                y2 += __shfl_xor_sync(0xffffffffu, y2, 16);
                if (owner) {
                    st_node(An + row, y0, y1, y2, tag_of(it));  // ONE 32-byte store: row and tag arrive together
                    a0 = p0 * y0 + p1 * y1 + p2 * y2;
                    a1 = r0 * y0 + r1 * y1 + r2 * y2;
                    a2 = y0 * y0 + y1 * y1 + y2 * y2;
                }
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
        FMARK(2);
        auto update = [&](int ss, double pAp, double rAp, double ApAp, int ln) {
            fused_update_scalars(sh, ss, pAp, rAp, ApAp, ln, it, a.maxiter, a.tol, true, 0);
        };
        auto prefetch_halo = [&]() {
            // the next iteration's first round of halo loads (this iteration's Ap) travels while the scalars are formed
            if (FAST || (on && nh > 0 && a.prefetch)) {
                const double4 *An_c = cur ? A_1 : A_0;
#pragma unroll
                for (int k = 0; k < F_HALO_BATCH; k++) {
                    const int h = n_own + (int)threadIdx.x + k * (int)blockDim.x;
                    pre[k] = ld_node_l2(An_c + hn[h < nh ? h : 0]);
                }
            }
        };
        if ((FAST || (a.exact_sums && a.n_sys == 1)) && sh.expo_ok[0] && sh.expo_ok[1]) {  // the same decision in every CTA: it comes from shared totals
            fused_fx_allreduce<5, FAST>(a, sh, epoch, xe, fx_prev, light_target, !FAST && a.fenced != 0, it, true, 0, update, prefetch_halo, Grp() FT_PASS);
        } else {
            const SlotCounters sc = fused_slot_allreduce<5>(a, &sh, epoch, light_target, it, true, 0);
            epoch = sc.epoch;
            light_target = sc.light_target;
            prefetch_halo();
            FMARK(7);
        }
        cur ^= 1;
    }
#ifdef JF_CG_TIMING
    if (threadIdx.x == 32) {
        const double n = sh.iters[0] > 0 ? sh.iters[0] : 1;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("cgt cta %d sm %u nh %d iters %d | top %.0f halo %.0f spmv %.0f | sync_a %.0f fence %.0f words %.0f redpoll %.0f arrive %.0f decode %.0f update %.0f sync_b %.0f | polls %.2f retries %.3f\n",
               blockIdx.x, smid, nh, sh.iters[0], tacc[0] / n, tacc[1] / n, tacc[2] / n, tacc[3] / n, tacc[4] / n, tacc[5] / n, tacc[6] / n,
               tacc[10] / n, tacc[11] / n, tacc[8] / n, tacc[7] / n, tacc[9] / n, tacc[12] / n);
    }
#endif

#ifdef JF_FX_CHECK
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("fx-check: %d fixed-point reductions of %d iterations, max rel diff to slot sums %.3e\n", sh.dbg_n, sh.iters[0], sh.dbg_max);
#endif
    // the last iteration's x += alpha p, then (A12) U[free] += step * x ; du = step^2 * sum x^2
    acc = 0.0;
    if (have && a.state[s].active) {
        const double alpha = sh.alpha[s];
        x0 = fma(alpha, p0, x0);
        x1 = fma(alpha, p1, x1);
        x2 = fma(alpha, p2, x2);
        if (owner) {
            ((double4 *)(a.x + s * vstride))[row] = make_double4(x0, x1, x2, 0.0);
            if (a.apply_update && row < a.Nf) {
                double *U = a.U + ((long long)s * a.N + row) * 3;
                U[0] += a.step * x0;
                U[1] += a.step * x1;
                U[2] += a.step * x2;
                acc = x0 * x0 + x1 * x1 + x2 * x2;
            }
        }
    }
    if (a.apply_update) {
        acc = warp_sum(acc);
        if (lane == 0) sh.warp[0][s][warp] = acc;
        fused_allreduce<1, 5>(a, grid, sh, epoch, [&](int ss, double t0, double, double) { sh.tot0[ss] = t0; });
    }
    if (blockIdx.x == 0 && threadIdx.x < a.n_sys && a.state[threadIdx.x].active) {
        const int q = threadIdx.x;
        a.state[q].cg_iters = sh.iters[q];
        a.state[q].normb = sh.normb[q];
        a.state[q].resid = sh.rho[q];
        if (a.apply_update) a.state[q].du = a.step * a.step * sh.tot0[q];
    }
}

}  // namespace

// usable when the register-resident layout applies (decided by jf_cg_resident_configure: cg_resident == 2)
int jf_cg_fused_configure(jf_ctx *c) {
    c->cg_fused = 0;
    // the collect keeps five partial sums per lane in registers: at most 160 CTAs (B200: 148 SMs)
    if (c->cg_resident != 2 || c->n_sys > F_MAX_SYS || c->cg_grid > 160 || getenv("JF_CG_CLASSIC")) return 0;
    const size_t smem = (size_t)c->cg_hmax * (6 * sizeof(double) + sizeof(int)) + (size_t)c->cg_res_warps * 64 * (9 * sizeof(double) + sizeof(int));
    JF_CUDA(cudaFuncSetAttribute(jf_cg_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JF_CUDA(cudaFuncSetAttribute(jf_cg_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, per_sm_fast = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_cg_fused_kernel<false>, c->cg_res_warps * 64, smem));
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fast, jf_cg_fused_kernel<true>, c->cg_res_warps * 64, smem));
    if (per_sm_fast < 1) per_sm = 0;
    if (per_sm < 1) return 0;
    c->cg_fused_smem = smem;
    c->cg_fused = 1;
    return 1;
}

int jf_cg_fused_launch(jf_ctx *c, CgArgs &a) {
    // read per launch (one CG solve = hundreds of iterations): tests switch them within a process
    const bool no_prefetch = getenv("JF_CG_NO_PREFETCH") != nullptr, cg_sync = getenv("JF_CG_CGSYNC") != nullptr,
               slot_sums = getenv("JF_CG_SLOT_SUMS") != nullptr;
    a.prefetch = !no_prefetch;
    a.fenced = getenv("JF_CG_FENCED") != nullptr;  // order the Ap rows by a release fence instead of relying on their tags
    a.acquire = getenv("JF_CG_ACQUIRE") != nullptr;
    a.launch_seq = (int)(++c->cg_launch_seq & 0x7fffffff);
    a.light_barrier = !cg_sync;
    a.exact_sums = !slot_sums && !cg_sync;
    a.fx_test = getenv("JF_FX_TEST_OVERFLOW") ? atoi(getenv("JF_FX_TEST_OVERFLOW")) : 0;
    fused_place_accumulators(c, a);
    a.halo_ptr = c->d_halo_ptr;
    a.halo_nodes = c->d_halo_nodes;
    a.lcol = c->d_lcol;
    a.hmax = c->cg_hmax;
    a.ctas_per_sys = c->cg_ctas_per_sys;
    a.r1 = c->d_r1;
    a.Ap1 = c->d_Ap1;
    void *args[] = {&a};
    // the specialised instantiation where its assumptions hold (see the kernel): one system, default switches, one round
    // of halo loads per thread
    const bool fast = c->n_sys == 1 && a.prefetch && a.exact_sums && a.light_barrier && !a.fenced && a.fx_test == 0 &&
                      c->cg_foreign_max <= F_HALO_BATCH * c->cg_res_warps * 64 && getenv("JF_CG_GENERAL") == nullptr;
    JF_CUDA(cudaLaunchCooperativeKernel(fast ? (void *)jf_cg_fused_kernel<true> : (void *)jf_cg_fused_kernel<false>, dim3(c->cg_grid),
                                        dim3(c->cg_res_warps * 64), args, c->cg_fused_smem, c->stream));
    return 0;
}
