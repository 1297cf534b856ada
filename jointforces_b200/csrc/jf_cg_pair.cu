// Two load cases of one mesh side by side: see the comment at the kernel (split out of jf_cg_fused.cu).
#include "jf_cg_allreduce.cuh"

namespace {

// =================================================================================================
// Two load cases of one mesh side by side (sweeps): jf_cg_pair_kernel.
//
// The single-solve kernel above leaves an SM idle for the ~60 % of an iteration its one CTA spends in the grid-wide
// all-reduce, and it cannot share the SM: its 256 threads x 255 registers are the whole register file.  Here a CTA is
// 128 threads -- ONE thread per block row, four slices -- so two CTAs fit an SM (2 x 128 x 255 registers), and the
// two belong to DIFFERENT systems (the first ctas_per_sys CTAs of the grid are system 0, the rest system 1): while one
// system's CTA waits for its all-reduce the other's computes, the warp schedulers interleave them for free.  Each
// system is its own reduction group (own accumulators, counters and partial-sum tables, `Grp`), so the two CG solves
// of a launch are independent: different iteration counts, no lockstep; only the launch begins and ends together.
// A row keeps its first P_JR blocks in registers as before; the rest of the row (SELL columns P_JR.. of its slice)
// lives in shared memory, component-major per warp (conflict-free 8-byte loads) -- registers and shared memory of the
// 148 SMs together hold both 18 MB matrices of BASELINE configs[2].  Same arithmetic per system as the single-solve
// kernel, but a row is summed by one thread in column order instead of two half rows: the bits of Ap differ in the
// last place from the single-solve kernel's (tests compare the trajectories, not bits).
// =================================================================================================
#ifndef P_JR_REGS
#define P_JR_REGS 8
#endif
#ifndef P_JR_TMEM_REGS
#define P_JR_TMEM_REGS 4
#endif
constexpr int P_JR = P_JR_REGS;  // block columns of a row in registers (tail in shared memory)
// ... with the tail in tensor memory: a block there costs what a register block costs (the p gathers set both), and
// four register blocks leave the kernel without spills: 8 / 6 / 4 in registers: 3.74 / 3.53 / 3.51 us per pair of iterations
constexpr int P_JR_TMEM = P_JR_TMEM_REGS;
constexpr int P_THREADS = 128;   // one thread per row, four 32-row slices per CTA
constexpr int P_HALO_BATCH = 4;  // foreign halo nodes per row thread: one round of loads covers 512 per CTA
#ifndef P_DEFAULT_LANES
#define P_DEFAULT_LANES 1
#endif
#ifndef P_PREFETCH
#define P_PREFETCH 2
#endif
#ifndef P_DEFAULT_TMEM
#define P_DEFAULT_TMEM 1
#endif
constexpr size_t P_PARTIAL_STRIDE = 8192;  // doubles of a.partial per group (2 parities x F_REPLICAS x 3 x <= 160 CTAs fit)



// LANES = threads per block row: 1 (128 threads, 255 registers: a row's P_JR register blocks in one thread) or 2 (256 threads,
// 128 registers: two warps per slice, lane half h takes the SELL columns j with j % 2 == h as in jf_cg_fused_kernel,
// P_JR / 2 of them in registers; twice the warps to hide the latencies of a phase, the same shared-memory traffic).
// TMEM = the tail of the rows in tensor memory instead of shared memory.  With one lane per row the first P_JR_TMEM columns stay
// in registers; with two lanes per row (128 registers per thread) only the first four do -- two per thread -- and up to
// seven more per thread live in the thread's 128 tensor-memory columns (warps w and w + 4 share a lane quarter and take
// one half of the 256 columns each).
template <int LANES, bool TMEM = false>
__global__ void __launch_bounds__(P_THREADS * LANES, 2) jf_cg_pair_kernel(CgArgs a) {
    constexpr int NT = P_THREADS * LANES;      // threads per CTA
    constexpr int RJ = TMEM ? (LANES == 2 ? 4 : P_JR_TMEM) : P_JR;  // SELL columns of a row that live in registers
    constexpr int JR = RJ / LANES;             // register blocks per thread
    constexpr int HB = P_HALO_BATCH / LANES;   // foreign halo nodes per thread
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FusedSharedT<1> sh;
    __shared__ int s_tail[P_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / a.ctas_per_sys;
    const int cb = blockIdx.x - s * a.ctas_per_sys;
    const Grp grp((unsigned)a.ctas_per_sys, (unsigned)cb);
    // this system's view of the arguments: one system, its own counters, tables and accumulators
    CgArgs g = a;
    g.n_sys = 1;
    g.partial = a.partial + (size_t)s * P_PARTIAL_STRIDE;
    g.sync_ctr = a.sync_ctr + s * 32;
    g.acc = a.acc + (size_t)s * 16 * a.acc_stride;
    g.state = a.state + s;
    const int W = P_THREADS / 32;  // slices per CTA
    const int half = LANES == 2 ? lane >> 4 : 0;
    const int rloc = LANES == 2 ? warp * 16 + (lane & 15) : (int)threadIdx.x;  // local row
    const int rl = rloc & 31;                                                  // row within its slice
    const long long sl = (long long)cb * W + (rloc >> 5);
    const bool have = sl < a.n_slices;
    const bool owner = have && half == 0;
    const long long row = sl * 32 + rl;
    const long long vstride = a.Nf_pad * 4;
    const int hs = a.hmax, fs = a.fmax;
    double *hp = (double *)dyn_smem;                 // 3 * hmax: p of the halo nodes (own rows first), component-major
    double *hr = hp + (size_t)hs * 3;                // 3 * fmax: r of the FOREIGN halo nodes
    double *mt = hr + (size_t)fs * 3;                // tail_max * 288: blocks P_JR.. of the CTA's slices, per slice: [column][component][row]
    int *hn = (int *)(mt + (TMEM ? (size_t)0 : (size_t)a.tail_max * 288));   // fmax: node ids of the foreign halo nodes
    unsigned short *lct = (unsigned short *)(hn + fs);  // tail_max * 32: local columns of the tail blocks
    const int n_own = min(W, (int)max(0ll, a.n_slices - (long long)cb * W)) * 32;
    const int h0 = a.halo_ptr[cb], nh = a.halo_ptr[cb + 1] - h0;
    const int nf = nh - n_own;
    unsigned epoch = 0;

    if (threadIdx.x == 0) {
        sh.done[0] = g.state[0].active ? 0 : 1;
        sh.iters[0] = 0;
        sh.alpha[0] = 0.0;
        sh.beta[0] = 0.0;
        sh.expo_ok[0] = sh.expo_ok[1] = sh.fallback = 0;
    }
    for (int i = threadIdx.x; i < 3 * 8; i += NT) (&sh.warp[0][0][0])[i] = 0.0;
    int w = 0, q0 = 0;
    if (have) {
        q0 = a.slice_ptr[sl];
        w = (a.slice_ptr[sl + 1] - q0) >> 5;
    }
    const int wt = max(w - RJ, 0);  // block columns of this row's slice that live in shared / tensor memory
    const int wtu = (wt + LANES - 1) / LANES;  // ... per thread (tensor memory: the same trip count for both lane halves)
    if (rl == 0 && half == 0) s_tail[rloc >> 5] = wt;
    __shared__ unsigned s_tmem;
    if (TMEM) {
        if (warp == 0) {  // one warp allocates; the permit is given up at once so that the SM's other CTA can allocate too
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((unsigned long long)__cvta_generic_to_shared(&s_tmem)),
                         "n"(TM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (TMEM) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this warp's lanes; warps 4..7 (two lanes per row) use the upper half of the columns
    const unsigned tm = TMEM ? s_tmem + ((unsigned)((warp & 3) * 32) << 16) + (warp >= 4 ? TM_COLS / 2 : 0) : 0u;
    unsigned short *lctt = lct + threadIdx.x;  // tensor-memory tail: local columns per thread, [t][NT]
    const bool sys_on = !sh.done[0];
    int toff = 0;
    for (int k = 0; k < (rloc >> 5); k++) toff += s_tail[k];
    double *mtw = mt + (size_t)toff * 288 + rl;
    unsigned short *lctw = lct + toff * 32 + rl;

    double m[JR][9];
    int lc[JR];
#pragma unroll
    for (int jj = 0; jj < JR; jj++) {
        lc[jj] = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) m[jj][q] = 0.0;
    }
    if (have && sys_on) {
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288 + rl;
#pragma unroll
        for (int jj = 0; jj < JR; jj++) {
            const int j = LANES * jj + half;
            if (j < w) {
                lc[jj] = a.lcol[q0 + j * 32 + rl];
#pragma unroll
                for (int q = 0; q < 9; q++) m[jj][q] = v[288 * j + 32 * q];
            }
        }
        if (!TMEM) {
            for (int t = half; t < wt; t += LANES) {
                lctw[t * 32] = a.lcol[q0 + (RJ + t) * 32 + rl];
#pragma unroll
                for (int q = 0; q < 9; q++) mtw[t * 288 + q * 32] = v[288 * (RJ + t) + 32 * q];
            }
        }
    }
    if (TMEM) {  // (warp-collective stores: the slice width is the same for the whole warp; a lane half without a block stores zeros)
        const double *v = a.val + (long long)s * a.n_slots * 9 + (long long)(q0 >> 5) * 288 + rl;
        for (int t = 0; t < wtu; t++) {
            const int j = RJ + LANES * t + half;
            const bool ld = have && sys_on && j < w;
            lctt[t * NT] = ld ? a.lcol[q0 + j * 32 + rl] : (unsigned short)0;
#pragma unroll
            for (int q = 0; q < 9; q++) tmem_st_double(tm + t * 18 + 2 * q, ld ? v[288 * j + 32 * q] : 0.0);
        }
        tmem_wait_st();
    }
    for (int h = threadIdx.x; h < nf; h += NT) hn[h] = a.halo_nodes[h0 + n_own + h];
    __syncthreads();

    double4 *const A_0 = (double4 *)(a.Ap + s * vstride), *const A_1 = (double4 *)(a.Ap1 + s * vstride);
    const long long tag_base = (long long)a.launch_seq << 32;
    auto tag_of = [&](long long iter) { return __longlong_as_double(tag_base | (long long)(unsigned)iter); };
    if (sys_on) {
        const double4 *b4 = (const double4 *)(a.b + s * vstride);
        for (int h = threadIdx.x; h < nf; h += NT) {
            const double4 bv = b4[hn[h]];
            hr[h] = bv.x; hr[fs + h] = bv.y; hr[2 * fs + h] = bv.z;
            hp[n_own + h] = 0.0; hp[hs + n_own + h] = 0.0; hp[2 * hs + n_own + h] = 0.0;
        }
    }
    double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, y0 = 0, y1 = 0, y2 = 0;
    double acc = 0.0;
    if (have && sys_on) {
        const double4 bv = ((const double4 *)(a.b + s * vstride))[row];
        r0 = bv.x; r1 = bv.y; r2 = bv.z;
        if (owner) {
            st_node(A_1 + row, 0.0, 0.0, 0.0, tag_of(0));  // generation "-1": Ap = 0 makes iteration 0 the general case
            acc = r0 * r0 + r1 * r1 + r2 * r2;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) sh.warp[0][0][warp] = acc;
    unsigned light_target = 0, xe = 0;
    // (group-local counter barriers also at the two ends of the solve: the kernel never needs the whole grid, so its
    //  tensor-memory variants can be launched without the cooperative-launch occupancy check)
    {
        const double *part = fused_publish_sync<1, true>(g, grid, sh, epoch, &light_target, grp);
        fused_collect<1, 5>(g, sh, part, [&](int, double t0, double, double) {
            sh.normb[0] = t0;
            sh.rho[0] = t0;
            sh.tot0[0] = t0;
            if (!sh.done[0] && t0 == 0.0) sh.done[0] = 1;  // cg returns 0 when normb == 0
        }, grp);
    }

#ifdef JF_CG_TIMING
    long long tacc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    int cur = 0;
    FxPrev fx_prev;
    // PREFETCH (LANES == 1): the foreign Ap rows of the next iteration are requested before the all-reduce's poll, as in the
    // single-solve kernel -- one L2 round trip less on the CTA's critical path (the period of a lone CTA, not the SM's
    // throughput, bounds the pair: tools/pair_probe.py); with two lanes per row there are no registers for it
    constexpr int NPRE = LANES == 1 ? P_PREFETCH : 0;  // of a thread's HB foreign nodes, how many are requested ahead
    double4 pre[HB];
#pragma unroll
    for (int k = 0; k < HB; k++) pre[k] = make_double4(0.0, 0.0, 0.0, tag_of(0));
    for (long long it = 1; it <= a.maxiter; it++) {
        if (sh.done[0]) break;
        FMARK(0);
        const double alpha = sh.alpha[0], beta = sh.beta[0];  // of the previous iteration (0, 0 for the first)
        const double4 *Ao = cur ? A_0 : A_1;  // Ap of the previous iteration
        double4 *An = cur ? A_1 : A_0;        // this iteration's
        const long long expect = tag_base | (long long)(unsigned)(it - 1);
        // ---- foreign halo: r_j -= alpha Ap_j ; p_j = r_j + beta p_j, r and p kept here
        {
            int hh[HB];
#pragma unroll
            for (int k = 0; k < HB; k++) {
                const int h = (int)threadIdx.x + k * NT;
                hh[k] = h < nf ? h : 0;
                if (k >= NPRE) pre[k] = ld_node_l2(Ao + hn[hh[k]]);
            }
            // ---- own row meanwhile: the deferred updates of the previous iteration ---------------------------------
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
                if (half == 0) {
                    hp[rloc] = p0;
                    hp[hs + rloc] = p1;
                    hp[2 * hs + rloc] = p2;
                }
            }
#pragma unroll
            for (int k = 0; k < HB; k++) {  // one node at a time: few temporaries beside the matrix registers
                if ((int)threadIdx.x + k * NT < nf) {
                    while (__double_as_longlong(pre[k].w) != expect) {  // not there yet (rare): read again
                        pre[k] = ld_node_l2(Ao + hn[hh[k]]);
#ifdef JF_CG_TIMING
                        if (threadIdx.x == 32) tacc[12]++;
#endif
                    }
                    const int h = hh[k];
                    const double q0r = fma(-alpha, pre[k].x, hr[h]), q1r = fma(-alpha, pre[k].y, hr[fs + h]),
                                 q2r = fma(-alpha, pre[k].z, hr[2 * fs + h]);
                    const double q0p = hp[n_own + h], q1p = hp[hs + n_own + h], q2p = hp[2 * hs + n_own + h];
                    hr[h] = q0r;
                    hr[fs + h] = q1r;
                    hr[2 * fs + h] = q2r;
                    hp[n_own + h] = fma(beta, q0p, q0r);
                    hp[hs + n_own + h] = fma(beta, q1p, q1r);
                    hp[2 * hs + n_own + h] = fma(beta, q2p, q2r);
                }
            }
        }
        FMARK(1);
        __syncthreads();
        // ---- block row: Ap_i = sum_j K_ij p_j -- registers first, then the tail of the slice from shared memory -------
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (have) {
            y0 = y1 = y2 = 0.0;
            double z0 = 0.0, z1 = 0.0, z2 = 0.0;
#pragma unroll
            for (int jj = 0; jj < JR; jj += 2) {
                const double *pa = hp + lc[jj], *pb = hp + lc[jj + 1];
                const double ax = pa[0], ay = pa[hs], az = pa[2 * hs];
                const double bx = pb[0], by = pb[hs], bz = pb[2 * hs];
                JF_BLOCK_FMA_TO(y0, y1, y2, m[jj], 1, ax, ay, az)
                JF_BLOCK_FMA_TO(z0, z1, z2, m[jj + 1], 1, bx, by, bz)
            }
            int t = half;
            if (TMEM) {
                int u = 0;
                if (LANES == 1) {
                    for (; u + 1 < wtu; u += 2) {  // two blocks = 36 words per tcgen05.ld pair, while the p gathers go through shared memory
                        unsigned mw[36];
                        tmem_ld_x32(tm + u * 18, mw);
                        tmem_ld_x4(tm + u * 18 + 32, mw + 32);
                        const double *pa = hp + lctt[u * NT], *pb = hp + lctt[(u + 1) * NT];
                        const double ax = pa[0], ay = pa[hs], az = pa[2 * hs];
                        const double bx = pb[0], by = pb[hs], bz = pb[2 * hs];
                        tmem_wait_ld();
                        double md[18];
#pragma unroll
                        for (int q = 0; q < 18; q++) md[q] = __hiloint2double((int)mw[2 * q + 1], (int)mw[2 * q]);
                        JF_BLOCK_FMA_TO(y0, y1, y2, md, 1, ax, ay, az)
                        JF_BLOCK_FMA_TO(z0, z1, z2, (md + 9), 1, bx, by, bz)
                    }
                }
                for (; u < wtu; u++) {
                    unsigned mw[18];
                    tmem_ld_x16(tm + u * 18, mw);
                    tmem_ld_x2(tm + u * 18 + 16, mw + 16);
                    const double *pa = hp + lctt[u * NT];
                    const double ax = pa[0], ay = pa[hs], az = pa[2 * hs];
                    tmem_wait_ld();
                    double md[9];
#pragma unroll
                    for (int q = 0; q < 9; q++) md[q] = __hiloint2double((int)mw[2 * q + 1], (int)mw[2 * q]);
                    if (u & 1) {
                        JF_BLOCK_FMA_TO(z0, z1, z2, md, 1, ax, ay, az)
                    } else {
                        JF_BLOCK_FMA_TO(y0, y1, y2, md, 1, ax, ay, az)
                    }
                }
                t = wt;
            }
            for (; t + LANES < wt; t += 2 * LANES) {
                const double *pa = hp + lctw[t * 32], *pb = hp + lctw[(t + LANES) * 32];
                const double *ma = mtw + t * 288, *mb = ma + LANES * 288;
                const double ax = pa[0], ay = pa[hs], az = pa[2 * hs];
                const double bx = pb[0], by = pb[hs], bz = pb[2 * hs];
                JF_BLOCK_FMA_TO(y0, y1, y2, ma, 32, ax, ay, az)
                JF_BLOCK_FMA_TO(z0, z1, z2, mb, 32, bx, by, bz)
            }
            if (t < wt) {
                const double *pa = hp + lctw[t * 32];
                const double *ma = mtw + t * 288;
                const double ax = pa[0], ay = pa[hs], az = pa[2 * hs];
                JF_BLOCK_FMA_TO(y0, y1, y2, ma, 32, ax, ay, az)
            }
            y0 += z0;
            y1 += z1;
            y2 += z2;
            if (LANES == 2) {
                y0 += __shfl_xor_sync(0xffffffffu, y0, 16);
                y1 += __shfl_xor_sync(0xffffffffu, y1, 16);
                y2 += __shfl_xor_sync(0xffffffffu, y2, 16);
            }
            if (owner) {
                st_node(An + row, y0, y1, y2, tag_of(it));  // ONE 32-byte store: row and tag arrive together
                a0 = p0 * y0 + p1 * y1 + p2 * y2;
                a1 = r0 * y0 + r1 * y1 + r2 * y2;
                a2 = y0 * y0 + y1 * y1 + y2 * y2;
            }
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
            sh.warp[0][0][warp] = a0;
            sh.warp[1][0][warp] = a1;
            sh.warp[2][0][warp] = a2;
        }
        FMARK(2);
        auto update = [&](int ss, double pAp, double rAp, double ApAp, int ln) {
            fused_update_scalars(sh, ss, pAp, rAp, ApAp, ln, it, a.maxiter, a.tol, true, 0);
        };
        auto request = [&]() {
            const double4 *An_c = cur ? A_1 : A_0;
#pragma unroll
            for (int k = 0; k < NPRE; k++) {
                const int h = (int)threadIdx.x + k * NT;
                pre[k] = ld_node_l2(An_c + hn[h < nf ? h : 0]);
            }
        };
        if (sh.expo_ok[0] && sh.expo_ok[1]) {  // the same decision in every CTA of the group: it comes from shared totals
            fused_fx_allreduce<5, true>(g, sh, epoch, xe, fx_prev, light_target, false, it, true, 0, update, request, grp FT_PASS);
        } else {
            const SlotCounters sc = fused_slot_allreduce<5>(g, &sh, epoch, light_target, it, true, 0, grp);
            epoch = sc.epoch;
            light_target = sc.light_target;
            request();
            FMARK(7);
        }
        cur ^= 1;
    }
#ifdef JF_CG_TIMING
    if (threadIdx.x == 32 && sys_on) {
        const double n = sh.iters[0] > 0 ? sh.iters[0] : 1;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("cgp sys %d cta %d sm %u nf %d iters %d | top %.0f halo %.0f spmv %.0f | sync_a %.0f words %.0f redpoll %.0f arrive %.0f decode %.0f update %.0f sync_b %.0f | polls %.2f retries %.3f\n",
               s, cb, smid, nf, sh.iters[0], tacc[0] / n, tacc[1] / n, tacc[2] / n, tacc[3] / n, tacc[5] / n, tacc[6] / n,
               tacc[10] / n, tacc[11] / n, tacc[8] / n, tacc[7] / n, tacc[9] / n, tacc[12] / n);
    }
#endif
    // the last iteration's x += alpha p, then (A12) U[free] += step * x ; du = step^2 * sum x^2
    acc = 0.0;
    if (owner && g.state[0].active) {
        const double alpha = sh.alpha[0];
        x0 = fma(alpha, p0, x0);
        x1 = fma(alpha, p1, x1);
        x2 = fma(alpha, p2, x2);
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
        if (lane == 0) sh.warp[0][0][warp] = acc;
        const double *part = fused_publish_sync<1, true>(g, grid, sh, epoch, &light_target, grp);
        fused_collect<1, 5>(g, sh, part, [&](int, double t0, double, double) { sh.tot0[0] = t0; }, grp);
    }
    if (cb == 0 && threadIdx.x == 0 && g.state[0].active) {
        g.state[0].cg_iters = sh.iters[0];
        g.state[0].normb = sh.normb[0];
        g.state[0].resid = sh.rho[0];
        if (a.apply_update) g.state[0].du = a.step * a.step * sh.tot0[0];
    }
    if (TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TM_COLS) : "memory");
    }
}

}  // namespace

// Two systems side by side (jf_cg_pair_kernel): contexts of exactly two systems whose rows fit four slices per CTA with
// one CTA group of <= 160 CTAs per system and two CTAs per SM -- the sweeps' case on config-2-size meshes.  Builds the
// kernel's own per-CTA halo lists (own rows first, then the sorted foreign block columns; the streaming kernels of the
// same context keep theirs).  Returns 1 (configured), 0 (does not apply) or a negative status.
int jf_cg_pair_configure(jf_ctx *c) {
    c->cg_pair = 0;
    if (c->n_sys != 2 || c->n_slices < 1 || getenv("JF_CG_NO_PAIR") || getenv("JF_CG_CLASSIC") || getenv("JF_CG_CGSYNC") ||
        getenv("JF_CG_SLOT_SUMS") || (c->flags & JF_FLAG_NO_RESIDENT))
        return 0;
    const int W = P_THREADS / 32;
    const int64_t ctas = (c->n_slices + W - 1) / W;
    if (ctas > 160 || 2 * ctas > 2 * (int64_t)c->sm_count) return 0;  // the collect keeps five partial sums per lane; two CTAs per SM
    std::vector<int32_t> hptr((size_t)ctas + 1, 0), hnodes;
    std::vector<uint16_t> lcol((size_t)c->n_slots, 0);
    std::vector<std::vector<int32_t>> lists((size_t)ctas);
    std::vector<int> tails((size_t)ctas, 0);
    jf_parallel_for(ctas, [&](int64_t lo, int64_t hi) {
        for (int64_t cb = lo; cb < hi; cb++) {
            const int64_t s0 = cb * W, s1 = std::min<int64_t>(c->n_slices, s0 + W);
            const int64_t q0 = c->h_slice_ptr[s0], q1 = c->h_slice_ptr[s1];
            const int32_t own0 = (int32_t)(s0 * JF_SLICE), own1 = (int32_t)(s1 * JF_SLICE);
            std::vector<int32_t> &foreign = lists[cb];
            for (int64_t q = q0; q < q1; q++) {
                const int32_t col = c->h_sell_col[q];
                if (col < own0 || col >= own1) foreign.push_back(col);
            }
            std::sort(foreign.begin(), foreign.end());
            foreign.erase(std::unique(foreign.begin(), foreign.end()), foreign.end());
            const int32_t n_own = own1 - own0;
            for (int64_t q = q0; q < q1; q++) {
                const int32_t col = c->h_sell_col[q];
                lcol[q] = (col >= own0 && col < own1)
                              ? (uint16_t)(col - own0)
                              : (uint16_t)(n_own + (std::lower_bound(foreign.begin(), foreign.end(), col) - foreign.begin()));
            }
            int tail = 0;
            for (int64_t sl = s0; sl < s1; sl++) tail += std::max(0, (c->h_slice_ptr[sl + 1] - c->h_slice_ptr[sl]) / JF_SLICE - P_JR);
            tails[cb] = tail;
        }
    }, 8);
    int fmax = 1, tail_max = 1;
    for (int64_t cb = 0; cb < ctas; cb++) {
        const int32_t own0 = (int32_t)(cb * W * JF_SLICE), own1 = (int32_t)(std::min<int64_t>(c->n_slices, (cb + 1) * W) * JF_SLICE);
        for (int32_t i = own0; i < own1; i++) hnodes.push_back(i);
        hnodes.insert(hnodes.end(), lists[cb].begin(), lists[cb].end());
        hptr[cb + 1] = (int32_t)hnodes.size();
        fmax = std::max(fmax, (int)lists[cb].size());
        tail_max = std::max(tail_max, tails[cb]);
    }
    if (fmax > P_HALO_BATCH * P_THREADS) return 0;  // one round of halo loads per thread
    const int hmax = W * JF_SLICE + fmax;
    size_t smem = (size_t)hmax * 3 * sizeof(double) + (size_t)fmax * (3 * sizeof(double) + sizeof(int)) +
                  (size_t)tail_max * (288 * sizeof(double) + 32 * sizeof(unsigned short));
    const char *lanes_env = getenv("JF_CG_PAIR_LANES");
    const int lanes = lanes_env ? (atoi(lanes_env) == 2 ? 2 : 1) : P_DEFAULT_LANES;
    const char *tmem_env = getenv("JF_CG_PAIR_TMEM");
    bool tmem = tmem_env ? atoi(tmem_env) != 0 : P_DEFAULT_TMEM != 0;
    // columns of a row beyond the register blocks, per thread: at most what its tensor-memory columns hold
    const int tm_tail = lanes == 2 ? (c->cg_wmax - 4 + 1) / 2 : c->cg_wmax - P_JR_TMEM;
    if (tm_tail > (lanes == 2 ? TM_MAX_TAIL / 2 : TM_MAX_TAIL)) tmem = false;
    if (tmem)  // no matrix tail in shared memory; the tail's local columns per thread instead of per slice
        smem = (size_t)hmax * 3 * sizeof(double) + (size_t)fmax * (3 * sizeof(double) + sizeof(int)) +
               (size_t)std::max(tm_tail, 1) * P_THREADS * lanes * sizeof(unsigned short);
    const void *kernel = tmem ? (lanes == 2 ? (const void *)jf_cg_pair_kernel<2, true> : (const void *)jf_cg_pair_kernel<1, true>)
                              : (lanes == 2 ? (const void *)jf_cg_pair_kernel<2> : (const void *)jf_cg_pair_kernel<1>);
    const bool verbose = getenv("JF_TIMING") != nullptr;
    cudaError_t fe = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (fe != cudaSuccess) {
        if (verbose) fprintf(stderr, "jf_cg_pair_configure: cudaFuncSetAttribute(%zu bytes): %s\n", smem, cudaGetErrorString(fe));
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    // (the calculator answers 1 for any kernel that allocates tensor memory; registers and shared memory are what limit
    //  these kernels, so the twin without tensor memory is asked with the same launch shape)
    const void *twin = lanes == 2 ? (const void *)jf_cg_pair_kernel<2> : (const void *)jf_cg_pair_kernel<1>;
    if (tmem && cudaFuncSetAttribute(twin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tmem ? twin : kernel, P_THREADS * lanes, smem));
    if (verbose)
        fprintf(stderr, "jf_cg_pair_configure: lanes %d tmem %d, %lld CTAs per system, %zu bytes of shared memory, %d CTAs per SM\n", lanes, (int)tmem,
                (long long)ctas, smem, per_sm);
    if ((int64_t)per_sm * c->sm_count < 2 * ctas) return 0;  // both systems co-resident (two CTAs per SM)
    int rc;
    if ((rc = jf_dev_alloc(c, &c->d_pair_halo_ptr, hptr.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_pair_halo_nodes, hnodes.size()))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_pair_lcol, lcol.size()))) return rc;
    JF_CUDA(cudaMemcpy(c->d_pair_halo_ptr, hptr.data(), hptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_pair_halo_nodes, hnodes.data(), hnodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    JF_CUDA(cudaMemcpy(c->d_pair_lcol, lcol.data(), lcol.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(hptr.size() * 4 + hnodes.size() * 4 + lcol.size() * 2);
    c->cg_pair = lanes + (tmem ? 2 : 0);  // 1, 2: lanes per row, tail in shared memory; 3, 4: the same with the tail in tensor memory
    c->cg_pair_ctas = (int)ctas;
    c->cg_pair_smem = smem;
    c->cg_pair_hmax = hmax;
    c->cg_pair_fmax = fmax;
    c->cg_pair_tail = tail_max;
    return 1;
}

int jf_cg_pair_launch(jf_ctx *c, CgArgs &a) {
    a.prefetch = 1;
    a.fenced = 0;
    a.acquire = 0;
    a.launch_seq = (int)(++c->cg_launch_seq & 0x7fffffff);
    a.light_barrier = 1;
    a.exact_sums = 1;
    a.fx_test = 0;
    // accumulators of the two groups (16 words each) at the start of d_acc, their barrier counters 512 bytes in
    a.acc_stride = 1;
    a.acc = c->d_acc;
    a.sync_ctr = (unsigned *)(c->d_acc + 64);
    JF_CUDA(cudaMemsetAsync(c->d_acc, 0, 1024, c->stream));
    a.halo_ptr = c->d_pair_halo_ptr;
    a.halo_nodes = c->d_pair_halo_nodes;
    a.lcol = c->d_pair_lcol;
    a.hmax = c->cg_pair_hmax;
    a.fmax = c->cg_pair_fmax;
    a.tail_max = c->cg_pair_tail;
    a.ctas_per_sys = c->cg_pair_ctas;
    a.Ap1 = c->d_Ap1;
    void *args[] = {&a};
    void *kernel = c->cg_pair == 4   ? (void *)jf_cg_pair_kernel<2, true>
                   : c->cg_pair == 3 ? (void *)jf_cg_pair_kernel<1, true>
                   : c->cg_pair == 2 ? (void *)jf_cg_pair_kernel<2>
                                     : (void *)jf_cg_pair_kernel<1>;
    const int lanes = (c->cg_pair - 1) % 2 + 1;
    if (c->cg_pair >= 3)  // tensor memory: plain launch (see JF_SPIN_LIMIT); 2 * ctas <= two CTAs per SM was checked at configure time
        JF_CUDA(cudaLaunchKernel(kernel, dim3(2 * c->cg_pair_ctas), dim3(P_THREADS * lanes), args, c->cg_pair_smem, c->stream));
    else
        JF_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(2 * c->cg_pair_ctas), dim3(P_THREADS * lanes), args, c->cg_pair_smem, c->stream));
    return 0;
}
