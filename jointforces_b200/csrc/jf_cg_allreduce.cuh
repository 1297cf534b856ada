// Pieces shared by the single-reduction CG kernels (jf_cg_fused.cu, jf_cg_pair.cu, jf_cg_fused_stream.cu): the shared
// scalar state, reduction groups, the counter barrier, the partial-sum tables and the fixed-point all-reduce through integer
// atomics that doubles as the grid barrier.  Everything lives in an unnamed namespace: one copy per translation unit.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "jf_cg_common.cuh"

namespace {

// -DJF_CG_TIMING: thread 32 (lane 0 of the synchronising warp) of EVERY CTA accumulates clock64 intervals (FMARK) and prints one line at the end
#ifdef JF_CG_TIMING
#define FMARK(i) do { if (threadIdx.x == 32) { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } } while (0)
#define FT_ARGS , long long *tacc, long long &tlast
#define FT_PASS , tacc, tlast
#else
#define FMARK(i) do { } while (0)
#define FT_ARGS
#define FT_PASS
#endif

constexpr int F_JR = 8;         // block columns per lane half in registers; a ninth (slices 17-18 blocks wide: few) lives in shared memory
constexpr int F_HALO_BATCH = 2;  // one round of loads covers 512 foreign halo nodes (config 2: 154-434 per CTA)
#ifndef F_REPLICAS
#define F_REPLICAS 4
#endif
#ifndef JF_ACC_DEFAULT_STRIDE
#define JF_ACC_DEFAULT_STRIDE 1
#endif
// polls of a grid-wide wait before the kernel gives up (tens of seconds): the spin barriers need every CTA of the group
// resident, which a cooperative launch guarantees; jf_cg_pair_kernel with its tail in tensor memory is launched plainly
// (the occupancy calculator grants kernels that allocate tensor memory one CTA per SM, the hardware runs two)
constexpr unsigned JF_SPIN_LIMIT = 1u << 25;
constexpr int F_MAX_SYS = 16;   // systems per launch in this kernel (shared tables are sized by it)

template <int MS>
struct FusedSharedT {
    double warp[3][MS][8];  // per-warp partials of (p.Ap, r.Ap, Ap.Ap)
    double normb[MS], rho[MS], alpha[MS], beta[MS], tot0[MS];
    int done[MS], iters[MS], lastcur[MS];
    // fixed-point all-reduce (one system): exponents of the previous iteration's p.Ap and Ap.Ap, words to add
    int expo[2], expo_ok[2], fallback;  // exponents of the previous p.Ap and Ap.Ap (units of the fixed-point words) and whether they are usable
    unsigned long long word[8];  // 0..5: two limbs of each sum, 6: overflow mark; each with this CTA's arrival in the top bits
#ifdef JF_FX_CHECK
    double dbg[3], dbg_max;
    int dbg_n;
#endif
};
using FusedShared = FusedSharedT<F_MAX_SYS>;

// The CTAs that reduce together: the whole grid, or (jf_cg_pair_kernel) the CTAs of one system -- `n` of them, this
// one being number `i`.  The counters, partial-sum tables and accumulators a group uses come with the CgArgs it passes.
struct Grp {
    unsigned n, i;
    __device__ __forceinline__ Grp() : n(gridDim.x), i(blockIdx.x) {}
    __device__ __forceinline__ Grp(unsigned n_, unsigned i_) : n(n_), i(i_) {}
};

// all-reduce of `nq` values per system + grid barrier (cooperative groups), fixed summation order.
// After it the lane that owns system s's totals runs `update(s, t0, t1, t2)` before the CTA barrier.
// Grid barrier for the iteration loop: monotone arrival counter (zeroed before every launch), release fence before
// the arrival, no L1 invalidation after it -- everything this kernel reads from other CTAs afterwards (halo Ap,
// partial sums) is loaded from L2 (ld.global.cg), and nothing else it reads changes during the launch.
// cg::grid.sync() adds MEMBAR + CCTL.IVALL on the way out (8 % of an iteration in the ncu source view).
__device__ __forceinline__ void light_grid_barrier(unsigned *ctr, unsigned &target, unsigned n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n_ctas;
        __threadfence();  // (red.release.gpu compiles to the same MEMBAR + REDG sequence)
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned seen, spins = 0;
        do {  // relaxed poll; what is read from other CTAs afterwards comes from L2 (ld.global.cg), see fused_fx_allreduce
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            if (++spins > JF_SPIN_LIMIT) __trap();  // a CTA of the group is not resident (cannot happen under a cooperative launch)
        } while (seen < target);
    }
    __syncthreads();
}

// release fence at gpu scope: MEMBAR.ALL.GPU (__threadfence() is fence.sc: MEMBAR.SC.GPU)
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// named CTA barriers of the fixed-point all-reduce (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- all-reduce through integer atomics that double as the grid barrier (one system) --------------------------
// Each of the three sums is accumulated as an 88-bit fixed-point number in two 64-bit words (44 bits each, the high part
// biased to stay non-negative); every addend also adds 1 to the top ten bits of its word, so a word whose top bits have
// advanced by the number of CTAs is complete: arrival and data are the same atomic, and one poll of two sectors replaces
// the barrier's poll plus the separate read of 3 x grid partial sums.  The unit is 2^-23 (high) and 2^-67 (low) of 2^E,
// E = exponent of the previous iteration's total (all CTAs derive it from the same bits), so a partial may be 2^19 times
// larger than the previous total before it does not fit; such a partial raises a mark in a seventh word (which carries
// its own arrival count, so a complete poll has seen every mark) and the iteration is reduced again through the slot
// tables -- written only then.  Integer addition is order independent: bitwise reproducible by construction.
// Precision: every partial is truncated to the low unit (2^-67 of 2^E: far below a double's ulp of the total unless the
// total collapses by more than 2^14 against the previous one -- then the sum still carries 53 - log2(collapse) + 14
// bits, and CG's next iteration re-derives the scale); the decode rounds twice.  tests/test_fixed_point_model.py restates
// encode / accumulate / decode with Python integers.
constexpr int FX_COUNT_SHIFT = 54;  // arrivals in the top 10 bits: grids of up to 1023 CTAs
constexpr int FX_BITS = 44;         // bits per limb: 1023 biased limbs stay below 2^54
constexpr int FX_HEADROOM = 19;
constexpr int FX_HI_SHIFT = FX_BITS - 2 - FX_HEADROOM, FX_LO_SHIFT = FX_HI_SHIFT + FX_BITS;  // unit_hi = 2^(E-23), unit_lo = 2^(E-67)

__device__ __forceinline__ double pow2(int k) { return __longlong_as_double((long long)(k + 1023) << 52); }  // -1022 <= k <= 1023

// e = ilogb(v) + 1 for a positive normal v, straight from the exponent field (a handful of integer instructions on the
// serial path of every iteration); false for zero, negative, subnormal, infinite, NaN or out-of-range values
__device__ __forceinline__ bool fx_exponent_ok(double v, int *e) {
    const int hi = __double2hiint(v);
    const int k = ((hi >> 20) & 0x7ff) - 1022;
    *e = k;
    return hi > 0 && k > -900 + FX_LO_SHIFT && k < 900 - FX_HEADROOM;
}

// first half: publish the CTA's partial sums and pass the grid barrier
template <int NQ, bool LIGHT = false, typename SH>
__device__ __forceinline__ double *fused_publish_sync(const CgArgs &a, cg::grid_group &grid, SH &sh, unsigned &epoch,
                                                      unsigned *light_target = nullptr, const Grp grp = Grp()) {
    epoch++;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // F_REPLICAS copies of the partial table: CTA b reads copy b % F_REPLICAS, so a line has fewer simultaneous readers
    const size_t table = (size_t)3 * a.n_sys * grp.n;
    double *part = a.partial + (size_t)(epoch & 1u) * F_REPLICAS * table;
    __syncthreads();
    for (int v = warp; v < NQ * a.n_sys; v += nwarps) {
        const int q = v / a.n_sys, s = v - q * a.n_sys;
        double x = lane < nwarps ? sh.warp[q][s][lane] : 0.0;
        x = warp_sum(x);
        if (lane < F_REPLICAS) part[lane * table + (size_t)v * grp.n + grp.i] = x;
    }
    if (LIGHT && a.light_barrier) light_grid_barrier(a.sync_ctr, *light_target, grp.n);
    else grid.sync();
    return part;
}

// second half: every CTA adds the partial sums in the same order; the lane that owns system s's totals runs
// `update(s, t0, t1, t2)` before the CTA barrier
template <int NQ, int KMAX, typename SH, typename F>
__device__ __forceinline__ void fused_collect(const CgArgs &a, SH &sh, const double *part, F update, const Grp grp = Grp()) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const size_t table = (size_t)3 * a.n_sys * grp.n;
    for (int s = warp; s < a.n_sys; s += nwarps) {
        // all loads of all sums in flight before the first shuffle (grid <= 32*KMAX CTAs: KMAX per lane and sum)
        double v[3][KMAX];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double *p = part + (grp.i % F_REPLICAS) * table + (size_t)(q * a.n_sys + s) * grp.n;
#pragma unroll
            for (int k = 0; k < KMAX; k++) {
                const int i = lane + 32 * k;
                v[q][k] = i < (int)grp.n ? __ldcg(p + i) : 0.0;
            }
        }
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double x = v[q][0];
#pragma unroll
            for (int k = 1; k < KMAX; k++) x += v[q][k];
            t[q] = warp_sum(x);
        }
        if (lane == 0) update(s, t[0], t[1], t[2]);
    }
    __syncthreads();
}

// all-reduce of `nq` values per system + grid barrier (cooperative groups), fixed summation order
template <int NQ, int KMAX, typename SH, typename F>
__device__ __forceinline__ void fused_allreduce(const CgArgs &a, cg::grid_group &grid, SH &sh, unsigned &epoch, F update,
                                                const Grp grp = Grp()) {
    const double *part = fused_publish_sync<NQ>(a, grid, sh, epoch, nullptr, grp);
    fused_collect<NQ, KMAX>(a, sh, part, update, grp);
}

// alpha, beta, the new r.r, the stop decision and the units of the next fixed-point sums from the three totals of an
// iteration.  With alpha = rho / pAp:
//       beta = r'.r' / rho = 1 - 2 (r.Ap / p.Ap) + alpha (Ap.Ap / p.Ap)        r'.r' = rho beta
// -- three quotients by the same total that do not depend on each other.  ln < 0: one thread does everything (slot-table
// path); ln >= 0: called by a whole warp with ln = lane: lanes 0, 1, 2 divide side by side, lanes 4, 5 take the exponents
// meanwhile and write them themselves; the dependent chain is one division, two shuffles and three multiply-adds.  Both
// forms execute the same operations on the same inputs, so the two paths agree bit for bit.
template <typename SH>
__device__ __forceinline__ void fused_update_scalars(SH &sh, int ss, double pAp, double rAp, double ApAp, int ln, long long it,
                                                     long long maxiter, double tol, bool want_expo, int cur) {
    if (sh.done[ss]) return;
    const double rho = sh.rho[ss];
    double alpha, t1, t2;
    if (ln < 0) {
        alpha = rho / pAp;
        t1 = rAp / pAp;
        t2 = ApAp / pAp;
    } else {
        const double qv = (ln == 1 ? rAp : ln == 2 ? ApAp : rho) / pAp;
        if (want_expo && (ln == 4 || ln == 5)) {  // units of the next iteration's fixed-point sums
            int e = 0;
            const bool ok = fx_exponent_ok(ln == 5 ? ApAp : pAp, &e);
            sh.expo[ln - 4] = e;
            sh.expo_ok[ln - 4] = ok;
        }
        alpha = __shfl_sync(0xffffffffu, qv, 0);
        t1 = __shfl_sync(0xffffffffu, qv, 1);
        t2 = __shfl_sync(0xffffffffu, qv, 2);
    }
    if (ln <= 0) {
        const double beta = fma(alpha, t2, fma(-2.0, t1, 1.0));
        const double rsnew = rho * beta;  // r'.r' without a second pass over the vectors
        sh.alpha[ss] = alpha;
        sh.iters[ss] = (int)it;
        sh.lastcur[ss] = cur;  // streaming kernel: generation that holds this iteration's p
        if (rsnew < tol * sh.normb[ss] || it == maxiter || !(rsnew > 0.0)) {
            sh.done[ss] = 1;  // the pending x += alpha p is applied after the loop
        } else {
            sh.beta[ss] = beta;
            sh.rho[ss] = rsnew;
        }
        if (want_expo && ln < 0) {
            int e0 = 0, e2 = 0;
            sh.expo_ok[0] = fx_exponent_ok(pAp, &e0);
            sh.expo_ok[1] = fx_exponent_ok(ApAp, &e2);
            sh.expo[0] = e0;
            sh.expo[1] = e2;
        }
    }
}

struct SlotCounters { unsigned epoch, light_target; };
struct FxPrev { unsigned long long w0 = 0, w1 = 0; };  // per lane: value of its accumulator word when the set's previous use completed
template <int KMAX, typename SH>
__device__ __noinline__ SlotCounters fused_slot_allreduce(CgArgs a, SH *shp, unsigned epoch, unsigned light_target, long long it,
                                                          bool want_expo, int cur, const Grp grp = Grp()) {
    // (arguments by value and the counters returned: a reference parameter would pin the caller's copy to local memory)
    SH &sh = *shp;
    cg::grid_group grid = cg::this_grid();
    const double *part = fused_publish_sync<3, true>(a, grid, sh, epoch, &light_target, grp);
    fused_collect<3, KMAX>(a, sh, part, [&](int ss, double s0, double s1, double s2) {
        fused_update_scalars(sh, ss, s0, s1, s2, -1, it, a.maxiter, a.tol, want_expo, cur);
    }, grp);
    return SlotCounters{epoch, light_target};
}

// The fixed-point all-reduce of one iteration (one system).  `update(s, pAp, rAp, ApAp, lane)` is called by the first
// lanes of the synchronising warp together and leaves the scalars in shared memory; `request()` runs once in every
// thread and issues the loads of what the thread needs from other CTAs in the next iteration.
//
// Roles (the critical path of an iteration is this function, so its serial pieces run side by side):
//   warp 0  "prep"  adds the warps' partial sums and converts them to fixed-point words
//   warp 1  "sync"  adds the words to the accumulators (integer atomics that carry the arrival count), polls until every
//                   CTA's are in, decodes the totals and forms alpha, beta, the stop decision and the next exponents
//   others          wait for the scalars
// Two orderings of the data other CTAs read afterwards (the Ap rows):
//   fenced = false  the rows validate themselves (epoch tag in their fourth word, written and read with the row as one
//                   32-byte access), so nothing orders them against the atomics: no fence, relaxed poll, and
//                   `request()` runs BEFORE the poll -- the loads travel while the CTAs arrive.
//   fenced = true   producer: stores -> bar.sync -> fence.gpu -> red (same thread); consumer: ld.acquire poll that reads
//                   the complete counts -> bar -> `request()`.  The formally ordered path (JF_CG_FENCED=1, and always
//                   for the streaming kernel, whose rows carry no tag).
// The overflow mark travels in a seventh word that carries its own arrival count, so "complete" covers it.
template <int KMAX, bool NEVER_FENCED = false, typename SH, typename U, typename P>
__device__ __forceinline__ void fused_fx_allreduce(const CgArgs &a, SH &sh, unsigned &epoch, unsigned &xe, FxPrev &prev,
                                                   unsigned &light_target, bool fenced_arg, long long it, bool want_expo, int cur, U update,
                                                   P request, const Grp grp FT_ARGS) {
    const bool fenced = !NEVER_FENCED && fenced_arg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    epoch++;
    xe++;
    const int nwarps = blockDim.x >> 5;
#ifdef JF_FX_CHECK
    const size_t table = (size_t)3 * grp.n;
    double *part = a.partial + (size_t)(epoch & 1u) * F_REPLICAS * table;
#endif
    // Two accumulator sets alternate and are never cleared during a launch: the words only grow, and the sum of one use is
    // the difference to the value the set had when its previous use completed (`prev`, the same in every CTA because a
    // set is complete -- and stays untouched until everyone has read it -- when a CTA leaves the poll).  Nothing has to
    // be ordered against a clearing store.
    unsigned long long *acc = a.acc + (size_t)(xe & 1u) * 8 * a.acc_stride;
    __syncthreads();  // the warps' partials are in sh.warp; every Ap store of this CTA has been issued
    FMARK(3);
    // lanes 8q .. 8q+7 of one warp add the (at most eight) warp partials of sum q -- the same shuffle tree as warp_sum over 8
    // lanes -- and lanes 0, 8, 16 turn the three sums into words; then lanes 0..6 hold the words they add
    auto make_words = [&]() -> unsigned long long {
        const int q = lane >> 3, wsrc = lane & 7;
        double x = (q < 3 && wsrc < nwarps) ? sh.warp[q][0][wsrc] : 0.0;
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
#ifdef JF_FX_CHECK
        if (q < 3 && wsrc < F_REPLICAS) part[wsrc * table + (size_t)q * grp.n + grp.i] = x;
#endif
        bool ok = true;
        long long hi = 0, lo = 0;
        if (q < 3 && wsrc == 0) {
            const int E = sh.expo[q == 2 ? 1 : 0];
            ok = fabs(x) < pow2(E + FX_HEADROOM) &&  // false for NaN
                 (NEVER_FENCED || !(a.fx_test > 0 && grp.i == 0 && xe % (unsigned)a.fx_test == 0));
            if (ok) {
                hi = __double2ll_rd(x * pow2(FX_HI_SHIFT - E));
                const double rem = fma(-(double)hi, pow2(E - FX_HI_SHIFT), x);  // in [0, unit_hi]; rounds only for tiny negative x, by < 2^-9 low units
                lo = __double2ll_rd(rem * pow2(FX_LO_SHIFT - E));
                lo = min(max(lo, 0ll), (1ll << FX_BITS) - 1);
            }
        }
        const unsigned bad = __ballot_sync(0xffffffffu, !ok);
        const unsigned long long whi = (1ull << FX_COUNT_SHIFT) + (unsigned long long)(hi + (1ll << (FX_BITS - 1)));
        const unsigned long long wlo = (1ull << FX_COUNT_SHIFT) + (unsigned long long)lo;
        const int src = (lane >> 1) << 3;  // lane 2q and 2q+1 take the words of sum q from lane 8q
        const unsigned long long a_hi = __shfl_sync(0xffffffffu, whi, src & 31), a_lo = __shfl_sync(0xffffffffu, wlo, src & 31);
        if (lane == 6) return (1ull << FX_COUNT_SHIFT) + (bad ? 1ull : 0ull);  // the overflow mark arrives like a sum
        return (lane & 1) ? a_lo : a_hi;
    };
    if (warp == 0 && fenced) {
        const unsigned long long wd = make_words();
        if (lane < 7) sh.word[lane] = wd;
        __syncwarp();
        named_bar_arrive(1, 64);            // words are in shared memory (bar.arrive orders the stores before it)
        named_bar_sync(2, blockDim.x);      // every CTA has arrived
        request();
    } else if (warp == 1) {
        unsigned long long wd;
        if (fenced) {                       // the fence and the conversion run side by side in two warps
            if (lane == 0) fence_release_gpu();  // this CTA's rows (and slot partials) are in L2 before its words arrive
            __syncwarp();
            FMARK(4);
            named_bar_sync(1, 64);
            wd = lane < 7 ? sh.word[lane] : 0ull;
        } else {
            wd = make_words();              // nothing to wait for: this warp converts and adds
        }
        FMARK(5);
        unsigned long long w = 0, raw = 0;  // lanes 0..6: the accumulator words (6: overflow marks), this use's share of them
        const unsigned long long before = (xe & 1u) ? prev.w1 : prev.w0;
        if (lane < 7) asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc + lane * a.acc_stride), "l"(wd) : "memory");
        if (!fenced) request();
        bool complete;
        unsigned spins = 0;
        do {
            if (++spins > JF_SPIN_LIMIT) __trap();
            if (lane < 7) {
                // (acquire = LDG.STRONG.GPU + CCTL.IVALL: an L1 invalidation per poll, which costs the streaming kernel 13 % --
                //  it keeps its index arrays in L1; everything read from other CTAs afterwards is loaded with ld.global.cg,
                //  i.e. from L2, so the invalidation buys nothing on this hardware.  JF_CG_ACQUIRE=1 asks for it anyway.)
                if (fenced && a.acquire) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(raw) : "l"(acc + lane * a.acc_stride) : "memory");
                else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(raw) : "l"(acc + lane * a.acc_stride) : "memory");
            }
            w = raw - before;
            complete = lane >= 7 || (unsigned)(w >> FX_COUNT_SHIFT) == grp.n;
#ifdef JF_CG_TIMING
            if (threadIdx.x == 32) tacc[9]++;
#endif
        } while (!__all_sync(0xffffffffu, complete));
        FMARK(6);
        if (fenced) named_bar_arrive(2, blockDim.x);  // the barrier proper: the other warps may read what other CTAs wrote
        if (xe & 1u) prev.w1 = raw;
        else prev.w0 = raw;
        FMARK(10);
        double d = 0.0;
        if (lane < 6) {
            const unsigned long long body = w & ((1ull << FX_COUNT_SHIFT) - 1);
            const int E = sh.expo[(lane >> 1) == 2 ? 1 : 0];
            if (lane & 1) d = (double)body * pow2(E - FX_LO_SHIFT);
            else d = (double)((long long)body - (long long)grp.n * (1ll << (FX_BITS - 1))) * pow2(E - FX_HI_SHIFT);
        }
        const double t = d + __shfl_down_sync(0xffffffffu, d, 1);  // lanes 0, 2, 4: high + low part
        const double t0 = __shfl_sync(0xffffffffu, t, 0), t1 = __shfl_sync(0xffffffffu, t, 2), t2 = __shfl_sync(0xffffffffu, t, 4);
        const unsigned long long ovf = __shfl_sync(0xffffffffu, w, 6) & ((1ull << FX_COUNT_SHIFT) - 1);
        if (lane == 0) sh.fallback = ovf ? 1 : 0;
        FMARK(11);
        if (!ovf) update(0, t0, t1, t2, lane);  // the whole warp: the lanes share the serial pieces
        FMARK(8);
        if (fenced) request();  // after the scalars: loads in flight would hold up the scalar chain
#ifdef JF_FX_CHECK
        if (lane == 0) { sh.dbg[0] = t0; sh.dbg[1] = t1; sh.dbg[2] = t2; }
#endif
    } else {
        if (fenced) named_bar_sync(2, blockDim.x);
        request();
    }
    __syncthreads();  // the scalars are in shared memory
    FMARK(7);
    if (sh.fallback) {  // some partial did not fit: this iteration goes through the slot tables (per-warp partials are still in sh.warp)
        const SlotCounters sc = fused_slot_allreduce<KMAX>(a, &sh, epoch, light_target, it, want_expo, cur, grp);
        epoch = sc.epoch;
        light_target = sc.light_target;
    }
#ifdef JF_FX_CHECK
    else {  // debug build: the fixed-point totals against the slot-table sums of the same partials
        fused_collect<3, KMAX>(a, sh, part, [&](int, double s0, double s1, double s2) {
            const double r0 = fabs(sh.dbg[0] - s0) / fabs(s0), r1 = fabs(sh.dbg[1] - s1) / fabs(s1), r2 = fabs(sh.dbg[2] - s2) / fabs(s2);
            sh.dbg_max = fmax(sh.dbg_max, fmax(r0, fmax(r1, r2)));
            sh.dbg_n++;
        }, grp);
    }
#endif
}

#define JF_BLOCK_FMA_TO(z0, z1, z2, vj, STRIDE, px, py, pz) \
    z0 = fma((vj)[0 * (STRIDE)], px, z0);                    \
    z0 = fma((vj)[1 * (STRIDE)], py, z0);                    \
    z0 = fma((vj)[2 * (STRIDE)], pz, z0);                    \
    z1 = fma((vj)[3 * (STRIDE)], px, z1);                    \
    z1 = fma((vj)[4 * (STRIDE)], py, z1);                    \
    z1 = fma((vj)[5 * (STRIDE)], pz, z1);                    \
    z2 = fma((vj)[6 * (STRIDE)], px, z2);                    \
    z2 = fma((vj)[7 * (STRIDE)], py, z2);                    \
    z2 = fma((vj)[8 * (STRIDE)], pz, z2);

constexpr int TM_COLS = 256;                 // columns per CTA: two CTAs per SM share the 512
constexpr int TM_MAX_TAIL = TM_COLS / 18;    // block columns of a row that fit (18 words per 3x3 block of doubles)

}  // namespace

// JF_CG_ACC_STRIDE=n (64-bit words, power of two <= JF_ACC_MAX_STRIDE): the seven accumulator words of a set on separate
// lines, so their atomics are served by different L2 slices instead of queueing on two sectors of one
static inline void fused_place_accumulators(jf_ctx *c, CgArgs &a) {
    const char *e = getenv("JF_CG_ACC_STRIDE");
    int stride = e ? atoi(e) : JF_ACC_DEFAULT_STRIDE;
    if (stride < 1 || stride > JF_ACC_MAX_STRIDE) stride = 1;
    a.acc_stride = stride;
    if (stride == 1) {
        a.acc = (unsigned long long *)((char *)c->d_sync + 64);  // zeroed with the barrier counter (jf_launch_cg)
    } else {
        a.acc = c->d_acc;
        cudaMemsetAsync(c->d_acc, 0, (size_t)16 * stride * sizeof(unsigned long long), c->stream);
    }
}
