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

// JF_CG_ACC_STRIDE=n (64-bit words, power of two <= JF_ACC_MAX_STRIDE): the seven accumulator words of a set on separate
// lines, so their atomics are served by different L2 slices instead of queueing on two sectors of one
static void fused_place_accumulators(jf_ctx *c, CgArgs &a) {
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
