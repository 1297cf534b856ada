// Shared pieces of the CG kernels (jf_cg_resident.cu, jf_cg_stream.cu): arguments, the barrier +
// all-reduce primitive, the per-system scalar state machine of Hestenes-Stiefel CG.
#pragma once
#include <cooperative_groups.h>

#include "jf_internal.h"

namespace cg = cooperative_groups;

#ifdef JF_CG_TIMING
#include <cstdio>
#define TMARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } } while (0)
#else
#define TMARK(i) do { } while (0)
#endif

constexpr int MAXW = 32;      // warps per CTA upper bound (shared partial table stride)
constexpr int RES_WARPS = 8;  // resident kernel: at most 8 warps per CTA -> 255 registers per thread
constexpr int STR_BATCH = 3;  // block columns gathered per batch (streaming)

struct CgArgs {
    const int32_t *slice_ptr;
    const int32_t *sell_col;
    const double *val;
    const double *b;
    double *x, *r, *p0, *p1, *Ap;
    double *r1, *Ap1;  // second generation of r and Ap (single-reduction kernel)
    double *partial;  // (2, n_sys, grid) per-CTA partial sums, double-buffered by barrier parity
    int sums_only;          // 1: p.Ap reduction through the fence-free sums-only barrier
    unsigned *sync_ctr;     // arrival counter of the sums-only barrier (zeroed before every launch)
    ulonglong2 *sum_slots;  // (2, n_sys, grid) flag-carrying partial sums of the sums-only barrier
    jf_sys_state *state;
    double *U;
    long long N, Nf, Nf_pad, n_slices, n_slots;
    int n_sys;
    int wmax;  // widest slice (resident kernel: shared-memory stride)
    // resident kernel only: per-CTA halo (sorted unique block columns its slices touch)
    const int32_t *halo_ptr;    // (ctas_per_sys + 1)
    const int32_t *halo_nodes;  // concatenated halo lists
    const uint16_t *lcol;       // (n_slots) block column as index into the owning CTA's halo list
    int hmax, ctas_per_sys;
    int fs_sm_blocks = 0;        // jf_cg_fused_stream_kernel: block columns of a CTA's first group cached in shared memory
    int fmax = 0, tail_max = 0;  // jf_cg_pair_kernel: most foreign halo nodes / most shared-memory block columns of any CTA
    double tol;
    long long maxiter;
    double step;
    int apply_update;
    unsigned long long *acc = nullptr;  // single-reduction kernel: 3 x 8 fixed-point accumulators (zeroed before every launch)
    int fx_test = 0;            // JF_FX_TEST_OVERFLOW=n: CTA 0 declares every n-th fixed-point reduction unrepresentable (tests)
    int exact_sums = 1;         // sums through integer atomics that double as the barrier (JF_CG_SLOT_SUMS=1: slot tables)
    int light_barrier = 1;      // single-reduction kernel: counter barrier without the L1 invalidation (JF_CG_CGSYNC=1: grid.sync)
    int prefetch = 1;           // single-reduction kernel: request the halo Ap ahead of its use (JF_CG_NO_PREFETCH=1: off)
    int fenced = 0;             // register-resident single-reduction kernel: fence/acquire ordering of the Ap rows (JF_CG_FENCED=1)
    int acc_stride = 1;         // 64-bit words between two accumulator words of the fixed-point all-reduce (JF_CG_ACC_STRIDE)
    int acquire = 0;            // fenced ordering: poll with ld.acquire (L1 invalidation per poll) instead of ld.relaxed (JF_CG_ACQUIRE=1)
    int launch_seq = 0;         // CG launches of this context so far: upper half of the Ap rows' epoch tag
};

// one 32-byte vector entry with a single 256-bit load (LDG.E.256 on sm_100a): a gathered node costs
// one L1tex wavefront per distinct line instead of two
__device__ __forceinline__ double4 ld_node(const double4 *p) {
    double4 v;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// the same load served by L2 (never by this SM's L1): for data other CTAs wrote in this launch when the barrier in
// between does not invalidate L1
__device__ __forceinline__ double4 ld_node_l2(const double4 *p) {
    double4 v;
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// one 32-byte vector entry with a single 256-bit store (STG.E.256): `*p = make_double4(..)` compiles to two 128-bit
// stores (double4 is 16-byte aligned), which a concurrent reader can see half done
__device__ __forceinline__ void st_node(double4 *p, double x, double y, double z, double w) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
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

// All-reduce WITHOUT memory-ordering duties, for the p.Ap reduction: what follows it (x += alpha p,
// r -= alpha Ap) touches only the CTA's own rows, so no other CTA's vector writes have to become
// visible -- only the sums.  One relaxed arrival counter (polled by one lane per CTA) says "everyone has
// published"; the sums travel in 16-byte slots whose two halves each carry the 32-bit epoch, so a slot is
// self-validating and needs no ordering against the counter (re-read until both tags match).  Saves
// the two MEMBAR.SC + CCTL.IVALL of grid.sync: 1.6 us instead of 2.05 us on 134 CTAs
// (tools/micro/barrier_bench.cu).  Deterministic: every CTA adds the slots in the same order.
__device__ __forceinline__ void grid_allreduce_sums_only(const CgArgs &a, double *s_warp, double *s_tot, unsigned &nsum) {
    nsum++;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    ulonglong2 *slots = a.sum_slots + (size_t)(nsum & 1u) * a.n_sys * gridDim.x;
    __syncthreads();
    if (warp == 0) {
        for (int s = 0; s < a.n_sys; s++) {
            double v = lane < nwarps ? s_warp[s * MAXW + lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                const unsigned long long lo = ((unsigned long long)nsum << 32) | (bits & 0xffffffffull);
                const unsigned long long hi = ((unsigned long long)nsum << 32) | (bits >> 32);
                asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slots + (size_t)s * gridDim.x + blockIdx.x), "l"(lo), "l"(hi) : "memory");
            }
        }
        if (lane == 0) {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(a.sync_ctr) : "memory");
            const unsigned target = nsum * gridDim.x;
            unsigned seen;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.sync_ctr) : "memory");
            } while (seen < target);
        }
        __syncwarp();
        for (int s = 0; s < a.n_sys; s++) {
            const ulonglong2 *p = slots + (size_t)s * gridDim.x;
            double t = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                unsigned long long lo, hi;
                do {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p + i) : "memory");
                } while ((unsigned)(lo >> 32) != nsum || (unsigned)(hi >> 32) != nsum);
                t += __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
            }
            t = warp_sum(t);
            if (lane == 0) s_tot[s] = t;
        }
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

// ---- tensor memory as plain storage (jf_cg_pair_kernel, jf_cg_fused_stream_kernel, jf_cg_stream_kernel) -----------------------------------------------------
// TMEM (256 KB per SM, 128 lanes x 512 columns of 32 bits) is idle in a kernel without tensor-core work.  A CTA of four
// warps owns one lane per thread (warp w may touch lanes 32 (w % 4) ... + 31), so with 256 allocated columns every thread
// has 1 KB = 128 doubles of private storage that is read by tcgen05.ld straight into registers -- not through the
// LSU / shared-memory pipe the block rows' p gathers need.  The shared-memory tail of a row lives there instead.
__device__ __forceinline__ void tmem_st_double(unsigned taddr, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
#define TM_R4(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3])
__device__ __forceinline__ void tmem_ld_x2(unsigned taddr, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(unsigned taddr, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : TM_R4(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(unsigned taddr, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TM_R4(r, 0), TM_R4(r, 4), TM_R4(r, 8), TM_R4(r, 12)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(unsigned taddr, unsigned *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : TM_R4(r, 0), TM_R4(r, 4), TM_R4(r, 8), TM_R4(r, 12), TM_R4(r, 16), TM_R4(r, 20), TM_R4(r, 24), TM_R4(r, 28)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

