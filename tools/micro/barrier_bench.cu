// Micro-benchmark: cost of one grid-wide barrier+all-reduce on B200, several implementations.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

__device__ __forceinline__ double warp_sum(double v) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ void ll_store(ulonglong2 *p, double v, unsigned flag) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    unsigned long long lo = ((unsigned long long)flag << 32) | (bits & 0xffffffffull), hi = ((unsigned long long)flag << 32) | (bits >> 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ double ll_wait(const ulonglong2 *p, unsigned flag) {
    unsigned long long lo, hi;
    do { asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    } while ((unsigned)(lo >> 32) != flag || (unsigned)(hi >> 32) != flag);
    return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
}
template <int FENCE>  // 0 none, 1 __threadfence (sc), 2 fence.acq_rel.gpu
__device__ __forceinline__ void fence() {
    if (FENCE == 1) __threadfence();
    if (FENCE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

template <int MODE, int FENCE>
__global__ void bench(ulonglong2 *slots, double *partial, double *out, int iters, double *junk) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_warp[32];
    __shared__ double s_tot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned epoch = 0;
    double acc = 1.0 + threadIdx.x * 1e-6, total = 0;
    for (int it = 0; it < iters; it++) {
        junk[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;  // a global write the barrier must publish
        double v = warp_sum(acc);
        if (lane == 0) s_warp[warp] = v;
        if (MODE == 0) {  // cooperative groups: publish, grid.sync, collect
            __syncthreads();
            if (warp == 0) { double t = lane < nwarps ? s_warp[lane] : 0; t = warp_sum(t); if (lane == 0) partial[(it & 1) * gridDim.x + blockIdx.x] = t; }
            grid.sync();
            if (warp == 0) { double t = 0; for (int i = lane; i < gridDim.x; i += 32) t += __ldcg(partial + (it & 1) * gridDim.x + i); t = warp_sum(t); if (lane == 0) s_tot = t; }
            __syncthreads();
        } else if (MODE == 4) {  // last arriver fans out per-CTA flags (no polling on the counter); full fence semantics
            epoch++;
            __syncthreads();
            if (warp == 0) {
                double t = lane < nwarps ? s_warp[lane] : 0; t = warp_sum(t);
                unsigned *ctr = (unsigned *)slots;                    // arrival counter
                unsigned *flags = (unsigned *)slots + 64;             // one flag per CTA, 32 words apart (own 128-B line)
                double *pp = partial + (epoch & 1) * gridDim.x;
                unsigned last = 0;
                if (lane == 0) {
                    pp[blockIdx.x] = t;
                    fence<FENCE>();
                    unsigned old = atomicAdd(ctr, 1u);
                    last = (old == epoch * gridDim.x - 1u);
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    for (int i = lane; i < gridDim.x; i += 32) asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(flags + 32 * i), "r"(epoch) : "memory");
                } else if (lane == 0) {
                    unsigned seen;
                    do { asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + 32 * blockIdx.x) : "memory"); } while (seen != epoch);
                }
                __syncwarp();
                fence<FENCE>();
                double u = 0;
                for (int i = lane; i < gridDim.x; i += 32) u += __ldcg(pp + i);
                u = warp_sum(u);
                if (lane == 0) s_tot = u;
            }
            __syncthreads();
        } else if (MODE == 3) {  // counter barrier that only carries the sums: no memory fences at all
            epoch++;
            __syncthreads();
            if (warp == 0) {
                double t = lane < nwarps ? s_warp[lane] : 0; t = warp_sum(t);
                unsigned *ctr = (unsigned *)slots;   // monotone arrival counter
                double *pp = partial + (epoch & 1) * gridDim.x;
                if (lane == 0) {
                    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(pp + blockIdx.x), "d"(t) : "memory");
                    if (FENCE == 1) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(ctr) : "memory");
                    else asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(ctr) : "memory");
                    unsigned seen;
                    const unsigned target = epoch * gridDim.x;
                    do { if (FENCE == 1) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                         else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
                }
                __syncwarp();
                double u = 0;
                for (int i = lane; i < gridDim.x; i += 32) { double x; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(x) : "l"(pp + i) : "memory"); u += x; }
                u = warp_sum(u);
                if (lane == 0) s_tot = u;
            }
            __syncthreads();
        } else {          // flag-carrying slots
            epoch++;
            ulonglong2 *sl = slots + (size_t)(epoch & 1u) * gridDim.x;
            __syncthreads();
            if (warp == 0) {
                double t = lane < nwarps ? s_warp[lane] : 0; t = warp_sum(t);
                if (lane == 0) { fence<FENCE>(); ll_store(sl + blockIdx.x, t, epoch); }
                double u = 0;
                if (MODE == 1) { for (int i = lane; i < gridDim.x; i += 32) u += ll_wait(sl + i, epoch); }
                if (MODE == 2) {  // all of a lane's slots polled together; retry only the missing ones
                    constexpr int K = 10;
                    double val[K]; unsigned okm = 0; const int n = (gridDim.x - lane + 31) / 32;
                    bool all;
                    do {
                        all = true;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            if (k < n && !(okm >> k & 1)) {
                                unsigned long long lo, hi;
                                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(sl + lane + 32 * k) : "memory");
                                if ((unsigned)(lo >> 32) == epoch && (unsigned)(hi >> 32) == epoch) { val[k] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull))); okm |= 1u << k; }
                                else all = false;
                            }
                        }
                    } while (!__all_sync(0xffffffffu, all));
#pragma unroll
                    for (int k = 0; k < K; k++) if (k < n) u += val[k];
                }
                u = warp_sum(u);
                if (lane == 0) s_tot = u;
                fence<FENCE>();
            }
            __syncthreads();
        }
        total += s_tot;
        acc = acc * 0.999 + 1e-3;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = total;
}

template <int MODE, int FENCE>
float run(int grid, int block, int iters) {
    ulonglong2 *slots; double *partial, *out, *junk;
    cudaMalloc(&slots, (2 * grid + 4096) * sizeof(ulonglong2)); cudaMemset(slots, 0, (2 * grid + 4096) * sizeof(ulonglong2));
    cudaMalloc(&partial, 2 * grid * sizeof(double)); cudaMalloc(&out, 8); cudaMalloc(&junk, (size_t)grid * block * 8);
    void *args[] = {&slots, &partial, &out, &iters, &junk};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemset(slots, 0, (2 * grid + 4096) * sizeof(ulonglong2));
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel((void *)bench<MODE, FENCE>, dim3(grid), dim3(block), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return -1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(slots); cudaFree(partial); cudaFree(out); cudaFree(junk);
    return best * 1e3f / iters;
}

int main() {
    const int iters = 20000;
    int grids[] = {134, 148}, blocks[] = {128, 256};
    printf("us per barrier+allreduce   grid block  coopgroups  slots+sc-fence  slots+acqrel  slots+nofence | parallel-poll: sc acqrel nofence\n");
    for (int g : grids) for (int b : blocks) {
        if (g * b > 148 * 2048) continue;
        printf("%5d %5d   %8.2f   %8.2f   %8.2f   %8.2f", g, b, run<0, 1>(g, b, iters), run<1, 1>(g, b, iters), run<1, 2>(g, b, iters), run<1, 0>(g, b, iters));
        printf(" | %8.2f   %8.2f   %8.2f", run<2, 1>(g, b, iters), run<2, 2>(g, b, iters), run<2, 0>(g, b, iters));
        printf(" | counter, sums only: rel/acq %8.2f  relaxed %8.2f", run<3, 1>(g, b, iters), run<3, 0>(g, b, iters));
        printf(" | fan-out flags: sc %8.2f acqrel %8.2f nofence %8.2f\n", run<4, 1>(g, b, iters), run<4, 2>(g, b, iters), run<4, 0>(g, b, iters));
    }
    return 0;
}
