// Tensor memory as plain storage: cycles per tcgen05.ld (32x32b shapes) for 1..4 warps of a 128-thread CTA, one or two
// CTAs per SM, against LDS.64 of the same bytes.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b));
}

__global__ void __launch_bounds__(128, 2) k_tmem(long long *cyc, double *out, int active_warps, int reps) {
    __shared__ uint32_t s_base;
    __shared__ double sm[128 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"l"((uint64_t)__cvta_generic_to_shared(&s_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 256; c += 2) tmem_st2(base + c, (uint32_t)(c * 131 + threadIdx.x), (uint32_t)(c + 7 * lane));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    for (int i = threadIdx.x; i < 128 * 32; i += 128) sm[i] = i * 0.5;
    __syncthreads();
    uint32_t acc = 0;
    double dacc = 0.0;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp < active_warps) {
        t0 = clock64();
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int c = 0; c < 256; c += 32) {
                uint32_t v[32];
                tmem_ld32(base + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                for (int i = 0; i < 32; i++) acc += v[i];
            }
        }
        t1 = clock64();
        // the same 1 KB per thread with LDS.64 (conflict-free, component-major)
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
                double v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = sm[((c * 2 + i) & 31) * 128 + threadIdx.x];
#pragma unroll
                for (int i = 0; i < 16; i++) dacc += v[i];
            }
        }
        t2 = clock64();
    }
    __syncthreads();
    if (lane == 0 && warp < active_warps) {
        cyc[(blockIdx.x * 4 + warp) * 2] = t1 - t0;
        cyc[(blockIdx.x * 4 + warp) * 2 + 1] = t2 - t1;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc + dacc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(s_base));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cyc;
    double *out;
    cudaMallocManaged(&cyc, sizeof(long long) * 8 * 2 * sms);
    cudaMalloc(&out, sizeof(double) * 128 * 2 * sms);
    const int reps = 200;
    for (int per_sm = 1; per_sm <= 2; per_sm++)
        for (int aw = 1; aw <= 4; aw++) {
            k_tmem<<<sms * per_sm, 128>>>(cyc, out, aw, reps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            double a = 0, b = 0;
            int n = 0;
            for (int i = 0; i < sms * per_sm; i++)
                for (int w = 0; w < aw; w++) { a += cyc[(i * 4 + w) * 2]; b += cyc[(i * 4 + w) * 2 + 1]; n++; }
            // per load of 32 columns (128 B per thread, 4 KB per warp) / per 16 LDS.64 (the same bytes)
            printf("CTAs/SM %d, %d warp(s) per CTA loading: tcgen05.ld x32 %.1f cycles per 4 KB per warp  -> %.1f B/clk/SM ; LDS.64 x16 %.1f cycles -> %.1f B/clk/SM\n",
                   per_sm, aw, a / n / (reps * 8), 4096.0 * aw * per_sm / (a / n / (reps * 8)), b / n / (reps * 8), 4096.0 * aw * per_sm / (b / n / (reps * 8)));
        }
    return 0;
}
