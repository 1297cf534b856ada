// Dependent-issue latencies on this GPU, in SM cycles (one warp, one CTA unless noted):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency_bench latency_bench.cu && ./latency_bench
// Used to decide how the CG kernels' serial pieces are laid out (DESIGN.md "What bounds it").
#include <cstdio>
#include <cuda_runtime.h>

#define N 512

__global__ void k_lat(double *out, long long *cyc, double seed, unsigned long long *gmem) {
    __shared__ double sm[64];
    double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    int k = 0;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 4 independent DFMA chains (ILP 4), N each
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    a += a1 + a2 + a3;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a + c;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // FFMA chain
    float f = (float)a, fb = 1.0001f, fc = 1e-6f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, fb, fc);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    a += f;
    // double division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) a = 1.0 / (a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // u64 -> double -> u64 chain
    unsigned long long u = (unsigned long long)(a * 1e6) + 12345;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 64; i++) { double d = (double)u; u = (unsigned long long)__double_as_longlong(d) >> 7; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // double shuffle chain (two SHFL)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 64; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // shared-memory round trip chain (STS + LDS)
    t0 = clock64();
    for (int i = 0; i < 64; i++) { sm[threadIdx.x] = a; __syncwarp(); a = sm[threadIdx.x ^ 1] + c; __syncwarp(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // L2 load chain (ld.cg pointer chase on one line)
    unsigned long long v = 0;
    t0 = clock64();
    for (int i = 0; i < 64; i++) { asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(gmem + (v & 1)) : "memory"); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // atomic with return (round trip)
    t0 = clock64();
    for (int i = 0; i < 64; i++) { v = atomicAdd(gmem + 8 + (v & 1), 1ull); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // store + fence.acq_rel.gpu
    t0 = clock64();
    for (int i = 0; i < 64; i++) { gmem[32 + threadIdx.x] = v + i; asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // fence.acq_rel.gpu with nothing pending
    t0 = clock64();
    for (int i = 0; i < 64; i++) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // store + __threadfence (fence.sc)
    t0 = clock64();
    for (int i = 0; i < 64; i++) { gmem[32 + threadIdx.x] = v + i; __threadfence(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    // red + relaxed load of the same word (what one poll costs)
    t0 = clock64();
    for (int i = 0; i < 64; i++) {
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(gmem + 16), "l"(1ull) : "memory");
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(gmem + 16 + (v & 0)) : "memory");
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * (N / 64); k++;
    out[threadIdx.x] = a + (double)u + (double)v;
}

int main() {
    double *out; long long *cyc; unsigned long long *g;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 32 * 8); cudaMalloc(&g, 4096 * 8); cudaMemset(g, 0, 4096 * 8);
    const char *names[] = {"DFMA dependent", "DFMA 4 chains (per 4 ops)", "DADD dependent", "FFMA dependent", "1/(x+1.5) double division + add",
                           "u64->f64 convert + shift", "double shfl_xor + DADD", "STS+LDS round trip + DADD", "ld.global.cg (L2 hit) chase",
                           "atomicAdd u64 with return", "st + fence.acq_rel.gpu", "fence.acq_rel.gpu alone", "st + __threadfence (fence.sc)",
                           "red + ld.relaxed same word"};
    for (int warps = 1; warps <= 8; warps *= 8) {
        for (int rep = 0; rep < 2; rep++) k_lat<<<1, 32 * warps>>>(out, cyc, 1.0, g);
        cudaDeviceSynchronize();
        long long h[32];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("-- %d warp(s) in the CTA, cycles per step (thread 0)\n", warps);
        for (int i = 0; i < 14; i++) printf("%-36s %8.1f\n", names[i], (double)h[i] / N);
    }
    // the same with all SMs busy: 148 CTAs of 8 warps
    for (int rep = 0; rep < 2; rep++) k_lat<<<148, 256>>>(out, cyc, 1.0, g);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("-- 148 CTAs x 8 warps (last writer's thread 0)\n");
    for (int i = 0; i < 14; i++) printf("%-36s %8.1f\n", names[i], (double)h[i] / N);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
