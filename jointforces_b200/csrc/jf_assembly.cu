// K2 -- deterministic assembly.  saenopy rebuilds a scipy COO->CSR matrix from
// 144*T triplets every Newton iteration; here the sparsity pattern, and for every non-zero 3x3
// block the ordered list of element blocks that sum into it, are precomputed once
// (jf_setup.cu), so assembly is a pure gather: no atomics, bitwise reproducible.
//
// Stiffness layout (SELL-32 by block rows): slot q = slice_ptr[k] + 32*j + lane holds the j-th
// block of row 32*k+lane; its nine values live at val[(q>>5)*288 + 32*comp + (q&31)], so a warp
// that owns a slice reads every component as one 256-byte coalesced line.
#include "jf_internal.h"

namespace {

__global__ void __launch_bounds__(256) jf_assemble_K_kernel(const int32_t *__restrict__ cptr, const int32_t *__restrict__ csrc,
                                                            const double *__restrict__ Kel, double *__restrict__ val,
                                                            const jf_sys_state *__restrict__ state, long long n_slots,
                                                            long long T) {
    const int s = blockIdx.y;
    if (!state[s].active) return;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slots) return;
    const double *K = Kel + (long long)s * T * 90;
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; i++) acc[i] = 0.0;
    const int beg = cptr[q], end = cptr[q + 1];
    for (int c = beg; c < end; c++) {
        const int src = csrc[c];
        const double *kb = K + (long long)(src >> 1) * 9;
        double v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = kb[i];
        if (src & 1) {  // block stored for (r,m): K_mr = K_rm^T
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int l = 0; l < 3; l++) acc[3 * i + l] += v[3 * l + i];
        } else {
#pragma unroll
            for (int i = 0; i < 9; i++) acc[i] += v[i];
        }
    }
    double *out = val + (long long)s * n_slots * 9 + (q >> 5) * 288 + (q & 31);
#pragma unroll
    for (int i = 0; i < 9; i++) out[32 * i] = acc[i];
}

// nodal forces f = gather of f_el (A10), and the CG right-hand side b = f - f_target on free nodes
__global__ void __launch_bounds__(256) jf_gather_f_kernel(const int32_t *__restrict__ fptr, const int32_t *__restrict__ fsrc,
                                                          const double *__restrict__ fel, const double *__restrict__ ft,
                                                          double *__restrict__ f, double *__restrict__ b,
                                                          const jf_sys_state *__restrict__ state, long long N, long long Nf,
                                                          long long Nf_pad, long long T) {
    const int s = blockIdx.y;
    if (!state[s].active) return;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double *fe = fel + (long long)s * T * 12;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int c = fptr[n]; c < fptr[n + 1]; c++) {
        const double *p = fe + 3ll * fsrc[c];
        a0 += p[0];
        a1 += p[1];
        a2 += p[2];
    }
    double *fo = f + ((long long)s * N + n) * 3;
    fo[0] = a0;
    fo[1] = a1;
    fo[2] = a2;
    if (n < Nf) {
        const double *t = ft + ((long long)s * N + n) * 3;
        double4 *bo = (double4 *)(b + ((long long)s * Nf_pad + n) * 4);
        *bo = make_double4(a0 - t[0], a1 - t[1], a2 - t[2], 0.0);
    }
}

// fixed-order reductions: strain energy, |f - f_t|^2 and |f|^2 over free nodes.  One CTA per system.
__global__ void __launch_bounds__(1024) jf_reduce_kernel(const double *__restrict__ Et, const double *__restrict__ f,
                                                         const double *__restrict__ b, jf_sys_state *__restrict__ state,
                                                         long long N, long long Nf, long long Nf_pad, long long T) {
    const int s = blockIdx.x;
    if (!state[s].active) return;
    __shared__ double sm[3][1024];
    double e = 0.0, r2 = 0.0, f2 = 0.0;
    const double *E = Et + (long long)s * T;
    for (long long t = threadIdx.x; t < T; t += 1024) e += E[t];
    const double *fs = f + (long long)s * N * 3;
    const double *bs = b + (long long)s * Nf_pad * 4;
    for (long long n = threadIdx.x; n < Nf; n += 1024) {
        const double x = fs[3 * n], y = fs[3 * n + 1], z = fs[3 * n + 2];
        f2 += x * x + y * y + z * z;
        const double bx = bs[4 * n], by = bs[4 * n + 1], bz = bs[4 * n + 2];
        r2 += bx * bx + by * by + bz * bz;
    }
    sm[0][threadIdx.x] = e;
    sm[1][threadIdx.x] = r2;
    sm[2][threadIdx.x] = f2;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + w];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + w];
            sm[2][threadIdx.x] += sm[2][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        state[s].energy = sm[0][0];
        state[s].residual2 = sm[1][0];
        state[s].force2 = sm[2][0];
    }
}

}  // namespace

int jf_launch_assembly(jf_ctx *c) {
    dim3 gk((unsigned)((c->n_slots + 255) / 256), (unsigned)c->n_sys);
    if (c->n_slots > 0) {
        jf_assemble_K_kernel<<<gk, 256, 0, c->stream>>>(c->d_contrib_ptr, c->d_contrib_src, c->d_Kel, c->d_val, c->d_state,
                                                        c->n_slots, c->T);
        JF_CUDA(cudaGetLastError());
        c->stats.kernel_launches++;
    }
    dim3 gf((unsigned)((c->N + 255) / 256), (unsigned)c->n_sys);
    jf_gather_f_kernel<<<gf, 256, 0, c->stream>>>(c->d_fn_ptr, c->d_fn_src, c->d_fel, c->d_ft, c->d_f, c->d_b, c->d_state, c->N,
                                                  c->Nf, c->Nf_pad, c->T);
    JF_CUDA(cudaGetLastError());
    jf_reduce_kernel<<<c->n_sys, 1024, 0, c->stream>>>(c->d_Et, c->d_f, c->d_b, c->d_state, c->N, c->Nf, c->Nf_pad, c->T);
    JF_CUDA(cudaGetLastError());
    c->stats.kernel_launches += 2;
    return 0;
}
