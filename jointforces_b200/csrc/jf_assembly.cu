// K2 -- deterministic assembly.  saenopy rebuilds a scipy COO->CSR matrix from
// 144*T triplets every Newton iteration; here the sparsity pattern, and for every non-zero 3x3
// block the ordered list of element blocks that sum into it, are precomputed once
// (jf_setup.cu), so assembly is a pure gather: no atomics, bitwise reproducible.
//
// Stiffness layout (SELL-32 by block rows): slot q = slice_ptr[k] + 32*j + lane holds the j-th
// block of row 32*k+lane; its nine values live at val[(q>>5)*288 + 32*comp + (q&31)], so a warp
// that owns a slice reads every component as one 256-byte coalesced line.
#include <string.h>

#include <vector>

#include "jf_internal.h"

namespace {

// One CTA per group of 32 slots (= the 288 values of one SELL column of a slice), nine threads per slot: thread
// (slot, comp) adds component comp of the slot's element blocks in list order, so the nine lanes of a slot read one
// contiguous 72-byte block per contribution; the 288 results leave through shared memory as one contiguous line.
__global__ void __launch_bounds__(288) jf_assemble_K_kernel(const int32_t *__restrict__ cptr, const int32_t *__restrict__ csrc,
                                                            const double *__restrict__ Kel, double *__restrict__ val,
                                                            const jf_sys_state *__restrict__ state, long long n_slots,
                                                            long long T) {
    const int s = blockIdx.y;
    if (!state[s].active) return;
    __shared__ double stage[288];
    const int slot_local = threadIdx.x / 9, comp = threadIdx.x - 9 * slot_local;
    const int comp_t = (comp % 3) * 3 + comp / 3;  // block stored for (r,m): K_mr = K_rm^T
    const long long q = (long long)blockIdx.x * 32 + slot_local;
    const double *K = Kel + (long long)s * T * 90;
    double acc = 0.0;
    const int beg = cptr[q], end = cptr[q + 1];
    for (int c = beg; c < end; c++) {
        const int src = csrc[c];
        acc += K[(long long)(src >> 1) * 9 + ((src & 1) ? comp_t : comp)];
    }
    stage[32 * comp + slot_local] = acc;
    __syncthreads();
    val[(long long)s * n_slots * 9 + (long long)blockIdx.x * 288 + threadIdx.x] = stage[threadIdx.x];
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

// fixed-order reductions: strain energy, |f - f_t|^2 and |f|^2 over free nodes.  The energy counts only tetrahedra
// with at least one free node (free nodes are numbered first: some corner < Nf), as saenopy's `_update_energy` does
// (SURVEY Appendix A5); Et itself keeps every tet's energy for jf_get_element_quantities.  RED_CTAS CTAs per system take
// fixed strided shares and publish partial sums; the CTA that arrives last (self-resetting ticket) adds them in
// index order -- deterministic whatever the arrival order.
constexpr int RED_CTAS = 32;
constexpr int RED_THREADS = 256;

__global__ void __launch_bounds__(RED_THREADS) jf_reduce_kernel(const int4 *__restrict__ tets, const double *__restrict__ Et, const double *__restrict__ f,
                                                                const double *__restrict__ b, jf_sys_state *__restrict__ state,
                                                                double *__restrict__ partial, unsigned *__restrict__ ticket, long long N,
                                                                long long Nf, long long Nf_pad, long long T, jf_relax_log rl) {
    const int s = blockIdx.y;
    if (!state[s].active) return;
    __shared__ double sm[3][RED_THREADS];
    __shared__ bool last;
    double e = 0.0, r2 = 0.0, f2 = 0.0;
    const double *E = Et + (long long)s * T;
    const long long stride = (long long)RED_CTAS * RED_THREADS, first = (long long)blockIdx.x * RED_THREADS + threadIdx.x;
    for (long long t = first; t < T; t += stride) {
        const int4 c = tets[t];
        if (min(min(c.x, c.y), min(c.z, c.w)) < Nf) e += E[t];
    }
    const double *fs = f + (long long)s * N * 3;
    const double4 *bs = (const double4 *)(b + (long long)s * Nf_pad * 4);
    for (long long n = first; n < Nf; n += stride) {
        const double x = fs[3 * n], y = fs[3 * n + 1], z = fs[3 * n + 2];
        f2 += x * x + y * y + z * z;
        const double4 bv = bs[n];
        r2 += bv.x * bv.x + bv.y * bv.y + bv.z * bv.z;
    }
    sm[0][threadIdx.x] = e;
    sm[1][threadIdx.x] = r2;
    sm[2][threadIdx.x] = f2;
    __syncthreads();
    for (int w = RED_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + w];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + w];
            sm[2][threadIdx.x] += sm[2][threadIdx.x + w];
        }
        __syncthreads();
    }
    double *mine = partial + ((long long)s * RED_CTAS + blockIdx.x) * 3;
    if (threadIdx.x == 0) {
        mine[0] = sm[0][0];
        mine[1] = sm[1][0];
        mine[2] = sm[2][0];
        __threadfence();
        last = atomicInc(&ticket[s], RED_CTAS - 1) == RED_CTAS - 1;  // wraps to 0: ready for the next launch
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double acc = 0.0;
        if (threadIdx.x < 3) {
            const volatile double *p = partial + (long long)s * RED_CTAS * 3 + threadIdx.x;
            for (int k = 0; k < RED_CTAS; k++) acc += p[3 * k];
            if (threadIdx.x == 0) state[s].energy = acc;
            if (threadIdx.x == 1) state[s].residual2 = acc;
            if (threadIdx.x == 2) state[s].force2 = acc;
        }
        const double energy = __shfl_sync(0xffffffffu, acc, 0), resid2 = __shfl_sync(0xffffffffu, acc, 1),
                     force2 = __shfl_sync(0xffffffffu, acc, 2);
        // ---- the relaxation driver's bookkeeping, on the device: log the record row of this evaluation and
        // decide whether the system goes on -- saenopy's rule: after Newton step i > 6, stop when
        // std(last five energies) / sqrt(5) / mean(last five energies) < rel_conv_crit (numpy's population std; IEEE
        // operations in the host loop's order, no contraction)
        if (rl.enabled && threadIdx.x == 0) {
            const int row = state[s].n_rows;                    // 0: start state, i + 1: after Newton step i
            if (row < rl.cap) {
                double *r = rl.rows + ((long long)s * rl.cap + row) * 4;
                r[0] = row == 0 ? 0.0 : state[s].du;
                r[1] = energy;
                r[2] = row == 0 ? force2 : resid2;              // saenopy logs sum f^2 (not f - f_target) for the start row
                r[3] = row == 0 ? 0.0 : (double)state[s].cg_iters;
                int go_on = row < rl.max_iter;                  // `row` Newton steps done
                if (row - 1 > 6) {
                    double mean = 0.0, var = 0.0;
                    for (int j = 0; j < 5; j++) mean = __dadd_rn(mean, r[1 - 4 * j]);
                    mean = __ddiv_rn(mean, 5.0);
                    for (int j = 0; j < 5; j++) {
                        const double d = __dadd_rn(r[1 - 4 * j], -mean);
                        var = __dadd_rn(var, __dmul_rn(d, d));
                    }
                    const double estd = __ddiv_rn(__dsqrt_rn(__ddiv_rn(var, 5.0)), __dsqrt_rn(5.0));
                    if (__ddiv_rn(estd, mean) < rl.conv_crit) go_on = 0;
                }
                state[s].n_rows = row + 1;
                if (!go_on) state[s].active = 0;
                if (rl.host_rows) {                             // progress for the host (mapped memory): row first, then the count
                    double *h = rl.host_rows + ((long long)s * rl.cap + row) * 4;
                    h[0] = r[0]; h[1] = r[1]; h[2] = r[2]; h[3] = r[3];
                    __threadfence_system();
                    rl.host_flags[2 * s + 1] = go_on;
                    rl.host_flags[2 * s] = row + 1;
                }
            }
        }
    }
}

// y = K x for one system, one thread per block row (parity hook: jf_apply_stiffness; the CG kernels have their own SpMV)
__global__ void __launch_bounds__(128) jf_spmv_kernel(const int32_t *__restrict__ slice_ptr, const int32_t *__restrict__ sell_col,
                                                      const double *__restrict__ val, const double *__restrict__ x4,
                                                      double *__restrict__ y4, long long n_slices) {
    const long long sl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sl >= n_slices) return;
    const int lane = threadIdx.x & 31;
    const int q0 = slice_ptr[sl], w = (slice_ptr[sl + 1] - q0) >> 5;
    const double *v = val + (long long)(q0 >> 5) * 288 + lane;
    double y0 = 0, y1 = 0, y2 = 0;
    for (int j = 0; j < w; j++) {
        const double *xj = x4 + 4ll * sell_col[q0 + 32 * j + lane];
        const double *vj = v + 288 * j;
        y0 += vj[0] * xj[0] + vj[32] * xj[1] + vj[64] * xj[2];
        y1 += vj[96] * xj[0] + vj[128] * xj[1] + vj[160] * xj[2];
        y2 += vj[192] * xj[0] + vj[224] * xj[1] + vj[256] * xj[2];
    }
    double *yo = y4 + 4 * (sl * 32 + lane);
    yo[0] = y0;
    yo[1] = y1;
    yo[2] = y2;
    yo[3] = 0.0;
}

}  // namespace

// y (N,3) = K x (N,3) with the assembled stiffness of the last evaluation; rows/columns of fixed nodes are zero
int jf_spmv_host(jf_ctx *c, int sys, const double *x, double *y) {
    std::vector<double> xin((size_t)c->Nf_pad * 4, 0.0), yout((size_t)c->Nf_pad * 4, 0.0);
    for (int64_t i = 0; i < c->Nf; i++)
        for (int k = 0; k < 3; k++) xin[4 * i + k] = x[3 * (int64_t)c->node_new2old[i] + k];
    // reuse the CG work vectors p0 (in) and Ap (out) of this system
    double *dx = c->d_p0 + (size_t)sys * c->Nf_pad * 4, *dy = c->d_Ap + (size_t)sys * c->Nf_pad * 4;
    JF_CUDA(cudaMemcpyAsync(dx, xin.data(), xin.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->n_slices > 0) {
        jf_spmv_kernel<<<(unsigned)((c->n_slices + 3) / 4), 128, 0, c->stream>>>(c->d_slice_ptr, c->d_sell_col,
                                                                                   c->d_val + (size_t)sys * c->n_slots * 9, dx, dy, c->n_slices);
        JF_CUDA(cudaGetLastError());
        c->stats.kernel_launches++;
    }
    JF_CUDA(cudaMemcpyAsync(yout.data(), dy, yout.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    memset(y, 0, (size_t)c->N * 3 * sizeof(double));
    for (int64_t i = 0; i < c->Nf; i++)
        for (int k = 0; k < 3; k++) y[3 * (int64_t)c->node_new2old[i] + k] = yout[4 * i + k];
    return 0;
}

int jf_launch_assembly(jf_ctx *c) {
    dim3 gk((unsigned)(c->n_slots / 32), (unsigned)c->n_sys);  // n_slots is a multiple of 32 (SELL-32)
    if (c->n_slots > 0) {
        jf_assemble_K_kernel<<<gk, 288, 0, c->stream>>>(c->d_contrib_ptr, c->d_contrib_src, c->d_Kel, c->d_val, c->d_state,
                                                        c->n_slots, c->T);
        JF_CUDA(cudaGetLastError());
        c->stats.kernel_launches++;
    }
    dim3 gf((unsigned)((c->N + 255) / 256), (unsigned)c->n_sys);
    jf_gather_f_kernel<<<gf, 256, 0, c->stream>>>(c->d_fn_ptr, c->d_fn_src, c->d_fel, c->d_ft, c->d_f, c->d_b, c->d_state, c->N,
                                                  c->Nf, c->Nf_pad, c->T);
    JF_CUDA(cudaGetLastError());
    jf_reduce_kernel<<<dim3(RED_CTAS, (unsigned)c->n_sys), RED_THREADS, 0, c->stream>>>((const int4 *)c->d_tets, c->d_Et, c->d_f, c->d_b, c->d_state, c->d_red_partial,
                                                                                      c->d_red_ticket, c->N, c->Nf, c->Nf_pad, c->T, c->rlog);
    JF_CUDA(cudaGetLastError());
    c->stats.kernel_launches += 2;
    return 0;
}
