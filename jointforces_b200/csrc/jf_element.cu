// K1 -- per-tetrahedron semi-affine fiber energy, force and stiffness, FP64.
//
// What saenopy does with three einsums over (tet, 286 directions) per Newton iteration
// (_get_s_bar / _get_applied_epsilon / _update_energy / _update_f_glo / _update_K_glo, driven from
// /root/reference/jointforces/simulation.py:167) is done here per tet as
//
//   l_b^2 = s_b^T (F^T F) s_b                      6 FMA   (pair products s_j s_k from the table)
//   w, w', w'' (l_b - 1)                           closed forms of strain.py:336-338 and integrals
//   g_b, h_b                                       (A6), with 1/l from the rsqrt
//   G_jk   += g_b s_j s_k                          6 FMA   (symmetric 3x3)
//   H_jkac += h_b s_j s_k s_a s_c                  15 FMA  (fully symmetric 4th order: 15 unique)
//
// over the FOLDED direction table (antipodal pairs merged with weight 2; all terms are even in
// s), and only after the direction loop contracts with the deformation gradient and the shape
// tensor:  P = G F^T, f_m = Phi_m P;   K_mr = 1/2 (F (Phi_m Phi_r : H) F^T - (Phi_m G Phi_r) I).
// Identical in exact arithmetic to (A7)/(A8); ~4x fewer FP64 operations than forming the
// 4x4x3x3 tensor per direction.
//
// Mapping: JF_ELEM_LANES lanes per tet, each lane takes every JF_ELEM_LANES-th direction; the 22
// partial sums are combined with warp shuffles; the epilogue's ten unique (m<=r) blocks are then
// split over the same lanes.  The direction table (172 x 22 doubles = 30 KB) is staged once per
// persistent CTA in shared memory, component-major so that the lanes of a group read one
// contiguous 16*L-byte run per LDS.128.
#include <stdlib.h>

#include "jf_internal.h"

namespace {

struct ElemArgs {
    const int4 *tets;
    const double *phi;
    const double *vol;
    const double2 *dirtab;
    const jf_material *mat;
    const jf_sys_state *state;
    const double *U;
    double *Et;
    double *fel;
    double *Kel;
    long long T, N;
    int n_sys, n_dirs_pad;
    double n_beams;
};

__host__ __device__ constexpr int hid4(int a, int b, int c, int d) {
    int ex = (a == 0) + (b == 0) + (c == 0) + (d == 0);
    int ey = (a == 1) + (b == 1) + (c == 1) + (d == 1);
    return (4 - ex) * (5 - ex) / 2 + (4 - ex - ey);
}
__host__ __device__ constexpr int id3(int a, int b, int c) {
    int ex = (a == 0) + (b == 0) + (c == 0);
    int ey = (a == 1) + (b == 1) + (c == 1);
    return (3 - ex) * (4 - ex) / 2 + (3 - ex - ey);
}
__host__ __device__ constexpr int sym2(int a, int b) {  // 00 11 22 01 02 12
    return a == b ? a : (a + b + 2);
}

// per-material constants of (A1), formed once per tet
struct MatC {
    double k, ls, inv_d0, inv_ds, kd0, kd02, kds, kds2, kls, hkls2;
    __device__ __forceinline__ explicit MatC(const jf_material &m)
        : k(m.k), ls(m.ls), inv_d0(m.inv_d0), inv_ds(m.inv_ds), kd0(m.k * m.d0), kd02(m.k * m.d0 * m.d0), kds(m.k * m.ds),
          kds2(m.k * m.ds * m.ds), kls(m.k * m.ls), hkls2(0.5 * m.k * m.ls * m.ls) {}
};

// (A1): lambda -> w, w', w''.  Both exponential regimes (buckling: x = lam/d0, stiffening:
// x = (lam-ls)/ds) share one evaluation of em = expm1(x) and emx = expm1(x) - x: a single
// polynomial gives both for |x| < 1/8 (no cancellation when d0 or ds is 1e30, i.e.
// materials.linear, /root/reference/jointforces/materials.py:8; this is also the only path the
// far field ever takes), exp(x) - 1 otherwise.  Selects instead of branches, one data-dependent
// branch (the exp) that whole warps skip in the far field.
__device__ __forceinline__ void material_eval(const MatC &m, double lam, double &w0, double &w1, double &w2) {
    const bool neg = lam < 0.0, stf = lam >= m.ls;
    const bool lin = !(neg || stf);
    const double dl = lam - m.ls;
    double x = neg ? lam * m.inv_d0 : dl * m.inv_ds;
    if (lin) x = 0.0;
    double p = 1.0 / 39916800.0;  // x^2/2! ... x^11/11!: truncation below 4e-18 relative for |x| < 1/8
    p = fma(p, x, 1.0 / 3628800.0);
    p = fma(p, x, 1.0 / 362880.0);
    p = fma(p, x, 1.0 / 40320.0);
    p = fma(p, x, 1.0 / 5040.0);
    p = fma(p, x, 1.0 / 720.0);
    p = fma(p, x, 1.0 / 120.0);
    p = fma(p, x, 1.0 / 24.0);
    p = fma(p, x, 1.0 / 6.0);
    p = fma(p, x, 0.5);
    double emx = p * x * x;
    double em = x + emx;
    if (fabs(x) >= 0.125) {
        em = exp(x) - 1.0;
        emx = em - x;
    }
    w2 = m.k * (1.0 + em);
    const double w1e = (neg ? m.kd0 : m.kds) * em;
    const double w0e = (neg ? m.kd02 : m.kds2) * emx;
    w1 = stf ? m.kls + w1e : w1e;
    w0 = stf ? (m.hkls2 + m.kls * dl) + w0e : w0e;
    if (lin) {
        w1 = m.k * lam;
        w0 = 0.5 * m.k * lam * lam;
    }
}

// (A4) F = I + sum_m u_m (x) Phi_m
__device__ __forceinline__ void deformation_gradient(const double *Us, const int4 nd, const double *phi, double F[3][3]) {
    const int nn[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) F[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const double *u = Us + 3ll * nn[m];
        const double u0 = u[0], u1 = u[1], u2 = u[2];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double pmj = phi[3 * m + j];
            F[0][j] = fma(u0, pmj, F[0][j]);
            F[1][j] = fma(u1, pmj, F[1][j]);
            F[2][j] = fma(u2, pmj, F[2][j]);
        }
    }
}

template <int L, int UN = 2, int MB = 3>
__global__ void __launch_bounds__(JF_ELEM_BLOCK, MB) jf_element_kernel(ElemArgs a) {
    extern __shared__ double2 tab[];
    const int ndp = a.n_dirs_pad;
    for (int i = threadIdx.x; i < (JF_NTAB / 2) * ndp; i += blockDim.x) tab[i] = a.dirtab[i];
    __syncthreads();

    const int sub = threadIdx.x % L;
    const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1)));
    const long long ngroups = (long long)gridDim.x * (blockDim.x / L);
    const long long items = (long long)a.n_sys * a.T;

    for (long long item = (long long)blockIdx.x * (blockDim.x / L) + threadIdx.x / L; item < items; item += ngroups) {
        const int s = (int)(item / a.T);
        const long long t = item - (long long)s * a.T;
        if (!a.state[s].active) continue;
        const MatC mat(a.mat[s]);
        const int4 nd = a.tets[t];
        const double *phi = a.phi + 12 * t;
        const double *Us = a.U + (long long)s * a.N * 3;
        // C = F^T F from a deformation gradient that is NOT kept across the direction loop (it is rebuilt for the
        // epilogue from L1-resident data: 36 FMA against 18 registers held through 43 directions)
        double C[6];
        {
        double F[3][3];
        deformation_gradient(Us, nd, phi, F);
        // off-diagonals doubled so that l^2 = sum_c C[c] * (s_j s_k)[c]
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = j; k < 3; k++) {
                double v = F[0][j] * F[0][k] + F[1][j] * F[1][k] + F[2][j] * F[2][k];
                C[sym2(j, k)] = (j == k) ? v : 2.0 * v;
            }
        }
        const double vnb = a.vol[t] / a.n_beams;

        double e = 0.0, G[6], H[15];
#pragma unroll
        for (int i = 0; i < 6; i++) G[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 15; i++) H[i] = 0.0;

#pragma unroll UN
        for (int b = sub; b < ndp; b += L) {
            const double2 t0 = tab[0 * ndp + b], t1 = tab[1 * ndp + b], t2 = tab[2 * ndp + b];
            double l2 = C[0] * t0.x;
            l2 = fma(C[1], t0.y, l2);
            l2 = fma(C[2], t1.x, l2);
            l2 = fma(C[3], t1.y, l2);
            l2 = fma(C[4], t2.x, l2);
            l2 = fma(C[5], t2.y, l2);
            const double r = rsqrt(l2);
            double ell = l2 * r;
            ell = fma(0.5 * r, fma(-ell, ell, l2), ell);  // one Newton step: l = sqrt(l2) to ~0.5 ulp
            double w0, w1, w2;
            material_eval(mat, ell - 1.0, w0, w1, w2);
            const double2 t10 = tab[10 * ndp + b];
            const double aw = vnb * t10.y;                 // V/N_b times the fold weight
            const double g = -(aw * w1) * r;               // (A6)
            const double h = aw * (ell * w2 - w1) * (r * r * r);
            e = fma(aw, w0, e);
            G[0] = fma(g, t0.x, G[0]);
            G[1] = fma(g, t0.y, G[1]);
            G[2] = fma(g, t1.x, G[2]);
            G[3] = fma(g, t1.y, G[3]);
            G[4] = fma(g, t2.x, G[4]);
            G[5] = fma(g, t2.y, G[5]);
#pragma unroll
            for (int q = 0; q < 7; q++) {
                const double2 tm = tab[(3 + q) * ndp + b];
                H[2 * q] = fma(h, tm.x, H[2 * q]);
                H[2 * q + 1] = fma(h, tm.y, H[2 * q + 1]);
            }
            H[14] = fma(h, t10.x, H[14]);
        }
        // direction-sphere reduction across the lanes of the group
#pragma unroll
        for (int off = 1; off < L; off <<= 1) {
            e += __shfl_xor_sync(gmask, e, off);
#pragma unroll
            for (int i = 0; i < 6; i++) G[i] += __shfl_xor_sync(gmask, G[i], off);
#pragma unroll
            for (int i = 0; i < 15; i++) H[i] += __shfl_xor_sync(gmask, H[i], off);
        }

        const long long el = (long long)s * a.T + t;
        if (sub == 0) a.Et[el] = e;                        // (A5)
        asm volatile("" ::: "memory");                     // make the compiler rebuild F instead of keeping it alive
        double F[3][3];
        deformation_gradient(Us, nd, phi, F);

        // (A7) f_m = Phi_m . (G F^T)
        {
            double PG[3][3];  // PG[j][i] = sum_k G_jk F_ik
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++)
                    PG[j][i] = G[sym2(j, 0)] * F[i][0] + G[sym2(j, 1)] * F[i][1] + G[sym2(j, 2)] * F[i][2];
            for (int m = sub; m < 4; m += L) {
                const double p0 = phi[3 * m], p1 = phi[3 * m + 1], p2 = phi[3 * m + 2];
                double *out = a.fel + el * 12 + 3 * m;
#pragma unroll
                for (int i = 0; i < 3; i++) out[i] = p0 * PG[0][i] + p1 * PG[1][i] + p2 * PG[2][i];
            }
        }
        // (A8) ten unique blocks K_mr, m <= r
        for (int p = sub; p < 10; p += L) {
            // pair index -> (m, r): 00 01 02 03 11 12 13 22 23 33
            const int m = (p >= 4) + (p >= 7) + (p >= 9);
            const int r = p - (m == 0 ? 0 : (m == 1 ? 3 : (m == 2 ? 5 : 6)));
            const double pm[3] = {phi[3 * m], phi[3 * m + 1], phi[3 * m + 2]};
            const double pr[3] = {phi[3 * r], phi[3 * r + 1], phi[3 * r + 2]};
            double W[10];  // W_kac = sum_j Phi_mj H_jkac, fully symmetric in (k,a,c)
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int a2 = k; a2 < 3; a2++)
#pragma unroll
                    for (int c2 = a2; c2 < 3; c2++)
                        W[id3(k, a2, c2)] = pm[0] * H[hid4(0, k, a2, c2)] + pm[1] * H[hid4(1, k, a2, c2)] + pm[2] * H[hid4(2, k, a2, c2)];
            double M[6];  // M_ac = sum_k Phi_rk W_kac
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
                for (int c2 = a2; c2 < 3; c2++)
                    M[sym2(a2, c2)] = pr[0] * W[id3(0, a2, c2)] + pr[1] * W[id3(1, a2, c2)] + pr[2] * W[id3(2, a2, c2)];
            double gamma = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++)
                gamma += pm[j] * (G[sym2(j, 0)] * pr[0] + G[sym2(j, 1)] * pr[1] + G[sym2(j, 2)] * pr[2]);
            double X[3][3];  // X = F M
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++)
                    X[i][c2] = F[i][0] * M[sym2(0, c2)] + F[i][1] * M[sym2(1, c2)] + F[i][2] * M[sym2(2, c2)];
            double *out = a.Kel + el * 90 + p * 9;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int l = 0; l < 3; l++) {
                    double v = X[i][0] * F[l][0] + X[i][1] * F[l][1] + X[i][2] * F[l][2];
                    if (i == l) v -= gamma;
                    out[3 * i + l] = 0.5 * v;
                }
        }
    }
}

}  // namespace

template <int L, int UN = 2, int MB = 3>
static int elem_configure_t(jf_ctx *c, size_t smem) {
    JF_CUDA(cudaFuncSetAttribute(jf_element_kernel<L, UN, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jf_element_kernel<L, UN, MB>, JF_ELEM_BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    c->elem_grid = c->sm_count * per_sm;
    return 0;
}

int jf_elem_configure(jf_ctx *c) {
    size_t smem = (size_t)(JF_NTAB / 2) * c->n_dirs_pad * sizeof(double2);
    c->elem_lanes = JF_ELEM_LANES;
    if (const char *e = getenv("JF_ELEM_LANES")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) c->elem_lanes = v;
    }
    c->elem_variant = 0;
    if (const char *e = getenv("JF_ELEM_VARIANT")) c->elem_variant = atoi(e);  // tuning experiments (4 lanes only)
    if (c->elem_lanes == 4) switch (c->elem_variant) {
        case 1: return elem_configure_t<4, 3, 3>(c, smem);
        case 2: return elem_configure_t<4, 4, 3>(c, smem);
        case 3: return elem_configure_t<4, 2, 2>(c, smem);
        case 4: return elem_configure_t<4, 4, 2>(c, smem);
        case 5: return elem_configure_t<4, 2, 4>(c, smem);
        case 6: return elem_configure_t<4, 1, 4>(c, smem);
        default: c->elem_variant = 0;
    }
    switch (c->elem_lanes) {
        case 1: return elem_configure_t<1>(c, smem);
        case 2: return elem_configure_t<2>(c, smem);
        case 8: return elem_configure_t<8>(c, smem);
        default: return elem_configure_t<4>(c, smem);
    }
}

int jf_launch_element(jf_ctx *c) {
    ElemArgs a;
    a.tets = (const int4 *)c->d_tets;
    a.phi = c->d_phi;
    a.vol = c->d_vol;
    a.dirtab = (const double2 *)c->d_dirtab;
    a.mat = c->d_mat;
    a.state = c->d_state;
    a.U = c->d_U;
    a.Et = c->d_Et;
    a.fel = c->d_fel;
    a.Kel = c->d_Kel;
    a.T = c->T;
    a.N = c->N;
    a.n_sys = c->n_sys;
    a.n_dirs_pad = c->n_dirs_pad;
    a.n_beams = (double)c->n_beams;
    size_t smem = (size_t)(JF_NTAB / 2) * c->n_dirs_pad * sizeof(double2);
    long long groups_per_block = JF_ELEM_BLOCK / c->elem_lanes;
    long long need = ((long long)c->n_sys * c->T + groups_per_block - 1) / groups_per_block;
    int grid = (int)(need < c->elem_grid ? (need > 0 ? need : 1) : c->elem_grid);
    if (c->elem_lanes == 4 && c->elem_variant) switch (c->elem_variant) {
        case 1: jf_element_kernel<4, 3, 3><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 2: jf_element_kernel<4, 4, 3><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 3: jf_element_kernel<4, 2, 2><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 4: jf_element_kernel<4, 4, 2><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 5: jf_element_kernel<4, 2, 4><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        default: jf_element_kernel<4, 1, 4><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
    }
    else switch (c->elem_lanes) {
        case 1: jf_element_kernel<1><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 2: jf_element_kernel<2><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        case 8: jf_element_kernel<8><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
        default: jf_element_kernel<4><<<grid, JF_ELEM_BLOCK, smem, c->stream>>>(a); break;
    }
    JF_CUDA(cudaGetLastError());
    c->stats.kernel_launches++;
    return 0;
}
