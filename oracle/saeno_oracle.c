/*
 * Plain-C CPU restatement of the saenopy relaxation path jointforces drives.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (jointforces_b200) never
 * touches it.
 *
 * PARITY UNPINNED: the algorithm lives in the un-vendored dependency `saenopy`
 * (/root/reference/setup.py:15, ">=0.7.4"); the reference holds no tests or golden vectors for
 * it.  This file restates SURVEY.md Appendix A (A1)-(A14) -- the same statement as
 * oracle/saeno_oracle.py, which is the readable version and which tests/ cross-check this file
 * against -- anchored on the reference's call sites:
 *   jointforces/simulation.py:100-104,143,146  (Solver setup, NaN-coded boundary conditions)
 *   jointforces/simulation.py:167              (solve_boundarycondition(step, max_iter, conv_crit))
 *   jointforces/strain.py:336-338              (the three-regime w''(lambda))
 *
 * Per tetrahedron and per direction b it accumulates, exactly in the shape saenopy's einsums
 * have (no folding of antipodal directions, no symmetric-tensor shortcut -- those belong to the
 * product kernels and are what this file checks):
 *   f[m][i]       += (Phi_m . s_b) * sbar_i * g_b                                   (A7)
 *   K[m][r][i][l] += 1/2 (Phi_m . s_b)(Phi_r . s_b)(h_b sbar_i sbar_l - g_b d_il)   (A8)
 * then assembles rows of free nodes into CSR, runs Hestenes-Stiefel CG (|r|^2 < tol |b|^2) and
 * the Newton/relaxation driver with the five-energy stop rule.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double k, d0, ls, ds; } so_material;

/* expm1(x) - x without cancellation (d0 = 1e30 for the `linear` material, materials.py:8) */
static double expm1_minus_x(double x) {
    if (fabs(x) < 0.125) {
        static const double c[] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                   1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0,
                                   1.0 / 24.0, 1.0 / 6.0, 0.5};
        double acc = c[0];
        for (int i = 1; i < 12; i++) acc = acc * x + c[i];
        return acc * x * x;
    }
    return expm1(x) - x;
}

/* (A1) lambda -> w, w', w'' */
static void material_eval(const so_material *m, double lam, double *w0, double *w1, double *w2) {
    if (lam < 0) {
        double x = lam / m->d0, em = expm1(x);
        *w2 = m->k * (1.0 + em);
        *w1 = m->k * m->d0 * em;
        *w0 = m->k * m->d0 * m->d0 * expm1_minus_x(x);
    } else if (lam >= m->ls) {
        double y = (lam - m->ls) / m->ds, em = expm1(y);
        *w2 = m->k * (1.0 + em);
        *w1 = m->k * m->ls + m->k * m->ds * em;
        *w0 = 0.5 * m->k * m->ls * m->ls + m->k * m->ls * (lam - m->ls) + m->k * m->ds * m->ds * expm1_minus_x(y);
    } else {
        *w2 = m->k;
        *w1 = m->k * lam;
        *w0 = 0.5 * m->k * lam * lam;
    }
}

void so_material_eval(double k, double d0, double ls, double ds, int n, const double *lam, double *w0,
                      double *w1, double *w2) {
    so_material m = {k, d0, ls, ds};
    for (int i = 0; i < n; i++) material_eval(&m, lam[i], w0 + i, w1 + i, w2 + i);
}

/* (A3) Phi = Chi * B^-1 (4x3 per tet), V = |det B| / 6 */
void so_shape_tensors(int T, const double *R, const int32_t *tets, double *Phi, double *V) {
    for (int t = 0; t < T; t++) {
        const int32_t *n = tets + 4 * t;
        double B[3][3];
        for (int c = 0; c < 3; c++)
            for (int i = 0; i < 3; i++) B[i][c] = R[3 * n[c + 1] + i] - R[3 * n[0] + i];
        double det = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) +
                     B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
        double inv[3][3];
        inv[0][0] = (B[1][1] * B[2][2] - B[1][2] * B[2][1]) / det;
        inv[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) / det;
        inv[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) / det;
        inv[1][0] = (B[1][2] * B[2][0] - B[1][0] * B[2][2]) / det;
        inv[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) / det;
        inv[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) / det;
        inv[2][0] = (B[1][0] * B[2][1] - B[1][1] * B[2][0]) / det;
        inv[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) / det;
        inv[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) / det;
        V[t] = fabs(det) / 6.0;
        double *P = Phi + 12 * t;
        for (int j = 0; j < 3; j++) {
            P[0 * 3 + j] = -(inv[0][j] + inv[1][j] + inv[2][j]);
            P[1 * 3 + j] = inv[0][j];
            P[2 * 3 + j] = inv[1][j];
            P[3 * 3 + j] = inv[2][j];
        }
    }
}

/* (A4)-(A8) for tets [0,T): E (T), f_el (T,4,3), K_el (T,4,4,3,3) */
void so_element_eval(int T, const int32_t *tets, const double *U, const double *Phi, const double *V, int Nb,
                     const double *s, double k, double d0, double ls, double ds, double *E, double *f_el,
                     double *K_el) {
    so_material mat = {k, d0, ls, ds};
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++) {
        const int32_t *n = tets + 4 * t;
        const double *P = Phi + 12 * t;
        double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int m = 0; m < 4; m++)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) F[i][j] += U[3 * n[m] + i] * P[3 * m + j];
        double f[4][3], K[4][4][3][3], e = 0.0;
        memset(f, 0, sizeof f);
        memset(K, 0, sizeof K);
        double vnb = V[t] / Nb;
        for (int b = 0; b < Nb; b++) {
            const double *sb = s + 3 * b;
            double sbar[3], sstar[4];
            for (int i = 0; i < 3; i++) sbar[i] = F[i][0] * sb[0] + F[i][1] * sb[1] + F[i][2] * sb[2];
            for (int m = 0; m < 4; m++) sstar[m] = P[3 * m] * sb[0] + P[3 * m + 1] * sb[1] + P[3 * m + 2] * sb[2];
            double ell = sqrt(sbar[0] * sbar[0] + sbar[1] * sbar[1] + sbar[2] * sbar[2]);
            double w0, w1, w2;
            material_eval(&mat, ell - 1.0, &w0, &w1, &w2);
            e += w0;
            double g = -(w1 / ell) * vnb;                      /* (A6) */
            double h = (ell * w2 - w1) / (ell * ell * ell) * vnb;
            double q[3][3];
            for (int i = 0; i < 3; i++)
                for (int l = 0; l < 3; l++) q[i][l] = 0.5 * (h * sbar[i] * sbar[l] - (i == l ? g : 0.0));
            for (int m = 0; m < 4; m++) {
                for (int i = 0; i < 3; i++) f[m][i] += sstar[m] * sbar[i] * g;
                for (int r = 0; r < 4; r++) {
                    double ss = sstar[m] * sstar[r];
                    for (int i = 0; i < 3; i++)
                        for (int l = 0; l < 3; l++) K[m][r][i][l] += ss * q[i][l];
                }
            }
        }
        E[t] = e / Nb * V[t];                                  /* (A5) */
        memcpy(f_el + 12 * (size_t)t, f, sizeof f);
        memcpy(K_el + 144 * (size_t)t, K, sizeof K);
    }
}

/* ---------------------------------------------------------------------------------------------
 * CSR pattern of the rows of free nodes (A10): scalar CSR of size 3N x 3N
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int N, T;
    int64_t nnz;
    int64_t *rowptr; /* 3N+1 */
    int32_t *col;    /* nnz   */
    int64_t *slot;   /* T*144 : CSR position of K_el entry, or -1 (row of a fixed node) */
    uint8_t *count_energy; /* T : 1 if the tet has at least one free node (its energy counts, see update_f_and_K) */
} so_pattern;

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

static so_pattern *pattern_build(int N, int T, const int32_t *tets, const uint8_t *movable) {
    so_pattern *p = (so_pattern *)calloc(1, sizeof *p);
    p->N = N;
    p->T = T;
    /* node -> neighbour nodes (with duplicates), via counting */
    int64_t *cnt = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    for (int t = 0; t < T; t++)
        for (int m = 0; m < 4; m++) cnt[tets[4 * t + m] + 1] += 4;
    for (int i = 0; i < N; i++) cnt[i + 1] += cnt[i];
    int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)cnt[N]);
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    memcpy(fill, cnt, sizeof(int64_t) * (size_t)N);
    for (int t = 0; t < T; t++)
        for (int m = 0; m < 4; m++)
            for (int r = 0; r < 4; r++) nb[fill[tets[4 * t + m]]++] = tets[4 * t + r];
    /* sort + unique per node */
    int64_t *nnb = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t)); /* block row ptr */
    for (int i = 0; i < N; i++) {
        int64_t a = cnt[i], b = cnt[i + 1];
        qsort(nb + a, (size_t)(b - a), sizeof(int32_t), cmp_i32);
        int64_t u = a;
        for (int64_t j = a; j < b; j++)
            if (j == a || nb[j] != nb[j - 1]) nb[u++] = nb[j];
        nnb[i + 1] = movable[i] ? (u - a) : 0;
        fill[i] = u - a; /* unique neighbour count */
    }
    for (int i = 0; i < N; i++) nnb[i + 1] += nnb[i];
    p->nnz = 9 * nnb[N];
    p->rowptr = (int64_t *)malloc(sizeof(int64_t) * (3 * (size_t)N + 1));
    p->col = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->nnz);
    p->rowptr[0] = 0;
    for (int i = 0; i < N; i++) {
        int64_t nn = nnb[i + 1] - nnb[i];
        for (int c = 0; c < 3; c++) {
            int64_t base = 9 * nnb[i] + c * 3 * nn;
            p->rowptr[3 * i + c + 1] = base + 3 * nn;
            for (int64_t j = 0; j < nn; j++)
                for (int l = 0; l < 3; l++) p->col[base + 3 * j + l] = 3 * nb[cnt[i] + j] + l;
        }
    }
    p->slot = (int64_t *)malloc(sizeof(int64_t) * 144 * (size_t)T);
    p->count_energy = (uint8_t *)malloc((size_t)T);
    for (int t = 0; t < T; t++)
        p->count_energy[t] = movable[tets[4 * t]] | movable[tets[4 * t + 1]] | movable[tets[4 * t + 2]] | movable[tets[4 * t + 3]];
    for (int t = 0; t < T; t++)
        for (int m = 0; m < 4; m++) {
            int32_t rn = tets[4 * t + m];
            int64_t nn = nnb[rn + 1] - nnb[rn];
            for (int r = 0; r < 4; r++) {
                int32_t cn = tets[4 * t + r];
                int64_t j = -1;
                if (movable[rn]) {
                    int32_t *hit = (int32_t *)bsearch(&cn, nb + cnt[rn], (size_t)fill[rn], sizeof(int32_t), cmp_i32);
                    j = hit - (nb + cnt[rn]);
                }
                for (int i = 0; i < 3; i++)
                    for (int l = 0; l < 3; l++)
                        p->slot[144 * (size_t)t + ((m * 4 + r) * 3 + i) * 3 + l] =
                            movable[rn] ? (9 * nnb[rn] + i * 3 * nn + 3 * j + l) : -1;
            }
        }
    free(cnt);
    free(nb);
    free(fill);
    free(nnb);
    return p;
}

static void pattern_free(so_pattern *p) {
    free(p->rowptr);
    free(p->col);
    free(p->slot);
    free(p->count_energy);
    free(p);
}

static void spmv(const so_pattern *p, const double *val, const double *x, double *y) {
    int n3 = 3 * p->N;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n3; i++) {
        double acc = 0.0;
        for (int64_t j = p->rowptr[i]; j < p->rowptr[i + 1]; j++) acc += val[j] * x[p->col[j]];
        y[i] = acc;
    }
}

static double dot(int n, const double *a, const double *b) {
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (int i = 0; i < n; i++) acc += a[i] * b[i];
    return acc;
}

/* (A13) returns iterations taken; x must hold n doubles */
static int cg_solve(const so_pattern *p, const double *val, const double *b, double *x, int maxiter, double tol,
                    double *r, double *pp, double *Ap) {
    int n = 3 * p->N;
    memset(x, 0, sizeof(double) * (size_t)n);
    double normb = dot(n, b, b);
    if (normb == 0) return 0;
    memcpy(r, b, sizeof(double) * (size_t)n);
    memcpy(pp, b, sizeof(double) * (size_t)n);
    double resid = dot(n, pp, pp);
    int it;
    for (it = 1; it <= maxiter; it++) {
        spmv(p, val, pp, Ap);
        double alpha = resid / dot(n, pp, Ap);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            x[i] += alpha * pp[i];
            r[i] -= alpha * Ap[i];
        }
        double rsnew = dot(n, r, r);
        if (rsnew < tol * normb) break;
        double beta = rsnew / resid;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) pp[i] = r[i] + beta * pp[i];
        resid = rsnew;
    }
    return it > maxiter ? maxiter : it;
}

/* evaluate + assemble: forces (N,3), CSR values, strain energy.
 * The global energy counts only tetrahedra with at least one free node (saenopy `_update_energy`, as SAENO's
 * `countEnergy`: a tet whose four nodes all carry a displacement condition cannot change and is left out of
 * E_glo).  With the reference's boundary conditions (outer shell held at zero, simulation.py:127-129) such tets
 * have zero energy anyway; the rule matters for non-zero prescribed displacements. */
static double update_f_and_K(const so_pattern *p, const int32_t *tets, const double *U, const double *Phi,
                             const double *V, int Nb, const double *s, const so_material *m, double *E, double *f_el,
                             double *K_el, double *forces, double *val) {
    int T = p->T;
    so_element_eval(T, tets, U, Phi, V, Nb, s, m->k, m->d0, m->ls, m->ds, E, f_el, K_el);
    memset(forces, 0, sizeof(double) * 3 * (size_t)p->N);
    memset(val, 0, sizeof(double) * (size_t)p->nnz);
    double e = 0.0;
    for (int t = 0; t < T; t++) {
        if (p->count_energy[t]) e += E[t];
        for (int mm = 0; mm < 4; mm++)
            for (int i = 0; i < 3; i++) forces[3 * tets[4 * t + mm] + i] += f_el[12 * (size_t)t + 3 * mm + i];
        const int64_t *sl = p->slot + 144 * (size_t)t;
        const double *Ke = K_el + 144 * (size_t)t;
        for (int q = 0; q < 144; q++)
            if (sl[q] >= 0) val[sl[q]] += Ke[q];
    }
    return e;
}

/* (A12)+(A14).  U in/out (N,3).  relrec: (max_iter+1, 3).  cg_iters: (max_iter).  Returns the
 * number of Newton iterations done.  stop_after_cg_total > 0 bounds the work for timing samples. */
int so_relax(int N, int T, const double *nodes, const int32_t *tets, const uint8_t *movable, const double *f_target,
             double *U, int Nb, const double *beams, double k, double d0, double ls, double ds, double step,
             int max_iter, double conv_crit, double *relrec, int32_t *cg_iters, double *forces_out,
             double *timing /* [0]=element+assembly s, [1]=cg s, may be NULL */, int nthreads,
             double cg_tol /* <= 0: saenopy's 1e-5 (on squared norms) */) {
    if (!(cg_tol > 0)) cg_tol = 0.00001;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    so_material mat = {k, d0, ls, ds};
    so_pattern *p = pattern_build(N, T, tets, movable);
    double *Phi = (double *)malloc(sizeof(double) * 12 * (size_t)T);
    double *V = (double *)malloc(sizeof(double) * (size_t)T);
    double *E = (double *)malloc(sizeof(double) * (size_t)T);
    double *f_el = (double *)malloc(sizeof(double) * 12 * (size_t)T);
    double *K_el = (double *)malloc(sizeof(double) * 144 * (size_t)T);
    double *val = (double *)malloc(sizeof(double) * (size_t)p->nnz);
    int n3 = 3 * N;
    double *forces = (double *)malloc(sizeof(double) * (size_t)n3);
    double *ff = (double *)malloc(sizeof(double) * (size_t)n3);
    double *uu = (double *)malloc(sizeof(double) * (size_t)n3);
    double *w1 = (double *)malloc(sizeof(double) * (size_t)n3);
    double *w2 = (double *)malloc(sizeof(double) * (size_t)n3);
    double *w3 = (double *)malloc(sizeof(double) * (size_t)n3);
    so_shape_tensors(T, nodes, tets, Phi, V);
    double t_el = 0, t_cg = 0, t0;
#ifdef _OPENMP
#define NOW() omp_get_wtime()
#else
#define NOW() 0.0
#endif
    t0 = NOW();
    double e = update_f_and_K(p, tets, U, Phi, V, Nb, beams, &mat, E, f_el, K_el, forces, val);
    t_el += NOW() - t0;
    double r0 = 0.0;
    for (int i = 0; i < N; i++)
        if (movable[i])
            for (int c = 0; c < 3; c++) r0 += forces[3 * i + c] * forces[3 * i + c];
    relrec[0] = 0.0;
    relrec[1] = e;
    relrec[2] = r0;
    int it;
    for (it = 0; it < max_iter; it++) {
        for (int i = 0; i < N; i++)
            for (int c = 0; c < 3; c++) ff[3 * i + c] = movable[i] ? forces[3 * i + c] - f_target[3 * i + c] : 0.0;
        t0 = NOW();
        cg_iters[it] = cg_solve(p, val, ff, uu, 3 * N, cg_tol, w1, w2, w3);
        t_cg += NOW() - t0;
        double du = 0.0;
        for (int i = 0; i < N; i++)
            if (movable[i])
                for (int c = 0; c < 3; c++) {
                    U[3 * i + c] += uu[3 * i + c] * step;
                    du += uu[3 * i + c] * uu[3 * i + c];
                }
        du *= step * step;
        t0 = NOW();
        e = update_f_and_K(p, tets, U, Phi, V, Nb, beams, &mat, E, f_el, K_el, forces, val);
        t_el += NOW() - t0;
        double res = 0.0;
        for (int i = 0; i < N; i++)
            if (movable[i])
                for (int c = 0; c < 3; c++) {
                    double d = forces[3 * i + c] - f_target[3 * i + c];
                    res += d * d;
                }
        double *row = relrec + 3 * (it + 1);
        row[0] = du;
        row[1] = e;
        row[2] = res;
        if (it > 6) {
            double mean = 0, var = 0;
            for (int j = 0; j < 5; j++) mean += relrec[3 * (it + 1 - j) + 1];
            mean /= 5.0;
            for (int j = 0; j < 5; j++) {
                double d = relrec[3 * (it + 1 - j) + 1] - mean;
                var += d * d;
            }
            double estd = sqrt(var / 5.0) / sqrt(5.0);
            if (estd / mean < conv_crit) {
                it++;
                break;
            }
        }
    }
    if (forces_out) memcpy(forces_out, forces, sizeof(double) * (size_t)n3);
    if (timing) {
        timing[0] = t_el;
        timing[1] = t_cg;
    }
    free(Phi); free(V); free(E); free(f_el); free(K_el); free(val);
    free(forces); free(ff); free(uu); free(w1); free(w2); free(w3);
    pattern_free(p);
    return it;
}

/* one evaluation + assembly, for checking the product's assembled quantities:
 * forces (N,3); dense-free CSR (rowptr 3N+1, col, val) sized by so_pattern_nnz */
int64_t so_pattern_nnz(int N, int T, const int32_t *tets, const uint8_t *movable) {
    so_pattern *p = pattern_build(N, T, tets, movable);
    int64_t nnz = p->nnz;
    pattern_free(p);
    return nnz;
}

double so_assemble(int N, int T, const double *nodes, const int32_t *tets, const uint8_t *movable, const double *U,
                   int Nb, const double *beams, double k, double d0, double ls, double ds, double *forces,
                   int64_t *rowptr, int32_t *col, double *val) {
    so_material mat = {k, d0, ls, ds};
    so_pattern *p = pattern_build(N, T, tets, movable);
    double *Phi = (double *)malloc(sizeof(double) * 12 * (size_t)T);
    double *V = (double *)malloc(sizeof(double) * (size_t)T);
    double *E = (double *)malloc(sizeof(double) * (size_t)T);
    double *f_el = (double *)malloc(sizeof(double) * 12 * (size_t)T);
    double *K_el = (double *)malloc(sizeof(double) * 144 * (size_t)T);
    so_shape_tensors(T, nodes, tets, Phi, V);
    double e = update_f_and_K(p, tets, U, Phi, V, Nb, beams, &mat, E, f_el, K_el, forces, val);
    memcpy(rowptr, p->rowptr, sizeof(int64_t) * (3 * (size_t)N + 1));
    memcpy(col, p->col, sizeof(int32_t) * (size_t)p->nnz);
    free(Phi); free(V); free(E); free(f_el); free(K_el);
    pattern_free(p);
    return e;
}

/* ---------------------------------------------------------------------------------------------
 * Diagnostic (tools/cg_precision_study.py): how many iterations does the SAME CG (A13) take on the SAME system
 * when only the floating-point evaluation of its sums changes?  variant 0: this file's (plain double, products and
 * sums rounded separately, sequential sums); 1: fused multiply-adds and pairwise (tree) dot products, the way a GPU
 * evaluates them; 2: everything in long double (64-bit mantissa).  CG's convergence is delayed by rounding errors
 * (loss of orthogonality); the more exact the arithmetic, the fewer iterations to the same residual.
 * ------------------------------------------------------------------------------------------- */
static double dot_pairwise(const double *a, const double *b, int n) {
    if (n <= 32) {
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc = fma(a[i], b[i], acc);
        return acc;
    }
    int h = n / 2;
    return dot_pairwise(a, b, h) + dot_pairwise(a + h, b + h, n - h);
}

static int cg_variant(const so_pattern *p, const double *val, const double *b, int maxiter, double tol, int variant) {
    int n = 3 * p->N;
    if (variant == 2) {
        long double *x = calloc(n, sizeof(long double)), *r = malloc(n * sizeof(long double)), *pp = malloc(n * sizeof(long double)),
                    *Ap = malloc(n * sizeof(long double));
        long double normb = 0;
        for (int i = 0; i < n; i++) { r[i] = pp[i] = b[i]; normb += (long double)b[i] * b[i]; }
        long double resid = normb;
        int it = 0;
        if (normb > 0)
            for (it = 1; it <= maxiter; it++) {
                long double pAp = 0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
                for (int i = 0; i < n; i++) {
                    long double acc = 0;
                    for (int64_t j = p->rowptr[i]; j < p->rowptr[i + 1]; j++) acc += (long double)val[j] * pp[p->col[j]];
                    Ap[i] = acc;
                    pAp += pp[i] * acc;
                }
                long double alpha = resid / pAp, rsnew = 0;
                for (int i = 0; i < n; i++) { x[i] += alpha * pp[i]; r[i] -= alpha * Ap[i]; rsnew += r[i] * r[i]; }
                if (rsnew < tol * normb) break;
                long double beta = rsnew / resid;
                for (int i = 0; i < n; i++) pp[i] = r[i] + beta * pp[i];
                resid = rsnew;
            }
        free(x); free(r); free(pp); free(Ap);
        return it > maxiter ? maxiter : it;
    }
    double *x = calloc(n, sizeof(double)), *r = malloc(n * sizeof(double)), *pp = malloc(n * sizeof(double)), *Ap = malloc(n * sizeof(double));
    memcpy(r, b, n * sizeof(double));
    memcpy(pp, b, n * sizeof(double));
    double normb = variant ? dot_pairwise(b, b, n) : dot(n, b, b);
    double resid = normb;
    int it = 0;
    if (normb > 0)
        for (it = 1; it <= maxiter; it++) {
            if (variant) {
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; i++) {
                    double acc = 0.0;
                    for (int64_t j = p->rowptr[i]; j < p->rowptr[i + 1]; j++) acc = fma(val[j], pp[p->col[j]], acc);
                    Ap[i] = acc;
                }
            } else {
                spmv(p, val, pp, Ap);
            }
            double alpha = resid / (variant ? dot_pairwise(pp, Ap, n) : dot(n, pp, Ap));
            for (int i = 0; i < n; i++) {
                if (variant) { x[i] = fma(alpha, pp[i], x[i]); r[i] = fma(-alpha, Ap[i], r[i]); }
                else { x[i] += alpha * pp[i]; r[i] -= alpha * Ap[i]; }
            }
            double rsnew = variant ? dot_pairwise(r, r, n) : dot(n, r, r);
            if (rsnew < tol * normb) break;
            double beta = rsnew / resid;
            for (int i = 0; i < n; i++) pp[i] = variant ? fma(beta, pp[i], r[i]) : r[i] + beta * pp[i];
            resid = rsnew;
        }
    free(x); free(r); free(pp); free(Ap);
    return it > maxiter ? maxiter : it;
}

/* CG iteration counts of the three arithmetic variants for the Newton step at displacements U */
void so_cg_variants(int N, int T, const double *nodes, const int32_t *tets, const uint8_t *movable, const double *f_target,
                    const double *U, int Nb, const double *beams, double k, double d0, double ls, double ds, double tol,
                    int32_t *iters /* [3] */) {
    so_material mat = {k, d0, ls, ds};
    so_pattern *p = pattern_build(N, T, tets, movable);
    double *Phi = malloc(sizeof(double) * 12 * (size_t)T), *V = malloc(sizeof(double) * (size_t)T), *E = malloc(sizeof(double) * (size_t)T);
    double *f_el = malloc(sizeof(double) * 12 * (size_t)T), *K_el = malloc(sizeof(double) * 144 * (size_t)T);
    double *val = malloc(sizeof(double) * (size_t)p->nnz), *forces = malloc(sizeof(double) * 3 * (size_t)N), *ff = malloc(sizeof(double) * 3 * (size_t)N);
    so_shape_tensors(T, nodes, tets, Phi, V);
    update_f_and_K(p, tets, U, Phi, V, Nb, beams, &mat, E, f_el, K_el, forces, val);
    for (int i = 0; i < N; i++)
        for (int c = 0; c < 3; c++) ff[3 * i + c] = movable[i] ? forces[3 * i + c] - f_target[3 * i + c] : 0.0;
    for (int v = 0; v < 3; v++) iters[v] = cg_variant(p, val, ff, 3 * N, tol, v);
    free(Phi); free(V); free(E); free(f_el); free(K_el); free(val); free(forces); free(ff);
    pattern_free(p);
}
