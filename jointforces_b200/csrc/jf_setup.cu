// One-off structure build for a context: locality reordering, shape tensors,
// SELL-32 block-matrix pattern with its deterministic gather lists (replaces saenopy's
// _compute_connections + per-iteration COO->CSR), and the folded direction table
// (SURVEY A2).  Host code; runs once per mesh, not on the hot path.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <queue>

#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <thread>

#include "jf_internal.h"

// split [0,n) over host threads (structure build only; results do not depend on the thread count)
template <typename F>
static void parallel_for(int64_t n, F fn) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 2 || n < 4096) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (unsigned i = 0; i < nt; i++) {
        const int64_t a = i * chunk, b = std::min<int64_t>(n, a + chunk);
        if (a < b) th.emplace_back([=] { fn(a, b); });
    }
    for (auto &t : th) t.join();
}

// ------------------------------------------------------------------------------------------------
// direction sphere, saenopy build_beams(300)
// ------------------------------------------------------------------------------------------------
static std::vector<double> default_beams(int n_req) {
    int n = (int)floor(sqrt(n_req * M_PI + 0.5));
    std::vector<double> b;
    for (int i = 0; i < n; i++) {
        double theta = (2 * M_PI / n) * i;
        int jmax = (int)floor(n * sin(theta) + 0.5);
        for (int j = 0; j < jmax; j++) {
            double rho = (2 * M_PI / jmax) * j;
            b.push_back(sin(theta) * cos(rho));
            b.push_back(sin(theta) * sin(rho));
            b.push_back(cos(theta));
        }
    }
    return b;
}

// index of the quartic monomial x^ex y^ey z^ez (ex+ey+ez = 4) in the 15-entry table
static inline int mono_id(int ex, int ey) { return (4 - ex) * (5 - ex) / 2 + (4 - ex - ey); }

static int build_dirtab(jf_ctx *c, const double *beams, int n_beams) {
    std::vector<double> own;
    if (!beams) {
        own = default_beams(300);
        beams = own.data();
        n_beams = (int)(own.size() / 3);
    }
    if (n_beams <= 0) {
        jf_set_error("n_beams must be positive");
        return JF_ERR_ARG;
    }
    c->n_beams = n_beams;
    // fold antipodal pairs: every term of A5-A8 is even in s, so s and -s contribute identically
    std::vector<int> rep;
    std::vector<double> wt;
    std::vector<char> used(n_beams, 0);
    for (int i = 0; i < n_beams; i++) {
        if (used[i]) continue;
        used[i] = 1;
        double w = 1.0;
        if (!(c->flags & JF_FLAG_NO_FOLD)) {
            for (int j = i + 1; j < n_beams; j++) {
                if (used[j]) continue;
                double d = 0;
                for (int k = 0; k < 3; k++) {
                    double e = beams[3 * i + k] + beams[3 * j + k];
                    d += e * e;
                }
                if (d < 1e-24) {
                    used[j] = 1;
                    w = 2.0;
                    break;
                }
            }
        }
        rep.push_back(i);
        wt.push_back(w);
    }
    c->n_dirs = (int)rep.size();
    int quantum = 16;  // lanes (1, 2, 4 or 8) x unroll 2
    c->n_dirs_pad = (c->n_dirs + quantum - 1) / quantum * quantum;
    // layout: (JF_NTAB/2) planes of n_dirs_pad double2
    std::vector<double> tab((size_t)JF_NTAB * c->n_dirs_pad, 0.0);
    auto put = [&](int comp, int b, double v) { tab[((size_t)(comp >> 1) * c->n_dirs_pad + b) * 2 + (comp & 1)] = v; };
    for (int b = 0; b < c->n_dirs_pad; b++) {
        double s[3] = {1.0, 0.0, 0.0}, w = 0.0;  // padding: harmless direction with zero weight
        if (b < c->n_dirs) {
            for (int k = 0; k < 3; k++) s[k] = beams[3 * rep[b] + k];
            w = wt[b];
        }
        put(0, b, s[0] * s[0]);
        put(1, b, s[1] * s[1]);
        put(2, b, s[2] * s[2]);
        put(3, b, s[0] * s[1]);
        put(4, b, s[0] * s[2]);
        put(5, b, s[1] * s[2]);
        for (int ex = 4; ex >= 0; ex--)
            for (int ey = 4 - ex; ey >= 0; ey--) {
                int ez = 4 - ex - ey;
                double v = 1.0;
                for (int q = 0; q < ex; q++) v *= s[0];
                for (int q = 0; q < ey; q++) v *= s[1];
                for (int q = 0; q < ez; q++) v *= s[2];
                put(6 + mono_id(ex, ey), b, v);
            }
        put(21, b, w);
    }
    if (jf_dev_alloc(c, &c->d_dirtab, tab.size())) return JF_ERR_CUDA;
    JF_CUDA(cudaMemcpy(c->d_dirtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(tab.size() * sizeof(double));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// node adjacency (sorted, unique, including self)
// ------------------------------------------------------------------------------------------------
static void build_adjacency(int64_t N, int64_t T, const int32_t *tets, std::vector<int64_t> &ptr,
                            std::vector<int32_t> &adj) {
    std::vector<int64_t> cnt(N + 1, 0);
    for (int64_t t = 0; t < T; t++)
        for (int m = 0; m < 4; m++) cnt[tets[4 * t + m] + 1] += 4;
    for (int64_t i = 0; i < N; i++) cnt[i + 1] += cnt[i];
    std::vector<int32_t> raw((size_t)cnt[N]);
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
    for (int64_t t = 0; t < T; t++)
        for (int m = 0; m < 4; m++)
            for (int r = 0; r < 4; r++) raw[fill[tets[4 * t + m]]++] = tets[4 * t + r];
    ptr.assign(N + 1, 0);
    parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            auto a = raw.begin() + cnt[i], b = raw.begin() + cnt[i + 1];
            std::sort(a, b);
            ptr[i + 1] = std::unique(a, b) - a;
        }
    });
    for (int64_t i = 0; i < N; i++) ptr[i + 1] += ptr[i];
    adj.resize((size_t)ptr[N]);
    for (int64_t i = 0; i < N; i++) std::copy(raw.begin() + cnt[i], raw.begin() + cnt[i] + (ptr[i + 1] - ptr[i]), adj.begin() + ptr[i]);
}

// Cost of the SpMV's vector gathers for a candidate ordering of the free nodes: the number of
// distinct 128-byte lines touched by each 32-lane gather (one per slice and block column), summed.
// On B200 that is what bounds the CG kernel for these matrices (L1tex wavefronts; DESIGN.md
// "CG"), so the ordering is chosen by it rather than by bandwidth alone.  Also returns the padded
// slot count.
static double gather_cost(const std::vector<int32_t> &order, int64_t N, const std::vector<int64_t> &ptr,
                          const std::vector<int32_t> &adj, int64_t *slots_out) {
    const int64_t nf = (int64_t)order.size();
    std::vector<int32_t> o2n(N, -1);
    for (int64_t i = 0; i < nf; i++) o2n[order[i]] = (int32_t)i;
    std::vector<int32_t> cols[JF_SLICE];
    int32_t tmp[JF_SLICE];
    double lines = 0;
    int64_t slots = 0;
    for (int64_t s0 = 0; s0 < nf; s0 += JF_SLICE) {
        size_t w = 0;
        const int nrow = (int)std::min<int64_t>(JF_SLICE, nf - s0);
        for (int r = 0; r < nrow; r++) {
            cols[r].clear();
            const int32_t old = order[s0 + r];
            for (int64_t k = ptr[old]; k < ptr[old + 1]; k++)
                if (o2n[adj[k]] >= 0) cols[r].push_back(o2n[adj[k]]);
            std::sort(cols[r].begin(), cols[r].end());
            w = std::max(w, cols[r].size());
        }
        slots += (int64_t)w * JF_SLICE;
        for (size_t j = 0; j < w; j++) {
            int n = 0;
            for (int r = 0; r < nrow; r++)
                if (j < cols[r].size()) tmp[n++] = cols[r][j] >> 2;  // four 32-byte entries per line
            std::sort(tmp, tmp + n);
            lines += (double)(std::unique(tmp, tmp + n) - tmp);
        }
    }
    if (slots_out) *slots_out = slots;
    return lines;
}

// Candidate orderings of the free nodes: the caller's order, and Cuthill-McKee (breadth-first,
// neighbours by increasing degree); the one with the cheapest gathers wins.
static void order_nodes(jf_ctx *c, const uint8_t *movable, const std::vector<int64_t> &ptr,
                        const std::vector<int32_t> &adj) {
    int64_t N = c->N;
    std::vector<int32_t> order;
    order.reserve(N);
    for (int64_t i = 0; i < N; i++)
        if (movable[i]) order.push_back((int32_t)i);
    if (!(c->flags & JF_FLAG_NO_REORDER)) {
        std::vector<int32_t> cm;
        cm.reserve(order.size());
        std::vector<int32_t> deg(N, 0);
        for (int64_t i = 0; i < N; i++)
            if (movable[i])
                for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) deg[i] += movable[adj[k]] ? 1 : 0;
        std::vector<char> seen(N, 0);
        // seeds in order of increasing degree (first unvisited wins)
        std::vector<int32_t> seeds(order);
        std::stable_sort(seeds.begin(), seeds.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
        std::vector<int32_t> nb;
        for (int32_t seed : seeds) {
            if (seen[seed]) continue;
            seen[seed] = 1;
            size_t head = cm.size();
            cm.push_back(seed);
            while (head < cm.size()) {
                int32_t u = cm[head++];
                nb.clear();
                for (int64_t k = ptr[u]; k < ptr[u + 1]; k++) {
                    int32_t v = adj[k];
                    if (movable[v] && !seen[v]) {
                        seen[v] = 1;
                        nb.push_back(v);
                    }
                }
                std::stable_sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
                cm.insert(cm.end(), nb.begin(), nb.end());
            }
        }
        int64_t slots_id = 0, slots_cm = 0;
        double cost_id = 0, cost_cm = 0;
        {
            std::thread other([&] { cost_cm = gather_cost(cm, N, ptr, adj, &slots_cm); });
            cost_id = gather_cost(order, N, ptr, adj, &slots_id);
            other.join();
        }
        // a gathered line and ~2 padded block slots cost about the same L1tex time
        if (cost_cm + 0.5 * (double)slots_cm / JF_SLICE < cost_id + 0.5 * (double)slots_id / JF_SLICE) order.swap(cm);
    }
    c->Nf = (int64_t)order.size();
    for (int64_t i = 0; i < N; i++)
        if (!movable[i]) order.push_back((int32_t)i);
    c->node_new2old = order;
    c->node_old2new.assign(N, 0);
    for (int64_t i = 0; i < N; i++) c->node_old2new[order[i]] = (int32_t)i;
}

template <typename Tp>
static int upload(jf_ctx *c, Tp **dst, const std::vector<Tp> &src) {
    if (jf_dev_alloc(c, dst, src.size())) return JF_ERR_CUDA;
    if (!src.empty()) JF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)(src.size() * sizeof(Tp));
    return 0;
}

int jf_build_structure(jf_ctx *c, const double *nodes, const int32_t *tets_in, const uint8_t *movable,
                       const double *beams, int n_beams) {
    const int64_t N = c->N, T = c->T;
    for (int64_t i = 0; i < 4 * T; i++)
        if (tets_in[i] < 0 || tets_in[i] >= N) {
            jf_set_error("tetrahedron %lld references node %d outside [0,%lld)", (long long)(i / 4), tets_in[i], (long long)N);
            return JF_ERR_ARG;
        }
    StageTimer tm;
    int rc = build_dirtab(c, beams, n_beams);
    if (rc) return rc;

    std::vector<int64_t> aptr;
    std::vector<int32_t> adj;
    tm.mark("validate+dirtab");
    build_adjacency(N, T, tets_in, aptr, adj);
    tm.mark("adjacency");
    order_nodes(c, movable, aptr, adj);
    tm.mark("ordering");
    const int64_t Nf = c->Nf;
    const std::vector<int32_t> &o2n = c->node_old2new;

    // tets: renumber, order by smallest new node id (keeps U gathers and K_el reads local)
    std::vector<int32_t> torder(T);
    std::iota(torder.begin(), torder.end(), 0);
    if (!(c->flags & JF_FLAG_NO_REORDER)) {
        std::vector<int32_t> key(T);
        for (int64_t t = 0; t < T; t++) {
            int32_t k = o2n[tets_in[4 * t]];
            for (int m = 1; m < 4; m++) k = std::min(k, o2n[tets_in[4 * t + m]]);
            key[t] = k;
        }
        std::stable_sort(torder.begin(), torder.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    }
    c->tet_new2old = torder;
    std::vector<int32_t> tets((size_t)4 * T);
    for (int64_t t = 0; t < T; t++)
        for (int m = 0; m < 4; m++) tets[4 * t + m] = o2n[tets_in[4 * (int64_t)torder[t] + m]];
    std::vector<double> xyz((size_t)3 * N);
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < 3; k++) xyz[3 * i + k] = nodes[3 * (int64_t)c->node_new2old[i] + k];

    tm.mark("renumber tets");
    // shape tensors Phi = Chi * B^-1, V = |det B|/6  
    std::vector<double> phi((size_t)12 * T), vol(T);
    std::vector<int64_t> bad_tet(1, -1);
    parallel_for(T, [&](int64_t lo_t, int64_t hi_t) {
    for (int64_t t = lo_t; t < hi_t; t++) {
        const int32_t *n = &tets[4 * t];
        double B[3][3];
        for (int col = 0; col < 3; col++)
            for (int i = 0; i < 3; i++) B[i][col] = xyz[3 * (int64_t)n[col + 1] + i] - xyz[3 * (int64_t)n[0] + i];
        double det = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) +
                     B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
        if (det == 0.0 || !isfinite(det)) {
            bad_tet[0] = torder[t];  // any one of them will do for the message
            continue;
        }
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
        vol[t] = fabs(det) / 6.0;
        double *P = &phi[12 * t];
        for (int j = 0; j < 3; j++) {
            P[j] = -(inv[0][j] + inv[1][j] + inv[2][j]);
            P[3 + j] = inv[0][j];
            P[6 + j] = inv[1][j];
            P[9 + j] = inv[2][j];
        }
    }
    });
    if (bad_tet[0] >= 0) {
        jf_set_error("tetrahedron %d is degenerate (zero volume)", (int)bad_tet[0]);
        return JF_ERR_ARG;
    }

    tm.mark("shape tensors");
    // SELL-32 pattern over free rows / free columns, in new numbering
    c->n_slices = (Nf + JF_SLICE - 1) / JF_SLICE;
    c->Nf_pad = c->n_slices * JF_SLICE;
    std::vector<int64_t> rptr(Nf + 1, 0);   // free-neighbour lists in new ids, sorted
    std::vector<int32_t> rcol;
    rcol.reserve((size_t)aptr[N]);
    {
        std::vector<int32_t> tmp;
        for (int64_t i = 0; i < Nf; i++) {
            int32_t old = c->node_new2old[i];
            tmp.clear();
            for (int64_t k = aptr[old]; k < aptr[old + 1]; k++) {
                int32_t v = o2n[adj[k]];
                if (v < Nf) tmp.push_back(v);
            }
            std::sort(tmp.begin(), tmp.end());
            rcol.insert(rcol.end(), tmp.begin(), tmp.end());
            rptr[i + 1] = (int64_t)rcol.size();
        }
    }
    c->n_blocks = (int64_t)rcol.size();
    std::vector<int32_t> slice_ptr(c->n_slices + 1, 0);
    for (int64_t k = 0; k < c->n_slices; k++) {
        int64_t w = 0;
        for (int64_t i = k * JF_SLICE; i < std::min<int64_t>(Nf, (k + 1) * JF_SLICE); i++) w = std::max(w, rptr[i + 1] - rptr[i]);
        int64_t next = (int64_t)slice_ptr[k] + w * JF_SLICE;
        if (next * 9 > 2000000000LL) {
            jf_set_error("mesh too large for 32-bit slot offsets");
            return JF_ERR_ARG;
        }
        slice_ptr[k + 1] = (int32_t)next;
    }
    c->n_slots = slice_ptr[c->n_slices];
    std::vector<int32_t> sell_col((size_t)c->n_slots);
    for (int64_t k = 0; k < c->n_slices; k++) {
        int64_t w = (slice_ptr[k + 1] - slice_ptr[k]) / JF_SLICE;
        for (int64_t j = 0; j < w; j++)
            for (int lane = 0; lane < JF_SLICE; lane++) {
                int64_t row = k * JF_SLICE + lane;
                int32_t col = (int32_t)std::min<int64_t>(row, Nf - 1);  // padding: zero block on a valid column
                if (row < Nf && j < rptr[row + 1] - rptr[row]) col = rcol[rptr[row] + j];
                sell_col[slice_ptr[k] + j * JF_SLICE + lane] = col;
            }
    }
    auto slot_of = [&](int32_t row, int32_t col) -> int64_t {
        const int32_t *a = &rcol[rptr[row]], *b = &rcol[rptr[row + 1]];
        const int32_t *hit = std::lower_bound(a, b, col);
        return (int64_t)slice_ptr[row / JF_SLICE] + (hit - a) * JF_SLICE + row % JF_SLICE;
    };
    tm.mark("SELL pattern");
    // gather lists for the stiffness: slot <- ((t*10+pair)*2+transpose), ascending t
    static const int pair_of[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
    if ((int64_t)T * 20 > 2000000000LL) {
        jf_set_error("mesh too large for 32-bit element offsets");
        return JF_ERR_ARG;
    }
    // slot of every (t, m, r) contribution, -1 if it touches a fixed node (parallel: independent lookups)
    std::vector<int32_t> eslot((size_t)16 * T);
    parallel_for(T, [&](int64_t lo, int64_t hi) {
        for (int64_t t = lo; t < hi; t++)
            for (int m = 0; m < 4; m++) {
                const int32_t a = tets[4 * t + m];
                for (int r = 0; r < 4; r++) {
                    const int32_t b = tets[4 * t + r];
                    eslot[16 * t + 4 * m + r] = (a < Nf && b < Nf) ? (int32_t)slot_of(a, b) : -1;
                }
            }
    });
    // counting sort by slot, ascending t within a slot (deterministic summation order)
    std::vector<int32_t> cptr((size_t)c->n_slots + 1, 0);
    for (int64_t e = 0; e < 16 * T; e++)
        if (eslot[e] >= 0) cptr[eslot[e] + 1]++;
    for (int64_t q = 0; q < c->n_slots; q++) cptr[q + 1] += cptr[q];
    std::vector<int32_t> csrc((size_t)cptr[c->n_slots]);
    {
        std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
        for (int64_t t = 0; t < T; t++)
            for (int m = 0; m < 4; m++)
                for (int r = 0; r < 4; r++) {
                    const int32_t q = eslot[16 * t + 4 * m + r];
                    if (q >= 0) csrc[fill[q]++] = (int32_t)(((t * 10 + pair_of[m][r]) << 1) | (m > r));
                }
    }
    tm.mark("stiffness gather lists");
    // gather lists for nodal forces: node <- (t*4+m)
    std::vector<int32_t> fptr(N + 1, 0);
    for (int64_t i = 0; i < 4 * T; i++) fptr[tets[i] + 1]++;
    for (int64_t i = 0; i < N; i++) fptr[i + 1] += fptr[i];
    std::vector<int32_t> fsrc((size_t)4 * T);
    {
        std::vector<int32_t> fill(fptr.begin(), fptr.end() - 1);
        for (int64_t i = 0; i < 4 * T; i++) fsrc[fill[tets[i]]++] = (int32_t)i;
    }

    tm.mark("force gather lists");
    if ((rc = upload(c, &c->d_tets, tets))) return rc;
    if ((rc = upload(c, &c->d_nodes, xyz))) return rc;
    if ((rc = upload(c, &c->d_phi, phi))) return rc;
    if ((rc = upload(c, &c->d_vol, vol))) return rc;
    if ((rc = upload(c, &c->d_slice_ptr, slice_ptr))) return rc;
    if ((rc = upload(c, &c->d_sell_col, sell_col))) return rc;
    if ((rc = upload(c, &c->d_contrib_ptr, cptr))) return rc;
    if ((rc = upload(c, &c->d_contrib_src, csrc))) return rc;
    if ((rc = upload(c, &c->d_fn_ptr, fptr))) return rc;
    if ((rc = upload(c, &c->d_fn_src, fsrc))) return rc;
    tm.mark("uploads");
    c->h_rowlen.resize(Nf);
    for (int64_t i = 0; i < Nf; i++) c->h_rowlen[i] = (int32_t)(rptr[i + 1] - rptr[i]);
    c->h_sell_col.swap(sell_col);
    c->h_slice_ptr.swap(slice_ptr);
    return 0;
}
