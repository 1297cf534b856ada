// C ABI of libjf_b200 (include/jf_b200.h): context lifetime, host<->device transfers in the
// caller's node/tet numbering, the Newton/relaxation driver and statistics.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>

#include "jf_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void jf_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int jf_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    jf_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return JF_ERR_CUDA;
}

extern "C" const char *jf_last_error(void) { return g_err; }
extern "C" int jf_version(void) { return 100; }

extern "C" int jf_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return jf_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    return n;
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CHECK_CTX(c)                          \
    do {                                      \
        if (!(c)) {                           \
            jf_set_error("null context");     \
            return JF_ERR_ARG;                \
        }                                     \
        JF_CUDA(cudaSetDevice((c)->device));  \
    } while (0)

#define CHECK_SYS(c, s)                                                          \
    do {                                                                         \
        if ((s) < 0 || (s) >= (c)->n_sys) {                                      \
            jf_set_error("system index %d outside [0,%d)", (s), (c)->n_sys);     \
            return JF_ERR_ARG;                                                   \
        }                                                                        \
    } while (0)

// ------------------------------------------------------------------------------------------------
// mapped pinned blocks for the relaxation progress (record rows written by the device, read by the host without a
// CUDA call).  cudaHostAlloc costs up to ~90 ms, so freed blocks are kept for the next context of the process.
// ------------------------------------------------------------------------------------------------
#include <mutex>
static std::mutex g_pin_mutex;
static std::vector<std::pair<void *, size_t>> g_pin_free;

static void *pinned_take(size_t bytes, size_t *got) {
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        for (size_t i = 0; i < g_pin_free.size(); i++)
            if (g_pin_free[i].second >= bytes) {
                void *p = g_pin_free[i].first;
                *got = g_pin_free[i].second;
                g_pin_free.erase(g_pin_free.begin() + i);
                return p;
            }
    }
    size_t cap = 1 << 16;
    while (cap < bytes) cap <<= 1;
    void *p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    *got = cap;
    return p;
}

static void pinned_give(void *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    g_pin_free.emplace_back(p, bytes);
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
static void ctx_free(jf_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    pinned_give(c->h_progress, c->h_progress_bytes);
    c->h_progress = nullptr;
    for (auto &e : c->ev_ring)
        if (e) cudaEventDestroy(e);
    void *ptrs[] = {c->d_tets, c->d_nodes, c->d_phi, c->d_vol, c->d_dirtab, c->d_mat, c->d_state, c->d_U, c->d_ft, c->d_f,
                    c->d_U0, c->d_Et, c->d_fel, c->d_Kel, c->d_slice_ptr, c->d_sell_col, c->d_contrib_ptr, c->d_contrib_src,
                    c->d_fn_ptr, c->d_fn_src, c->d_halo_ptr, c->d_halo_nodes, c->d_lcol, c->d_bin_ptr, c->d_bin_nodes, c->d_curve_scratch,
                    c->d_curves, c->d_val, c->d_b, c->d_x, c->d_r, c->d_p0, c->d_p1, c->d_Ap, c->d_r1, c->d_Ap1, c->d_partial, c->d_sync, c->d_red_partial, c->d_red_ticket,
                    c->d_relrec, c->d_acc, c->d_pair_halo_ptr, c->d_pair_halo_nodes, c->d_pair_lcol};
    if (c->stream) {
        for (void *p : ptrs)
            if (p) cudaFreeAsync(p, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    free(c->h_state);  // pageable on purpose: cudaMallocHost/cudaFreeHost cost up to ~90 ms per context
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void jf_ctx_destroy(jf_ctx *c) { ctx_free(c); }

extern "C" int jf_ctx_create(jf_ctx **out, int device, int n_sys, int64_t n_nodes, const double *nodes, int64_t n_tets,
                             const int32_t *tets, const uint8_t *movable, int n_beams, const double *beams, uint32_t flags) {
    if (!out || !nodes || !tets || !movable) {
        jf_set_error("jf_ctx_create: null argument");
        return JF_ERR_ARG;
    }
    *out = nullptr;
    if (n_sys < 1 || n_sys > JF_MAX_SYS) {
        jf_set_error("n_sys must be in [1,%d]", JF_MAX_SYS);
        return JF_ERR_ARG;
    }
    if (n_nodes < 4 || n_tets < 1 || n_nodes > 700000000LL) {
        jf_set_error("need at least 4 nodes and 1 tetrahedron");
        return JF_ERR_ARG;
    }
    int ndev = jf_device_count();
    if (ndev < 0) return ndev;
    if (device < 0 || device >= ndev) {
        jf_set_error("device %d not available (%d CUDA devices)", device, ndev);
        return JF_ERR_CUDA;
    }
    double t0 = now_ms();
    JF_CUDA(cudaSetDevice(device));
    jf_ctx *c = new jf_ctx();
    c->device = device;
    c->n_sys = n_sys;
    c->N = n_nodes;
    c->T = n_tets;
    c->flags = flags;
    int rc = 0;
    do {
        // single attributes, not cudaGetDeviceProperties: the full property query costs 100-350 ms now and then
        if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "attr", __FILE__, __LINE__); break; }
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (!coop) { jf_set_error("device lacks cooperative launch"); rc = JF_ERR_CUDA; break; }
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "stream", __FILE__, __LINE__); break; }
        for (auto &e : c->ev)
            if (cudaEventCreate(&e) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__); break; }
        if (rc) break;
        for (auto &e : c->ev_ring)
            if (cudaEventCreate(&e) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__); break; }
        if (rc) break;
        {   // Keep freed workspace in the device's default stream-ordered pool for the next context (a sweep creates and
            // destroys contexts; cudaMalloc/cudaFree cost 30-250 ms per context).  This raises the release threshold of a
            // pool the host application may share: the memory stays reserved until cudaMemPoolTrimTo or process exit.
            // JF_POOL_KEEP_MB bounds it (default: 8192 MB); JF_POOL_KEEP_MB=0 leaves the pool's threshold alone.
            cudaMemPool_t pool;
            const char *env = getenv("JF_POOL_KEEP_MB");
            const unsigned long long keep_mb = env ? strtoull(env, nullptr, 10) : 8192ull;
            if (keep_mb > 0 && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long keep = keep_mb << 20, cur = 0;
                if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        }
        StageTimer tm;
        if ((rc = jf_build_structure(c, nodes, tets, movable, beams, n_beams))) break;
        tm.mark("(structure total)");
        const size_t S = (size_t)n_sys, N = (size_t)c->N, T = (size_t)c->T;
        // (blocking cudaMemcpy/cudaMemset below run on the legacy stream; c->stream is non-blocking,
        //  so every batch of stream-ordered allocations is followed by a stream synchronize)
        if ((rc = jf_dev_alloc(c, &c->d_mat, S))) break;
        if ((rc = jf_dev_alloc(c, &c->d_state, S))) break;
        c->h_state = (jf_sys_state *)malloc(S * sizeof(jf_sys_state));
        if (!c->h_state) { jf_set_error("out of host memory"); rc = JF_ERR_NOMEM; break; }
        if ((rc = jf_dev_alloc(c, &c->d_U, S * N * 3))) break;
        if ((rc = jf_dev_alloc(c, &c->d_U0, S * N * 3))) break;
        if ((rc = jf_dev_alloc(c, &c->d_ft, S * N * 3))) break;
        if ((rc = jf_dev_alloc(c, &c->d_f, S * N * 3))) break;
        if ((rc = jf_dev_alloc(c, &c->d_Et, S * T))) break;
        if ((rc = jf_dev_alloc(c, &c->d_fel, S * T * 12))) break;
        if ((rc = jf_dev_alloc(c, &c->d_Kel, S * T * 90))) break;
        if ((rc = jf_dev_alloc(c, &c->d_val, S * (size_t)c->n_slots * 9))) break;
        const size_t V = S * (size_t)c->Nf_pad * 4;
        double **vecs[] = {&c->d_b, &c->d_x, &c->d_r, &c->d_p0, &c->d_p1, &c->d_Ap, &c->d_r1, &c->d_Ap1};
        for (double **v : vecs) {
            if ((rc = jf_dev_alloc(c, v, V))) break;
            if (cudaMemset(*v, 0, (V ? V : 1) * sizeof(double)) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "memset", __FILE__, __LINE__); break; }
        }
        if (rc) break;
        tm.mark("allocations+memsets");
        if ((rc = jf_cg_configure(c))) break;
        tm.mark("cg configure");
        if ((rc = jf_elem_configure(c))) break;
        if ((rc = jf_dev_alloc(c, &c->d_partial, std::max<size_t>(64 * S * (size_t)std::max(c->cg_grid, c->cg_fused_stream_grid), 2 * 8192)))) break;  // (2 parities, <= 8 replicas, <= 3 sums, S, grid)
        // barrier counter (16 B), 3 x 8 fixed-point accumulators at byte 64, then (2, S, grid) 16-byte slots at byte 256
        if ((rc = jf_dev_alloc(c, &c->d_sync, 64 + 8 * S * (size_t)c->cg_grid))) break;
        if ((rc = jf_dev_alloc(c, &c->d_acc, 16 * JF_ACC_MAX_STRIDE))) break;
        if ((rc = jf_dev_alloc(c, &c->d_red_partial, S * 32 * 3))) break;
        if ((rc = jf_dev_alloc(c, &c->d_red_ticket, S))) break;
        cudaMemsetAsync(c->d_red_ticket, 0, S * sizeof(unsigned), c->stream);  // counter + (2, S, grid) 16-byte slots
        cudaMemset(c->d_U, 0, S * N * 3 * sizeof(double));
        cudaMemset(c->d_U0, 0, S * N * 3 * sizeof(double));
        cudaMemset(c->d_ft, 0, S * N * 3 * sizeof(double));
        cudaMemset(c->d_f, 0, S * N * 3 * sizeof(double));
        memset(c->h_state, 0, S * sizeof(jf_sys_state));
        for (size_t s = 0; s < S; s++) c->h_state[s].active = 1;
        if (cudaMemcpy(c->d_state, c->h_state, S * sizeof(jf_sys_state), cudaMemcpyHostToDevice) != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "state", __FILE__, __LINE__); break; }
        c->have_targets.assign(S, 0);
        tm.mark("elem configure+state");
        if (cudaDeviceSynchronize() != cudaSuccess) { rc = jf_cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__); break; }
        tm.mark("device synchronize");
    } while (0);
    if (rc) {
        ctx_free(c);
        return rc;
    }
    c->stats.ms_setup = now_ms() - t0;
    *out = c;
    return JF_OK;
}

// ------------------------------------------------------------------------------------------------
// setters / getters (caller numbering <-> internal numbering)
// ------------------------------------------------------------------------------------------------
extern "C" int jf_set_material(jf_ctx *c, int sys, double k, double d0, double ls, double ds) {
    CHECK_CTX(c);
    if (sys != -1) CHECK_SYS(c, sys);
    if (!(k > 0) || !(d0 > 0) || !(ds > 0) || !isfinite(k)) {
        jf_set_error("material needs k > 0, d0 > 0, ds > 0");
        return JF_ERR_ARG;
    }
    jf_material m = {k, d0, ls, ds, 1.0 / d0, 1.0 / ds};
    for (int s = (sys < 0 ? 0 : sys); s < (sys < 0 ? c->n_sys : sys + 1); s++)
        JF_CUDA(cudaMemcpy(c->d_mat + s, &m, sizeof m, cudaMemcpyHostToDevice));
    c->have_material = true;  // per-system materials: the caller sets each system it uses
    c->evaluated = false;
    return JF_OK;
}

static int put_nodal(jf_ctx *c, double *dst_dev, int sys, const double *src, bool zero_fixed) {
    std::vector<double> tmp((size_t)c->N * 3);
    for (int64_t i = 0; i < c->N; i++) {
        const double *p = src + 3 * (int64_t)c->node_new2old[i];
        const bool fixed = i >= c->Nf;
        for (int k = 0; k < 3; k++) {
            double v = p[k];
            if (v != v || (zero_fixed && fixed)) v = 0.0;
            tmp[3 * i + k] = v;
        }
    }
    JF_CUDA(cudaMemcpyAsync(dst_dev + (size_t)sys * c->N * 3, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.bytes_h2d += (int64_t)(tmp.size() * sizeof(double));
    return JF_OK;
}

static int get_nodal(jf_ctx *c, const double *src_dev, int sys, double *dst) {
    std::vector<double> tmp((size_t)c->N * 3);
    JF_CUDA(cudaMemcpyAsync(tmp.data(), src_dev + (size_t)sys * c->N * 3, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.bytes_d2h += (int64_t)(tmp.size() * sizeof(double));
    for (int64_t i = 0; i < c->N; i++) {
        double *p = dst + 3 * (int64_t)c->node_new2old[i];
        for (int k = 0; k < 3; k++) p[k] = tmp[3 * i + k];
    }
    return JF_OK;
}

extern "C" int jf_set_targets(jf_ctx *c, int sys, const double *f_target) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!f_target) { jf_set_error("null f_target"); return JF_ERR_ARG; }
    int rc = put_nodal(c, c->d_ft, sys, f_target, true);
    if (rc) return rc;
    c->have_targets[sys] = 1;
    c->evaluated = false;
    return JF_OK;
}

extern "C" int jf_set_displacements(jf_ctx *c, int sys, const double *U) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!U) { jf_set_error("null U"); return JF_ERR_ARG; }
    c->evaluated = false;
    int rc = put_nodal(c, c->d_U, sys, U, false);
    if (rc) return rc;
    JF_CUDA(cudaMemcpyAsync(c->d_U0 + (size_t)sys * c->N * 3, c->d_U + (size_t)sys * c->N * 3, (size_t)c->N * 3 * sizeof(double),
                            cudaMemcpyDeviceToDevice, c->stream));
    return JF_OK;
}

extern "C" int jf_restore_displacements(jf_ctx *c, int sys) {
    CHECK_CTX(c);
    if (sys != -1) CHECK_SYS(c, sys);
    const size_t per = (size_t)c->N * 3;
    const size_t off = sys < 0 ? 0 : (size_t)sys * per, cnt = sys < 0 ? per * c->n_sys : per;
    JF_CUDA(cudaMemcpyAsync(c->d_U + off, c->d_U0 + off, cnt * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->evaluated = false;
    return JF_OK;
}

extern "C" int jf_get_displacements(jf_ctx *c, int sys, double *U) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!U) { jf_set_error("null U"); return JF_ERR_ARG; }
    return get_nodal(c, c->d_U, sys, U);
}

extern "C" int jf_get_forces(jf_ctx *c, int sys, double *f) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!f) { jf_set_error("null f"); return JF_ERR_ARG; }
    if (!c->evaluated) { jf_set_error("jf_get_forces before jf_evaluate/jf_relax"); return JF_ERR_STATE; }
    return get_nodal(c, c->d_f, sys, f);
}

// ------------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------------
static int evaluate_async(jf_ctx *c, bool timed) {
    int rc;
    if (timed) JF_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = jf_launch_element(c))) return rc;
    if (timed) JF_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = jf_launch_assembly(c))) return rc;
    if (timed) JF_CUDA(cudaEventRecord(c->ev[2], c->stream));
    return 0;
}

static int fetch_state(jf_ctx *c) {
    JF_CUDA(cudaMemcpyAsync(c->h_state, c->d_state, c->n_sys * sizeof(jf_sys_state), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.bytes_d2h += (int64_t)(c->n_sys * sizeof(jf_sys_state));
    return 0;
}

static int account_eval(jf_ctx *c, int n_active) {
    float a = 0, b = 0;
    JF_CUDA(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
    JF_CUDA(cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
    c->stats.ms_element += a;
    c->stats.ms_assembly += b;
    c->stats.tet_updates += (int64_t)n_active * c->T;
    return 0;
}

static int count_active(jf_ctx *c) {
    int n = 0;
    for (int s = 0; s < c->n_sys; s++) n += c->h_state[s].active ? 1 : 0;
    return n;
}

extern "C" int jf_evaluate(jf_ctx *c) {
    CHECK_CTX(c);
    if (!c->have_material) { jf_set_error("jf_evaluate before jf_set_material"); return JF_ERR_STATE; }
    for (int s = 0; s < c->n_sys; s++) {
        c->h_state[s].active = 1;
        JF_CUDA(cudaMemcpyAsync(&c->d_state[s].active, &c->h_state[s].active, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    int rc = evaluate_async(c, true);
    if (rc) return rc;
    if ((rc = fetch_state(c))) return rc;
    if ((rc = account_eval(c, count_active(c)))) return rc;
    c->evaluated = true;
    return JF_OK;
}

extern "C" int jf_get_energy(jf_ctx *c, int sys, double *energy, double *residual2, double *force2) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!c->evaluated) { jf_set_error("jf_get_energy before jf_evaluate/jf_relax"); return JF_ERR_STATE; }
    if (energy) *energy = c->h_state[sys].energy;
    if (residual2) *residual2 = c->h_state[sys].residual2;
    if (force2) *force2 = c->h_state[sys].force2;
    return JF_OK;
}

extern "C" int jf_get_element_quantities(jf_ctx *c, int sys, double *E, double *f_el, double *K_el) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!c->evaluated) { jf_set_error("jf_get_element_quantities before jf_evaluate"); return JF_ERR_STATE; }
    const int64_t T = c->T;
    JF_CUDA(cudaStreamSynchronize(c->stream));
    if (E) {
        std::vector<double> tmp(T);
        JF_CUDA(cudaMemcpy(tmp.data(), c->d_Et + (size_t)sys * T, T * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t t = 0; t < T; t++) E[c->tet_new2old[t]] = tmp[t];
    }
    if (f_el) {
        std::vector<double> tmp((size_t)T * 12);
        JF_CUDA(cudaMemcpy(tmp.data(), c->d_fel + (size_t)sys * T * 12, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t t = 0; t < T; t++) memcpy(f_el + 12 * (int64_t)c->tet_new2old[t], &tmp[12 * t], 12 * sizeof(double));
    }
    if (K_el) {
        static const int pair_of[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
        const int64_t chunk = 1 << 18;
        std::vector<double> tmp((size_t)std::min(chunk, T) * 90);
        for (int64_t t0 = 0; t0 < T; t0 += chunk) {
            int64_t n = std::min(chunk, T - t0);
            JF_CUDA(cudaMemcpy(tmp.data(), c->d_Kel + ((size_t)sys * T + t0) * 90, (size_t)n * 90 * sizeof(double), cudaMemcpyDeviceToHost));
            for (int64_t t = 0; t < n; t++) {
                double *dst = K_el + 144 * (int64_t)c->tet_new2old[t0 + t];
                const double *src = &tmp[90 * t];
                for (int m = 0; m < 4; m++)
                    for (int r = 0; r < 4; r++)
                        for (int i = 0; i < 3; i++)
                            for (int l = 0; l < 3; l++)
                                dst[((m * 4 + r) * 3 + i) * 3 + l] = (m <= r) ? src[pair_of[m][r] * 9 + 3 * i + l] : src[pair_of[m][r] * 9 + 3 * l + i];
            }
        }
    }
    return JF_OK;
}

extern "C" int jf_get_stiffness_blocks(jf_ctx *c, int sys, int64_t *n_blocks, int32_t *row_node, int32_t *col_node, double *blocks) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!n_blocks) { jf_set_error("null n_blocks"); return JF_ERR_ARG; }
    *n_blocks = c->n_blocks;
    if (!row_node && !col_node && !blocks) return JF_OK;
    if (!row_node || !col_node || !blocks) { jf_set_error("row_node, col_node and blocks must all be given"); return JF_ERR_ARG; }
    if (!c->evaluated) { jf_set_error("jf_get_stiffness_blocks before jf_evaluate"); return JF_ERR_STATE; }
    std::vector<double> val((size_t)c->n_slots * 9);
    JF_CUDA(cudaStreamSynchronize(c->stream));
    JF_CUDA(cudaMemcpy(val.data(), c->d_val + (size_t)sys * c->n_slots * 9, val.size() * sizeof(double), cudaMemcpyDeviceToHost));
    // walk rows; a row's real blocks are the leading h_rowlen entries of its SELL row
    int64_t out = 0;
    for (int64_t row = 0; row < c->Nf; row++) {
        int64_t k = row / JF_SLICE, lane = row % JF_SLICE;
        int64_t q0 = c->h_slice_ptr[k], w = (c->h_slice_ptr[k + 1] - q0) / JF_SLICE;
        (void)w;
        for (int64_t j = 0; j < c->h_rowlen[row]; j++) {
            int64_t q = q0 + j * JF_SLICE + lane;
            int32_t col = c->h_sell_col[q];
            if (out >= c->n_blocks) { jf_set_error("internal: block count mismatch"); return JF_ERR_STATE; }
            row_node[out] = c->node_new2old[row];
            col_node[out] = c->node_new2old[col];
            for (int comp = 0; comp < 9; comp++) blocks[9 * out + comp] = val[(q >> 5) * 288 + 32 * comp + (q & 31)];
            out++;
        }
    }
    if (out != c->n_blocks) { jf_set_error("internal: walked %lld blocks, expected %lld", (long long)out, (long long)c->n_blocks); return JF_ERR_STATE; }
    return JF_OK;
}

// ------------------------------------------------------------------------------------------------
// CG and relaxation
// ------------------------------------------------------------------------------------------------
static int ready_for_solve(jf_ctx *c) {
    if (!c->have_material) { jf_set_error("material not set"); return JF_ERR_STATE; }
    for (int s = 0; s < c->n_sys; s++)
        if (!c->have_targets[s]) { jf_set_error("targets of system %d not set", s); return JF_ERR_STATE; }
    return 0;
}

extern "C" int jf_cg(jf_ctx *c, double tol, int64_t maxiter, int32_t *iters) {
    CHECK_CTX(c);
    if (!c->evaluated) { jf_set_error("jf_cg before jf_evaluate"); return JF_ERR_STATE; }
    if (tol <= 0) tol = 1e-5;
    if (maxiter <= 0) maxiter = 3 * c->N;
    JF_CUDA(cudaEventRecord(c->ev[3], c->stream));
    int rc = jf_launch_cg(c, tol, maxiter, 0.0, 0);
    if (rc) return rc;
    JF_CUDA(cudaEventRecord(c->ev[4], c->stream));
    if ((rc = fetch_state(c))) return rc;
    float ms = 0;
    JF_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]));
    c->stats.ms_cg += ms;
    for (int s = 0; s < c->n_sys; s++) {
        if (iters) iters[s] = c->h_state[s].cg_iters;
        c->stats.cg_iterations += c->h_state[s].cg_iters;
    }
    return JF_OK;
}

extern "C" int jf_apply_stiffness(jf_ctx *c, int sys, const double *x, double *y) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!x || !y) { jf_set_error("null argument"); return JF_ERR_ARG; }
    if (!c->evaluated) { jf_set_error("jf_apply_stiffness before jf_evaluate"); return JF_ERR_STATE; }
    return jf_spmv_host(c, sys, x, y);
}

extern "C" int jf_get_cg_solution(jf_ctx *c, int sys, double *uu) {
    CHECK_CTX(c);
    CHECK_SYS(c, sys);
    if (!uu) { jf_set_error("null uu"); return JF_ERR_ARG; }
    std::vector<double> tmp((size_t)c->Nf_pad * 4);
    JF_CUDA(cudaStreamSynchronize(c->stream));
    JF_CUDA(cudaMemcpy(tmp.data(), c->d_x + (size_t)sys * c->Nf_pad * 4, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    memset(uu, 0, (size_t)c->N * 3 * sizeof(double));
    for (int64_t i = 0; i < c->Nf; i++)
        for (int k = 0; k < 3; k++) uu[3 * (int64_t)c->node_new2old[i] + k] = tmp[4 * i + k];
    return JF_OK;
}

// Room for the record rows of a relaxation of `max_iter` Newton steps: device rows + a mapped pinned mirror with the
// per-system progress flags behind it.
static int relax_log_prepare(jf_ctx *c, int max_iter, double conv_crit) {
    const int S = c->n_sys, cap = max_iter + 1;
    if (cap > c->rlog.cap || !c->d_relrec) {
        if (c->d_relrec) JF_CUDA(cudaFreeAsync(c->d_relrec, c->stream));
        c->d_relrec = nullptr;
        int rc = jf_dev_alloc(c, &c->d_relrec, (size_t)S * cap * 4);
        if (rc) return rc;
        c->rlog.cap = cap;
    }
    const size_t need = (size_t)S * c->rlog.cap * 4 * sizeof(double) + (size_t)S * 2 * sizeof(int);
    if (need > c->h_progress_bytes) {
        pinned_give(c->h_progress, c->h_progress_bytes);
        c->h_progress = pinned_take(need, &c->h_progress_bytes);
        if (!c->h_progress) { c->h_progress_bytes = 0; jf_set_error("cannot allocate the pinned progress block"); return JF_ERR_NOMEM; }
    }
    memset(c->h_progress, 0, need);
    void *dev = nullptr;
    JF_CUDA(cudaHostGetDevicePointer(&dev, c->h_progress, 0));
    c->rlog.rows = c->d_relrec;
    c->rlog.host_rows = (double *)dev;
    c->rlog.host_flags = (int *)((char *)dev + (size_t)S * c->rlog.cap * 4 * sizeof(double));
    c->rlog.max_iter = max_iter;
    c->rlog.conv_crit = conv_crit;
    return 0;
}

static inline const double *progress_rows(const jf_ctx *c, int sys) { return (const double *)c->h_progress + (size_t)sys * c->rlog.cap * 4; }
static inline const volatile int *progress_flags(const jf_ctx *c) {
    return (const volatile int *)((const char *)c->h_progress + (size_t)c->n_sys * c->rlog.cap * 4 * sizeof(double));
}

// The relaxation (A12, A14).  The driver's bookkeeping runs on the device (jf_reduce_kernel: record row, five-energy stop
// rule, `active` flag that every kernel honours), so the host enqueues RELAX_AHEAD Newton steps at a time and only then
// looks at the progress the device wrote into mapped memory -- steps enqueued past a system's stop are no-ops.
constexpr int RELAX_AHEAD = 4;

extern "C" int jf_relax(jf_ctx *c, double step, int max_iter, double conv_crit, double cg_tol, double *relrec, int32_t *cg_iters,
                        int32_t *n_iter) {
    CHECK_CTX(c);
    if (!relrec || !n_iter || max_iter < 0) { jf_set_error("jf_relax: bad arguments"); return JF_ERR_ARG; }
    int rc = ready_for_solve(c);
    if (rc) return rc;
    if (cg_tol <= 0) cg_tol = 1e-5;
    const int S = c->n_sys;
    const int64_t cg_maxiter = 3 * c->N;
    const size_t stride = (size_t)(max_iter + 1) * 3;
    if ((rc = relax_log_prepare(c, max_iter, conv_crit))) return rc;
    // all systems start active, no rows logged
    for (int s = 0; s < S; s++) {
        c->h_state[s].active = 1;
        c->h_state[s].n_rows = 0;
    }
    JF_CUDA(cudaEventRecord(c->ev[5], c->stream));
    JF_CUDA(cudaMemcpyAsync(c->d_state, c->h_state, S * sizeof(jf_sys_state), cudaMemcpyHostToDevice, c->stream));
    c->rlog.enabled = 1;
    struct Off { jf_ctx *c; ~Off() { c->rlog.enabled = 0; } } off{c};   // plain jf_evaluate calls do not log
    if ((rc = evaluate_async(c, true))) return rc;                      // row 0 (a system with max_iter == 0 stops here)
    JF_CUDA(cudaEventSynchronize(c->ev[2]));
    if ((rc = account_eval(c, S))) return rc;
    c->evaluated = true;
    const volatile int *flags = progress_flags(c);
    int steps = 0;
    auto n_active = [&]() { int n = 0; for (int s = 0; s < S; s++) n += flags[2 * s + 1] ? 1 : 0; return n; };
    while (steps < max_iter && n_active() > 0) {
        const int batch = std::min(RELAX_AHEAD, max_iter - steps);
        for (int k = 0; k < batch; k++) {
            cudaEvent_t *e = c->ev_ring + 4 * k;
            JF_CUDA(cudaEventRecord(e[0], c->stream));
            if ((rc = jf_launch_cg(c, cg_tol, cg_maxiter, step, 1))) return rc;
            JF_CUDA(cudaEventRecord(e[1], c->stream));
            if ((rc = jf_launch_element(c))) return rc;
            JF_CUDA(cudaEventRecord(e[2], c->stream));
            if ((rc = jf_launch_assembly(c))) return rc;
            JF_CUDA(cudaEventRecord(e[3], c->stream));
        }
        JF_CUDA(cudaEventSynchronize(c->ev_ring[4 * (batch - 1) + 3]));
        for (int k = 0; k < batch; k++) {
            cudaEvent_t *e = c->ev_ring + 4 * k;
            float a = 0, b = 0, d = 0;
            JF_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
            JF_CUDA(cudaEventElapsedTime(&b, e[1], e[2]));
            JF_CUDA(cudaEventElapsedTime(&d, e[2], e[3]));
            c->stats.ms_cg += a;
            c->stats.ms_element += b;
            c->stats.ms_assembly += d;
        }
        steps += batch;
    }
    JF_CUDA(cudaEventRecord(c->ev[6], c->stream));
    if ((rc = fetch_state(c))) return rc;   // energies etc. of the final state for the getters; synchronises the stream
    float total = 0;
    JF_CUDA(cudaEventElapsedTime(&total, c->ev[5], c->ev[6]));
    c->stats.ms_relax += total;
    for (int s = 0; s < S; s++) {
        const int rows = flags[2 * s];
        const double *src = progress_rows(c, s);
        n_iter[s] = rows - 1;
        for (int i = 0; i < rows; i++) {
            double *row = relrec + s * stride + 3 * (size_t)i;
            row[0] = src[4 * i];
            row[1] = src[4 * i + 1];
            row[2] = src[4 * i + 2];
            if (i > 0) {
                if (cg_iters) cg_iters[(size_t)s * max_iter + (i - 1)] = (int32_t)src[4 * i + 3];
                c->stats.cg_iterations += (int64_t)src[4 * i + 3];
            }
        }
        c->stats.newton_steps += rows - 1;
        c->stats.tet_updates += (int64_t)(rows - 1) * c->T;
        c->h_state[s].active = 1;   // (the device flag is reset by the next jf_relax / jf_evaluate)
    }
    return JF_OK;
}

// Rows of the relaxation record the device has logged so far for system `sys` (a jf_relax running in ANOTHER thread
// fills them; no CUDA call here): copies up to max_rows rows of (du, energy, residual^2) into `rows` and returns
// their number in *n_rows.  saenopy rewrites relrec.txt and calls its callback after every Newton step
// (simulation.py:167); the Python side polls this to do the same while the GPU runs ahead.
extern "C" int jf_relax_progress(jf_ctx *c, int sys, int32_t *n_rows, double *rows, int max_rows) {
    if (!c || !n_rows) { jf_set_error("null argument"); return JF_ERR_ARG; }
    CHECK_SYS(c, sys);
    if (!c->h_progress) { *n_rows = 0; return JF_OK; }
    const volatile int *flags = progress_flags(c);
    int n = flags[2 * sys];
    if (n > c->rlog.cap) n = c->rlog.cap;
    std::atomic_thread_fence(std::memory_order_acquire);
    if (rows) {
        const double *src = progress_rows(c, sys);
        const int m = std::min(n, max_rows);
        for (int i = 0; i < m; i++)
            for (int k = 0; k < 3; k++) rows[3 * i + k] = src[4 * i + k];
        n = m;
    }
    *n_rows = n;
    return JF_OK;
}

// ------------------------------------------------------------------------------------------------
// stats / info
// ------------------------------------------------------------------------------------------------
extern "C" int jf_get_stats(jf_ctx *c, jf_stats *out) {
    if (!c || !out) { jf_set_error("null argument"); return JF_ERR_ARG; }
    *out = c->stats;
    return JF_OK;
}

extern "C" int jf_reset_stats(jf_ctx *c) {
    if (!c) { jf_set_error("null context"); return JF_ERR_ARG; }
    double setup = c->stats.ms_setup;
    c->stats = jf_stats();
    c->stats.ms_setup = setup;
    return JF_OK;
}

extern "C" int jf_get_info(jf_ctx *c, jf_info *o) {
    if (!c || !o) { jf_set_error("null argument"); return JF_ERR_ARG; }
    o->n_nodes = c->N;
    o->n_free = c->Nf;
    o->n_tets = c->T;
    o->n_blocks = c->n_blocks;
    o->n_slots = c->n_slots;
    o->n_sys = c->n_sys;
    o->n_dirs = c->n_dirs;
    o->n_beams = c->n_beams;
    o->sm_count = c->sm_count;
    o->cg_grid = c->cg_grid;
    o->cg_block = JF_CG_BLOCK;
    o->elem_grid = c->elem_grid;
    o->elem_block = JF_ELEM_BLOCK;
    o->bytes_device = c->bytes_device;
    o->cg_resident = c->cg_resident;
    o->cg_wmax = c->cg_wmax;
    o->cg_stream_halo = c->cg_stream_halo;
    o->cg_fused = c->cg_pair ? 4 : c->cg_fused + 2 * c->cg_fused_stream;  // 1: register-resident, 2: streaming single-reduction kernel, 4: two systems side by side
    return JF_OK;
}

extern "C" int jf_solve_boundarycondition(int device, int64_t n_nodes, const double *nodes, int64_t n_tets, const int32_t *tets,
                                          const uint8_t *movable, const double *f_target, double *U, int n_beams,
                                          const double *beams, double k, double d0, double lambda_s, double ds, double step,
                                          int max_iter, double conv_crit, double *relrec, int32_t *cg_iters, int32_t *n_iter,
                                          double *forces_out, jf_stats *stats_out) {
    jf_ctx *c = nullptr;
    const bool timing = getenv("JF_TIMING") != nullptr;
    double t0 = now_ms(), t1 = t0, t2 = t0, t3 = t0, t4 = t0;
    int rc = jf_ctx_create(&c, device, 1, n_nodes, nodes, n_tets, tets, movable, n_beams, beams, 0);
    if (rc) return rc;
    t1 = now_ms();
    do {
        if ((rc = jf_set_material(c, 0, k, d0, lambda_s, ds))) break;
        if ((rc = jf_set_targets(c, 0, f_target))) break;
        if ((rc = jf_set_displacements(c, 0, U))) break;
        t2 = now_ms();
        if ((rc = jf_relax(c, step, max_iter, conv_crit, 0.0, relrec, cg_iters, n_iter))) break;
        t3 = now_ms();
        if ((rc = jf_get_displacements(c, 0, U))) break;
        if (forces_out && (rc = jf_get_forces(c, 0, forces_out))) break;
        if (stats_out) jf_get_stats(c, stats_out);
        t4 = now_ms();
    } while (0);
    jf_ctx_destroy(c);
    if (timing)
        fprintf(stderr, "[jf solve] create %.1f  set %.1f  relax %.1f  get %.1f  destroy %.1f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3,
                now_ms() - t4);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA peak (roofline denominator of the element kernel; MEASURED_PEAKS.json has no FP64 figure)
// ------------------------------------------------------------------------------------------------
__global__ void jf_fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int jf_bench_fp64(int device, double *gflops) {
    if (!gflops) { jf_set_error("null gflops"); return JF_ERR_ARG; }
    int ndev = jf_device_count();
    if (ndev < 0) return ndev;
    if (device < 0 || device >= ndev) { jf_set_error("device %d not available", device); return JF_ERR_CUDA; }
    JF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    JF_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    double *d = nullptr;
    JF_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    JF_CUDA(cudaEventCreate(&e0));
    JF_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        JF_CUDA(cudaEventRecord(e0));
        jf_fp64_peak_kernel<<<blocks, threads>>>(d, iters);
        JF_CUDA(cudaEventRecord(e1));
        JF_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        JF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double gf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) * 1e-9;
        if (rep > 0 && gf > best) best = gf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *gflops = best;
    return JF_OK;
}
