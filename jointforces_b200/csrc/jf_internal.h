// Internal declarations shared by the translation units of libjf_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "jf_b200.h"

#define JF_MAX_SYS 64          // systems per context (lockstep batch); smem partial tables are sized by it
#define JF_SLICE 32            // SELL slice height = one warp of block rows
#define JF_NTAB 22             // doubles per direction-table entry: 6 pair products, 15 quartic monomials, weight
#define JF_ELEM_LANES 4        // lanes cooperating on one tetrahedron
#define JF_ELEM_BLOCK 128
#define JF_CG_BLOCK 512

#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <thread>

// JF_TIMING=1 in the environment prints the host-side setup stages to stderr
struct StageTimer {
    bool on;
    std::chrono::steady_clock::time_point t;
    StageTimer() : on(getenv("JF_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[jf setup] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// split [0,n) over host threads (structure build only; every use is written so that results do not depend on
// the thread count)
template <typename F>
static inline void jf_parallel_for(int64_t n, F fn, int64_t serial_below = 4096) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 2 || n < serial_below) {
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

void jf_set_error(const char *fmt, ...);
int jf_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define JF_CUDA(x)                                                              \
    do {                                                                        \
        cudaError_t _e = (x);                                                   \
        if (_e != cudaSuccess) return jf_cuda_fail(_e, #x, __FILE__, __LINE__); \
    } while (0)

struct jf_material {
    double k, d0, ls, ds, inv_d0, inv_ds;
};

// per-system scalars living on the device (one struct per system)
struct jf_sys_state {
    double energy;      // strain energy of the last evaluation
    double residual2;   // sum (f - f_target)^2 over free nodes
    double force2;      // sum f^2 over free nodes
    double du;          // step^2 * sum uu^2 of the last relaxation step
    double normb;       // CG: b.b
    double resid;       // CG: r.r carried between iterations
    int cg_iters;       // CG iterations of the last solve
    int active;         // 1 = still relaxing
    int n_rows;         // relaxation-record rows logged by the device-side driver (jf_reduce_kernel) in this jf_relax
    int pad;
};

// what the device-side relaxation driver needs (jf_reduce_kernel logs one record row per evaluation and applies the
// five-energy stop rule itself, so the host can enqueue Newton steps ahead instead of reading the state every step)
struct jf_relax_log {
    double *rows;            // (S, cap, 4) device: du, energy, residual^2 (row 0: sum f^2), CG iterations
    double *host_rows;       // the same, mapped pinned host memory (written by the kernel; polled by jf_relax_progress)
    int *host_flags;         // (S, 2) mapped: rows logged, still active
    int cap;                 // rows per system
    int max_iter;            // Newton steps allowed
    double conv_crit;
    int enabled;
};

struct jf_ctx {
    int device = 0;
    int n_sys = 1;
    int64_t N = 0, Nf = 0, T = 0;
    int n_beams = 0, n_dirs = 0, n_dirs_pad = 0;
    uint32_t flags = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};

    // host-side permutations (new <-> old)
    std::vector<int32_t> node_new2old, node_old2new, tet_new2old;

    // device: mesh
    int32_t *d_tets = nullptr;     // (T,4) new node ids
    double *d_nodes = nullptr;     // (N,3) new order
    double *d_phi = nullptr;       // (T,12)
    double *d_vol = nullptr;       // (T)
    double *d_dirtab = nullptr;    // (JF_NTAB/2, n_dirs_pad) double2: direction table
    // device: per-system state
    jf_material *d_mat = nullptr;  // (S)
    jf_sys_state *d_state = nullptr;
    jf_sys_state *h_state = nullptr;  // pinned mirror
    double *d_U = nullptr;         // (S,N,3)
    double *d_U0 = nullptr;        // (S,N,3) start values kept for jf_restore_displacements
    double *d_ft = nullptr;        // (S,N,3) target forces (0 on fixed nodes)
    double *d_f = nullptr;         // (S,N,3) internal forces
    double *d_Et = nullptr;        // (S,T) per-tet energy
    double *d_fel = nullptr;       // (S,T,12)
    double *d_Kel = nullptr;       // (S,T,90) 10 unique (m<=r) 3x3 blocks
    // device: SELL-32 block matrix structure (shared by all systems)
    int64_t n_slices = 0, n_slots = 0, n_blocks = 0;
    int32_t *d_slice_ptr = nullptr;    // (n_slices+1) slot offset of each slice
    int32_t *d_sell_col = nullptr;     // (n_slots) block-column (free node id); padding -> own row
    int32_t *d_contrib_ptr = nullptr;  // (n_slots+1)
    int32_t *d_contrib_src = nullptr;  // ((t*10+pair)*2 + transpose)
    int32_t *d_fn_ptr = nullptr;       // (N+1)
    int32_t *d_fn_src = nullptr;       // (t*4+m)
    double *d_val = nullptr;           // (S, n_slots*9)
    // device: CG vectors, (S, Nf_pad, 4) each
    int64_t Nf_pad = 0;
    double *d_b = nullptr, *d_x = nullptr, *d_r = nullptr, *d_p0 = nullptr, *d_p1 = nullptr, *d_Ap = nullptr;
    double *d_r1 = nullptr, *d_Ap1 = nullptr;  // second generation for the single-reduction CG kernel
    int cg_fused = 0, cg_fused_stream = 0, cg_fused_stream_grid = 0, cg_fused_stream_tmem = 0, cg_fs_sm_blocks = 0, cg_stream_cache = 0;
    size_t cg_stream_smem = 0;
    unsigned cg_launch_seq = 0;  // CG launches so far (epoch tag of the Ap rows, jf_cg_fused.cu)
    size_t cg_fused_smem = 0;
    double *d_partial = nullptr;       // (2, S, cg_grid) block partial sums
    double *d_red_partial = nullptr;   // (S, 32, 3) partial sums of jf_reduce_kernel
    unsigned long long *d_acc = nullptr;  // fixed-point accumulators on separate lines (JF_CG_ACC_STRIDE > 1)
    unsigned *d_red_ticket = nullptr;  // (S) arrival tickets of jf_reduce_kernel (self-resetting)
    unsigned int *d_sync = nullptr;    // barrier counter (16 B), fixed-point accumulators (byte 64), sums-only slots (byte 256)
    int cg_grid = 0, elem_grid = 0;
    int cg_stream_halo = 0, elem_lanes = JF_ELEM_LANES;
    int elem_variant = 0;  // JF_ELEM_VARIANT: unroll / occupancy experiments of the element kernel (4 lanes)
    int cg_wmax = 1, cg_resident = 0, cg_res_warps = 0, cg_hmax = 0, cg_ctas_per_sys = 0;
    int cg_foreign_max = 0;
    // two systems side by side (jf_cg_pair_kernel, jf_cg_pair.cu)
    int cg_pair = 0, cg_pair_ctas = 0, cg_pair_hmax = 0, cg_pair_fmax = 0, cg_pair_tail = 0;
    size_t cg_pair_smem = 0;
    int32_t *d_pair_halo_ptr = nullptr, *d_pair_halo_nodes = nullptr;
    uint16_t *d_pair_lcol = nullptr;  // most foreign halo nodes of any CTA (resident kernels)
    int32_t *d_halo_ptr = nullptr, *d_halo_nodes = nullptr;
    // radial curve extraction (jf_curves.cu)
    int32_t *d_bin_ptr = nullptr, *d_bin_nodes = nullptr;
    double *d_curve_scratch = nullptr, *d_curves = nullptr;
    int n_bins = 0, bin_total = 0;
    double curve_r_inner = 0.0;
    uint16_t *d_lcol = nullptr;
    size_t cg_res_smem = 0;

    // device-side relaxation driver
    jf_relax_log rlog = {};
    double *d_relrec = nullptr;
    void *h_progress = nullptr;   // mapped pinned block from the process-wide cache (jf_api.cu)
    size_t h_progress_bytes = 0;
    cudaEvent_t ev_ring[4 * 8] = {};

    bool have_material = false, evaluated = false;
    std::vector<char> have_targets;
    jf_stats stats = {};
    int64_t bytes_device = 0;

    // host copies kept for the parity hooks
    std::vector<int32_t> h_sell_col, h_slice_ptr, h_rowlen;
};

constexpr int JF_ACC_MAX_STRIDE = 1024;  // words between two fixed-point accumulator words (JF_CG_ACC_STRIDE)

// jf_setup.cu
int jf_build_structure(jf_ctx *c, const double *nodes, const int32_t *tets, const uint8_t *movable,
                       const double *beams, int n_beams);
// kernels (each returns after enqueueing on c->stream)
int jf_launch_element(jf_ctx *c);
int jf_launch_assembly(jf_ctx *c);
int jf_spmv_host(jf_ctx *c, int sys, const double *x, double *y);
int jf_launch_cg(jf_ctx *c, double tol, int64_t maxiter, double step, int apply_update);
int jf_cg_configure(jf_ctx *c);
int jf_elem_configure(jf_ctx *c);

template <typename Tp>
int jf_dev_alloc(jf_ctx *c, Tp **p, size_t n) {
    size_t bytes = n * sizeof(Tp);
    if (bytes == 0) bytes = sizeof(Tp);
    // stream-ordered allocation from the device's default pool: jf_ctx_create raises the pool's release
    // threshold, so a context created after another was destroyed reuses its memory instead of paying
    // cudaMalloc/cudaFree (tens to hundreds of ms with high variance for ~35 buffers)
    cudaError_t e = cudaMallocAsync((void **)p, bytes, c->stream);
    if (e != cudaSuccess) return jf_cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    // setup copies use blocking calls on the legacy stream while c->stream is non-blocking: make the
    // allocation visible to every stream before anyone touches it (idle stream: microseconds)
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return jf_cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    c->bytes_device += (int64_t)bytes;
    return 0;
}
