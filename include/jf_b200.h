/*
 * jf_b200 -- C ABI of the B200-native relaxation library behind jointforces' simulation path.
 *
 * The reference (christophmark/jointforces) has no FFI: its hot path is the Python object
 * protocol of the third-party `saenopy.Solver` as called from
 * /root/reference/jointforces/simulation.py:100-104,143,146,167-168.  Each entry point below
 * cites the call it replaces.  All pointers are HOST pointers unless the name ends in `_dev`;
 * arrays are C-contiguous; the caller owns every buffer it passes; the library owns only its
 * workspace (device memory + a stream per context).  One context <-> one CUDA stream <-> one
 * host thread at a time; contexts are independent, several may be live on one GPU.
 *
 * A context holds ONE mesh and boundary mask and `n_sys` independent load cases ("systems") on
 * it -- the pressures of a lookup-table sweep (simulation.py:171-249) -- which are evaluated,
 * assembled and relaxed in lockstep by the same kernels.
 *
 * Every function returns 0 on success or a negative jf_status; jf_last_error() gives the
 * message of the last failure on the calling thread.  Nothing here ever calls exit().
 * There is no CPU fallback: without a CUDA device every compute entry fails with
 * JF_ERR_CUDA.
 */
#ifndef JF_B200_H
#define JF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jf_ctx jf_ctx;

enum jf_status {
    JF_OK = 0,
    JF_ERR_ARG = -1,    /* bad argument (null pointer, size, index out of range, degenerate tet) */
    JF_ERR_CUDA = -2,   /* CUDA runtime error or no device */
    JF_ERR_STATE = -3,  /* call out of order (e.g. relax before material/targets are set) */
    JF_ERR_NOMEM = -4,
    JF_ERR_IO = -5,     /* file cannot be opened or read */
    JF_ERR_FORMAT = -6  /* input is outside what the fast path handles; the caller falls back to the reference's own loop */
};

/* jf_ctx_create flags */
#define JF_FLAG_NO_REORDER 1u /* keep the caller's node/tet order (default: locality reordering) */
#define JF_FLAG_NO_FOLD 2u    /* do not merge antipodal directions (default: merge, exact)        */
#define JF_FLAG_NO_RESIDENT 4u /* never use the shared-memory-resident CG kernel (testing)        */

const char *jf_last_error(void);
int jf_version(void);
/* number of CUDA devices visible, or JF_ERR_CUDA */
int jf_device_count(void);

/*
 * Replaces Solver() + set_nodes + set_tetrahedra + the `movable` part of
 * set_boundary_condition (simulation.py:100,103,104,143).
 *   nodes   (n_nodes,3) float64 [m]; tets (n_tets,4) int32, 0-based;
 *   movable (n_nodes)  1 = free node (NaN displacement condition), 0 = fixed node;
 *   beams   (n_beams,3) unit directions, or NULL for saenopy's default build_beams(300) = 286.
 */
int jf_ctx_create(jf_ctx **out, int device, int n_sys, int64_t n_nodes, const double *nodes, int64_t n_tets,
                  const int32_t *tets, const uint8_t *movable, int n_beams, const double *beams, uint32_t flags);
void jf_ctx_destroy(jf_ctx *ctx);

/* Replaces set_material_model(SemiAffineFiberMaterial(k, d0, lambda_s, ds)) (simulation.py:101-102).
 * sys = -1 sets every system. */
int jf_set_material(jf_ctx *ctx, int sys, double k, double d0, double lambda_s, double ds);

/* Replaces the force half of set_boundary_condition (simulation.py:132-143): f_target (n_nodes,3);
 * entries of fixed nodes are ignored (NaN allowed there). */
int jf_set_targets(jf_ctx *ctx, int sys, const double *f_target);

/* Replaces the displacement half of set_boundary_condition and set_initial_displacements
 * (simulation.py:127-129,146): U (n_nodes,3) for ALL nodes (fixed nodes carry their prescribed
 * value, free nodes the start value).  NaN entries are read as 0. */
int jf_set_displacements(jf_ctx *ctx, int sys, const double *U);
int jf_get_displacements(jf_ctx *ctx, int sys, double *U);
/* device-side reset of U to what the last jf_set_displacements stored (sys = -1: all systems) */
int jf_restore_displacements(jf_ctx *ctx, int sys);
/* internal nodal forces f (n_nodes,3) of the last evaluation (Solver.mesh.forces, simulation.py:255) */
int jf_get_forces(jf_ctx *ctx, int sys, double *f);

/* _update_glo_f_and_K for every active system: element kernel + assembly. */
int jf_evaluate(jf_ctx *ctx);
/* strain energy, sum (f-f_target)^2 over free nodes, sum f^2 over free nodes of the last evaluation */
int jf_get_energy(jf_ctx *ctx, int sys, double *energy, double *residual2, double *force2);

/* parity hooks (original node/tet order): per-tet E (T), f_el (T,4,3), K_el (T,4,4,3,3); any may be NULL */
int jf_get_element_quantities(jf_ctx *ctx, int sys, double *E, double *f_el, double *K_el);
/* assembled stiffness as block COO in original node ids: call with NULL arrays to get the count */
int jf_get_stiffness_blocks(jf_ctx *ctx, int sys, int64_t *n_blocks, int32_t *row_node, int32_t *col_node,
                            double *blocks /* (n_blocks,3,3) */);

/* cg(K_glo, f - f_target, maxiter, tol) from x = 0 on the last evaluation; does not move U.
 * iters: (n_sys).  Solution via jf_get_cg_solution (n_nodes,3; zero on fixed nodes). */
int jf_cg(jf_ctx *ctx, double tol, int64_t maxiter, int32_t *iters);
int jf_get_cg_solution(jf_ctx *ctx, int sys, double *uu);
/* parity hook: y (n_nodes,3) = K_glo x with the stiffness of the last evaluation (free rows and columns only) */
int jf_apply_stiffness(jf_ctx *ctx, int sys, const double *x, double *y);

/*
 * Replaces solve_boundarycondition(step_size, max_iterations, rel_conv_crit) (simulation.py:167;
 * SURVEY A12, A14) for all systems in lockstep; each system stops on its own five-energy rule.
 *   cg_tol    <= 0 selects saenopy's 1e-5 (on squared norms); cg maxiter is 3*n_nodes.
 *   relrec    (n_sys, max_iter+1, 3)  rows [du, strain energy, residual^2]; row 0 is the start state
 *   cg_iters  (n_sys, max_iter)       CG iterations of each Newton step (may be NULL)
 *   n_iter    (n_sys)                 Newton steps taken
 */
int jf_relax(jf_ctx *ctx, double step, int max_iter, double conv_crit, double cg_tol, double *relrec,
             int32_t *cg_iters, int32_t *n_iter);

/*
 * Progress of a jf_relax that is running in another host thread: rows of the relaxation record (du, energy,
 * residual^2) the device has logged so far for system `sys` (up to max_rows are copied; rows may be NULL to ask for the
 * count only).  Makes no CUDA call.  Replaces what saenopy does inside solve_boundarycondition after every Newton
 * step -- np.savetxt(relrecname, relrec) and callback(...) (jointforces/simulation.py:167 passes relrecname) -- the
 * caller polls this and does both on the host while the GPU runs ahead.
 */
int jf_relax_progress(jf_ctx *ctx, int sys, int32_t *n_rows, double *rows, int max_rows);

/*
 * Radial displacement curves on the device -- what extract_deformation_curve_solver computes on the host
 * (simulation.py:296-299): per bin the median of |u|/r_inner over the nodes listed for that bin.
 * The caller classifies the nodes once per mesh: bin_ptr (n_bins+1), bin_nodes (bin_ptr[n_bins]) in the caller's
 * node numbering.  jf_get_curves fills curves (n_sys, n_bins); empty bins give NaN like np.nanmedian.
 */
int jf_set_curve_bins(jf_ctx *ctx, int n_bins, const int32_t *bin_ptr, const int32_t *bin_nodes, double r_inner);
int jf_get_curves(jf_ctx *ctx, double *curves);

/* accumulated device time (CUDA events on the context's stream) and work counters since creation
 * or the last jf_reset_stats */
typedef struct jf_stats {
    double ms_element;      /* element kernel                         */
    double ms_assembly;     /* stiffness + force gather + reductions  */
    double ms_cg;           /* CG kernel (includes the U update)      */
    double ms_setup;        /* host+device setup in jf_ctx_create     */
    int64_t tet_updates;    /* tets evaluated (active systems only)   */
    int64_t cg_iterations;  /* sum over systems                       */
    int64_t cg_launches;    /* CG kernel launches                     */
    int64_t kernel_launches;/* all kernels launched                   */
    int64_t newton_steps;   /* sum over systems                       */
    double ms_relax;        /* device time from the first to the last kernel of jf_relax calls */
    int64_t bytes_h2d;      /* host->device bytes copied (incl. setup uploads)  */
    int64_t bytes_d2h;      /* device->host bytes copied                        */
} jf_stats;
int jf_get_stats(jf_ctx *ctx, jf_stats *out);
int jf_reset_stats(jf_ctx *ctx);

/* structure facts for roofline arithmetic and tests */
typedef struct jf_info {
    int64_t n_nodes, n_free, n_tets, n_blocks /* non-zero 3x3 blocks */, n_slots /* incl. SELL padding */;
    int32_t n_sys, n_dirs /* directions after folding */, n_beams, sm_count, cg_grid, cg_block, elem_grid, elem_block;
    int64_t bytes_device;
    int32_t cg_resident /* 2: matrix in registers, 1: in shared memory, for the whole CG solve; 0: streamed */, cg_wmax;
    int32_t cg_stream_halo /* streaming kernel with per-group halo staging */;
    int32_t cg_fused /* one grid barrier per CG iteration: 1 register-resident kernel, 2 streaming kernel */;
} jf_info;
int jf_get_info(jf_ctx *ctx, jf_info *out);

/*
 * One-call form of what spherical_contraction_solver does between simulation.py:100 and :167
 * for a single load case: create, set, relax, read back, destroy.  U (n_nodes,3) is in/out
 * (start value / prescribed values in, relaxed displacements out); forces_out may be NULL.
 */
int jf_solve_boundarycondition(int device, int64_t n_nodes, const double *nodes, int64_t n_tets, const int32_t *tets,
                               const uint8_t *movable, const double *f_target, double *U, int n_beams,
                               const double *beams, double k, double d0, double lambda_s, double ds, double step,
                               int max_iter, double conv_crit, double *relrec, int32_t *cg_iters, int32_t *n_iter,
                               double *forces_out, jf_stats *stats_out);

/* FP64 FMA micro-benchmark (roofline denominator for the element kernel): returns GFLOP/s */
int jf_bench_fp64(int device, double *gflops);

/*
 * Mesh ingestion: read_meshfile (jointforces/simulation.py:25-78) without the Python line loop.  Host code only.
 * jf_msh_open reads the file, finds the first "$Nodes" / "$Elements" lines and the optional "$Jointforces" block
 * (has_radii = 1: r_inner, r_outer in the file's unit, micrometres; the caller scales by 10**-6 as :52-53 does);
 * jf_msh_read converts the node lines (tokens after the first -> coords (n_nodes,3)) and the element lines (LAST
 * four tokens, minus one -> tets (n_elements,4), as doubles like the reference returns them and/or as int32 like
 * jf_ctx_create takes them; either pointer may be NULL) on n_threads host threads (<= 0: all cores).  Numbers are
 * converted correctly rounded, i.e. bit-identical to Python's float().  Files the line loop would treat
 * differently from a plain decimal ASCII file return JF_ERR_FORMAT: the caller then runs the reference's loop,
 * which raises what the reference raises.
 */
typedef struct jf_msh jf_msh;
int jf_msh_open(const char *path, jf_msh **out, int64_t *n_nodes, int64_t *n_elements, int32_t *has_radii, double *r_inner,
                double *r_outer);
int jf_msh_read(jf_msh *msh, double *coords, double *tets_f64, int32_t *tets_i32, int32_t n_threads);
void jf_msh_close(jf_msh *msh);

/*
 * Lookup-table interpolants and pressure inference on the device (the consumer side of the table).
 *
 * jf_table_create: one scipy LinearNDInterpolator (jointforces/simulation.py:608 and :613) -- piecewise-linear
 * interpolation of `values` (n_points) over the Delaunay triangulation of `points` (n_points,2), NaN outside the
 * hull.  The caller passes the triangulation scipy.spatial.Delaunay(points) gives (what LinearNDInterpolator builds
 * internally): simplices (n_simplices,3) int32, transform (n_simplices,3,2), neighbors (n_simplices,3) int32.
 * jf_table_eval: out[i] = f(x_i, y_i) with optional log of either coordinate and exp of the result:
 *   get_displacement(distance, pressure) = f(log distance, log pressure)         (:610-611)  log_x=1 log_y=1 exp_out=0
 *   get_pressure(distance, displacement) = exp(f_inv(log distance, displacement)) (:615-616)  log_x=1 log_y=0 exp_out=1
 * jf_infer_pressure: force.py:399-432 for one frame of PIV vectors, NaN-coded like the reference; outputs are the
 * kept vectors in input order.  cos_min = cos(2 pi angle_filter / 360); use_filter = 0 for angle_filter=None.
 * jf_reconstruct_frames: the per-frame loop of reconstruct, force.py:70-121, for all frames of a series in one
 * call: stats (n_frames,76) = nanmean, nanmedian, nanstd(ddof=1), number of pressures used, 72 sector nanmeans.
 * r_min / r_max: NaN stands for None.  sector_bounds (72,2) = (alpha-5)*pi/180, (alpha+5)*pi/180 for
 * alpha = -180, -175 .. 175, as the caller computes them.
 * Agreement with scipy/numpy: to rounding (libm and CUDA log/exp/atan2 differ in the last bit; a query on a shared
 * edge may be assigned to the other triangle); medians are exact order statistics.
 */
typedef struct jf_table jf_table;
int jf_table_create(jf_table **out, int device, int64_t n_points, const double *points, const double *values, int64_t n_simplices,
                    const int32_t *simplices, const double *transform, const int32_t *neighbors);
void jf_table_destroy(jf_table *table);
int jf_table_eval(jf_table *table, int64_t n, const double *xq, const double *yq, int log_x, int log_y, int exp_out, double *out);
int jf_infer_pressure(jf_table *table, int64_t n, const double *x, const double *y, const double *u, const double *v, double cx,
                      double cy, double r_sph, int use_filter, double cos_min, double *distance, double *displacement,
                      double *angle, double *pressure, int64_t *n_out);
int jf_reconstruct_frames(jf_table *table, int32_t n_frames, const int64_t *frame_ptr, const double *frames /* (n_frames,3): cx, cy, r */,
                          const double *x, const double *y, const double *u, const double *v, int use_filter, double cos_min,
                          double r_min, double r_max, const double *sector_bounds, double *stats);
int64_t jf_table_launches(jf_table *table); /* kernels launched so far (measurement) */

#ifdef __cplusplus
}
#endif
#endif /* JF_B200_H */
