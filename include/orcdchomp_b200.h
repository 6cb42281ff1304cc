/* orcdchomp_b200.h -- C ABI of the B200-native CHOMP engine.
 *
 * This is the drop-in boundary for the hot path of personalrobotics/or_cdchomp:
 *   - the batched CHOMP iteration that replaces the cd_chomp_iterate + callback
 *     ping-pong of  src/orcdchomp_mod.cpp:2752-2830  (mod::iterate loop),
 *     src/libcd/chomp.c:430-683 (cd_chomp_iterate) and the two OpenRAVE-bound
 *     callbacks sphere_cost_pre / sphere_cost (src/orcdchomp_mod.cpp:968-1327);
 *   - the signed-distance-field build  cd_grid_double_bin_sdf
 *     (src/libcd/grid.c:637-687), the flood fill + relabel of
 *     src/orcdchomp_mod.cpp:543-548 / src/libcd/grid_flood.c:30-111 and the
 *     occupancy loop of src/orcdchomp_mod.cpp:498-525.
 *
 * Plain C types only: pointers, sizes, integer error codes.  No exceptions cross
 * this boundary, no torch / CUDA types appear in a signature (streams and device
 * pointers travel as void*).  All host buffers are caller-owned.  One host thread
 * per engine handle.  Every entry point FAILS (negative code) when no CUDA device
 * is usable -- there is no CPU fallback behind this header.
 *
 * (All file:line citations are relative to the reference repository root.)
 */
#ifndef ORCDCHOMP_B200_H
#define ORCDCHOMP_B200_H

#ifndef __CUDACC_RTC__ /* the kernels are also compiled at run time (NVRTC), without host headers */
#include <stddef.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* error codes (libcd uses -1 alloc / -2 bad type, chomp.c:47,399; grid.c:648) */
#define OCB_OK              0
#define OCB_ERR_ALLOC      (-1)   /* host or device allocation failed                 */
#define OCB_ERR_ARG        (-2)   /* bad argument / unsupported configuration         */
#define OCB_ERR_CUDA       (-3)   /* CUDA runtime error (see ocb_last_error)          */
#define OCB_ERR_NODEVICE   (-4)   /* no usable sm_100 device: the engine refuses to run */
#define OCB_ERR_JLIMIT     (-5)   /* a run left joint limits (chomp.c:651-655)        */

const char *ocb_last_error(void);          /* thread-local, never NULL */
const char *ocb_version(void);

/* ------------------------------------------------------------------------- */
/* Robot description: what mod::create pulls out of OpenRAVE once per run
 * (src/orcdchomp_mod.cpp:2104-2134 dof bookkeeping, 2148-2300 spheres,
 * 2639-2660 limits) plus the kinematic tree that OpenRAVE's SetActiveDOFValues /
 * Link::GetTransform / CalculateJacobian evaluate inside sphere_cost_pre
 * (src/orcdchomp_mod.cpp:1022-1049).  Poses are libcd poses [x y z qx qy qz qw]
 * (src/libcd/kin.c:42-52).
 *
 *   T_world_link[i] = T_world_link[parent[i]] o pose_parent[i] o motion_i(value_i)
 *   value_i         = dof_coeff[i][0] * q[dof_index[i]] + dof_coeff[i][1]
 *                     (dof_index[i] < 0: frozen joint, value_i = dof_coeff[i][1])
 *   motion_i        = rotation by value_i about axis[i]   (OCB_JOINT_REVOLUTE)
 *                     translation by value_i along axis[i] (OCB_JOINT_PRISMATIC)
 *                     identity                             (OCB_JOINT_FIXED)
 * axis[i] is a unit vector in link i's frame through that frame's origin.
 */
#define OCB_JOINT_FIXED     0
#define OCB_JOINT_REVOLUTE  1
#define OCB_JOINT_PRISMATIC 2

typedef struct ocb_robot
{
   int n_links;                 /* link 0 is the root; parent[i] < i                  */
   const int *parent;           /* [n_links], parent[0] = -1                          */
   const double *pose_parent;   /* [n_links][7]                                       */
   const int *joint_type;       /* [n_links]                                          */
   const double *axis;          /* [n_links][3]                                       */
   const int *dof_index;        /* [n_links] index into the active-dof vector or -1   */
   const double *dof_coeff;     /* [n_links][2]                                       */
   double base_pose[7];         /* world pose of link 0 (robot->GetTransform())       */
   int n_dof;                   /* n: number of active dofs (GetActiveDOF)            */
   const double *limit_lower;   /* [n_dof] (GetDOFLimits, mod.cpp:2639-2660)          */
   const double *limit_upper;   /* [n_dof]                                            */
   int n_spheres;               /* sphere table in <orcdchomp><spheres> XML order     */
   const int *sphere_link;      /* [n_spheres] link index (run_sphere.robot_linkindex) */
   const double *sphere_pos;    /* [n_spheres][3] pos_wrt_link                        */
   const double *sphere_radius; /* [n_spheres]                                        */
} ocb_robot;

/* One rooted signed distance field (struct run_rsdf, mod.cpp:850-855): the grid
 * (struct cd_grid, src/libcd/grid.h:29-41; C order grid[x*NY*NZ+y*NZ+z]) and
 * the world pose of its frame snapshotted at create (mod.cpp:2347-2369). */
typedef struct ocb_sdf
{
   int sizes[3];
   double lengths[3];
   double pose_world_gsdf[7];
   const double *data;          /* host pointer, sizes[0]*sizes[1]*sizes[2] doubles   */
} ocb_sdf;

/* One hard end-effector constraint: a task space region whose zero-width axes are held at zero
 * (the `create` command's con_tsr / everyn_tsr / start_tsr arguments, src/orcdchomp_mod.cpp:1930-1996;
 * callbacks con_tsr / con_everyn_tsr / con_start_tsr 1330-1784; projection src/libcd/chomp.c:553-600).
 * The constrained frame is  robot link `link` * pose_link_ee  (a manipulator's end effector and its
 * local tool transform, or a bare link with the identity).  Its pose seen from the TSR's frame,
 * inv(T0w) * frame * inv(Twe), is written as [x y z roll pitch yaw]; entry i of that vector is held at
 * zero when Bw[i][0] == 0 && Bw[i][1] == 0 (mod.cpp:2466-2519); the other bounds are not used. */
#define OCB_CON_START 0   /* con_tsr 'start ...': first moving waypoint                      */
#define OCB_CON_END   1   /* con_tsr 'end ...'  : last moving waypoint                       */
#define OCB_CON_ALL   2   /* con_tsr 'all ...' and everyn_tsr: every moving waypoint         */
#define OCB_CON_START_TSR 3 /* start_tsr: the start point itself joins the optimised rows
                               (m = n_points-1, no initial boundary row, mod.cpp:2316, 2571-2578)
                               and carries the constraint; at most one, not with floating_base */
typedef struct ocb_constraint
{
   int where;                   /* OCB_CON_*                                            */
   int link;                    /* robot link index of the constrained frame            */
   double pose_link_ee[7];      /* frame in the link: [x y z qx qy qz qw]                */
   double T0w[7];               /* tsr->T0w (parsed by tsr_create_parse, mod.cpp:3068)   */
   double Twe[7];               /* tsr->Twe                                              */
   double Bw[6][2];             /* bounds, rows x y z roll pitch yaw                     */
} ocb_constraint;

/* Per-batch parameters = the `create` command's numeric arguments and defaults
 * (src/orcdchomp_mod.cpp:1818-1848, 1875, 1888-2079). */
typedef struct ocb_params
{
   int n_points;                /* default 101; m = n_points-2 moving waypoints       */
   int derivative;              /* D of cd_chomp_create, default 1                    */
   double lambda;               /* default 10, must be >= 0.01                        */
   int use_momentum;
   int use_hmc;                 /* 1 = on; 2 = on, momentum drawn by the serial generator (test hook, same stream) */
   double hmc_resample_lambda;  /* default 0.02                                       */
   double epsilon;              /* default 0.1                                        */
   double epsilon_self;         /* default 0.04                                       */
   double obs_factor;           /* default 200                                        */
   double obs_factor_self;      /* default 10                                         */
   int floating_base;           /* `floating_base` (mod.cpp:991-1021, 1050-1086, 2424-2443): every waypoint
                                   starts with the base pose [x y z qx qy qz qw], n = 7 + robot->n_dof,
                                   q_start / q_goal rows are that long, robot->base_pose is not used, all
                                   spheres are active, pose columns of the Jacobian are scaled by 0.01,
                                   quaternions are re-normalised after every iteration              */
   int n_constraints;           /* hard constraints, in the order the reference adds them:
                                   start_tsr, everyn_tsr, then the con_tsr list (mod.cpp:2571-2613) */
   const ocb_constraint *constraints;
} ocb_params;

void ocb_params_default(ocb_params *p);

/* ------------------------------------------------------------------------- */
/* Engine: one per GPU (owns a stream, the resident SDFs and scratch).        */
typedef struct ocb_engine ocb_engine;
typedef struct ocb_batch ocb_batch;

int ocb_engine_create(int device, ocb_engine **out);
int ocb_engine_destroy(ocb_engine *e);
/* Make the engine launch on a caller stream (cudaStream_t as void*; NULL = own). */
int ocb_engine_set_stream(ocb_engine *e, void *cuda_stream);
int ocb_engine_sync(ocb_engine *e);
/* The engine keeps the device memory of destroyed batches and removed fields in its own pool so
 * that the next create costs microseconds; this hands the unused part back to the driver. */
int ocb_engine_trim(ocb_engine *e);
/* Run-time specialisation: batches created while this is on get the persistent kernel compiled
 * (NVRTC, once per configuration, ~2 s, cached for the life of the process) with their sizes --
 * waypoints, dofs, spheres, joint frames, fields, mode flags -- as literal constants.  Same
 * source and arithmetic as the library's own kernel (results agree to ~1e-15), fewer integer
 * instructions.  Off by
 * default (environment variable OCB_JIT=1 turns it on for every engine); if NVRTC is missing
 * the library's kernel is used and ocb_last_error() says why. */
int ocb_engine_enable_jit(ocb_engine *e, int on);
/* Inspection hook (host only, works without a GPU): the tables the run-time compiler is given for this
 * robot -- its kinematic tree and static self-collision pair list as constexpr data, which
 * csrc/chomp_jit_robot.cuh turns into straight-line code.  Returns the length of the text (0: this
 * robot takes the table-driven kernel), copies at most cap-1 bytes into buf (may be NULL). */
long ocb_debug_jit_robot_header(const ocb_robot *robot, const ocb_params *params, char *buf, size_t cap);
/* The smoothness metric of a run as the engine builds it (host only, no device needed; for inspection and
 * tests): band of A [m][2 derivative + 1], its dense inverse [m][m], the coefficient vectors of B
 * (B = bi (x) q_start + bf (x) q_goal) [m] each, closed_form bit 0: A = c tridiag(-1, 2, -1) (the solve is the
 * product with the closed-form inverse), bit 1: B = -c (q_start e_1 + q_goal e_m) (the smoothness part of the
 * update is T minus the straight line), c.  m = n_points - 2, or n_points - 1 with free_start (start_tsr).
 * Any output may be NULL. */
int ocb_debug_metric(int n_points, int derivative, int free_start, double *Aband, double *Ainv, double *bi,
                     double *bf, int *closed_form, double *c);

/* --- SDF residency (replaces mod::sdfs[], mod.cpp:584-586 / 716-718 / 836) --- */
/* copies the grid to HBM; *id indexes it in later calls */
int ocb_sdf_upload(ocb_engine *e, const ocb_sdf *sdf, int *id);
/* adopt an SDF that already lives in HBM (e.g. the output of ocb_sdf_build_device) */
int ocb_sdf_adopt_device(ocb_engine *e, const int sizes[3], const double lengths[3],
                         const double pose_world_gsdf[7], const double *d_data, int *id);
int ocb_sdf_remove(ocb_engine *e, int id);
/* cd_grid_double_interp + cd_grid_double_grad (grid.c:331-454) of resident field `id` at k points
 * given in the GRID frame [k][3]: the device function the CHOMP kernels sample with.  errs[i] = 1
 * where cd_grid_lookup_index rejects the point (grid.c:191-209; values / grads are then 0), else 0;
 * values[i] may be HUGE_VAL (grid.c:402, 431, 441). */
int ocb_sdf_sample_host(ocb_engine *e, int id, const double *points, int k, double *values,
                        double *grads, int *errs);

/* --- SDF build: cd_grid_double_bin_sdf (grid.c:637-687) --------------------- *
 * obs: 0.0 = free, HUGE_VAL = obstacle (any non-zero finite value is treated as
 * the parabola height grid.c:274-304 gives it).  sdf: > 0 in free space, < 0
 * inside obstacles.  _host copies in/out; _device works on HBM pointers.      */
int ocb_sdf_build_host(ocb_engine *e, const double *obs, const int sizes[3],
                       const double lengths[3], double *sdf);
int ocb_sdf_build_device(ocb_engine *e, const double *d_obs, const int sizes[3],
                         const double lengths[3], double *d_sdf);
/* ocb_sdf_build_* picks an exact integer path when the grid holds only 0 / HUGE_VAL and
 * its cells are cubes (what computedistancefield produces), else the general fp64 path
 * that follows grid.c:462-569 operation for operation.  Test hook: force the latter. */
int ocb_engine_force_general_sdf(ocb_engine *e, int on);
/* squared Euclidean distance transform alone: cd_grid_double_dt_sqeuc (grid.c:462-569) */
int ocb_dt_sqeuc_device(ocb_engine *e, const double *d_func, const int sizes[3],
                        const double lengths[3], double *d_out);
int ocb_dt_sqeuc_host(ocb_engine *e, const double *func, const int sizes[3],
                      const double lengths[3], double *out);

/* --- occupancy pipeline of computedistancefield (mod.cpp:386-410, 498-548) --- */
/* Analytic stand-in for the per-voxel OpenRAVE CheckCollision(cube): a voxel is
 * an obstacle iff the axis-aligned cube of half-extent cube_extent centred at
 * the voxel centre (grid frame) overlaps any primitive.  Primitives are given
 * in the GRID frame.  OCB_PRIM_BOX: pose[7] + half extents (oriented box, SAT
 * test); OCB_PRIM_SPHERE: centre + radius; OCB_PRIM_TRIANGLE: three vertices (triangle
 * meshes, the usual OpenRAVE geometry).  d_grid receives 1.0 (free) or
 * HUGE_VAL (obstacle) exactly as mod.cpp:398,522 leave it.                     */
#define OCB_PRIM_BOX      0
#define OCB_PRIM_SPHERE   1
#define OCB_PRIM_TRIANGLE 2   /* one triangle of a mesh: v0 = pose[0..2], v1 = pose[3..5],
                                 v2 = (pose[6], extents[0], extents[1]); cube-vs-triangle by the
                                 13-axis separating-axis test (touching counts as a hit) */
typedef struct ocb_prim
{
   int type;
   double pose[7];              /* box: pose in grid frame; sphere: pose[0..2] = centre */
   double extents[3];           /* box: half extents; sphere: extents[0] = radius       */
} ocb_prim;
int ocb_occupancy_device(ocb_engine *e, const ocb_prim *prims, int n_prims,
                         const int sizes[3], const double lengths[3], double cube_extent,
                         double *d_grid);
/* cd_grid_flood_fill(g, 0, no wrap, replace_1_to_0) followed by the 1.0 -> HUGE_VAL
 * relabel (grid_flood.c:30-111, mod.cpp:143-151, 543-548): 6-connected fill of
 * the value 1.0 with 0.0 starting at index_start, then every remaining 1.0
 * becomes HUGE_VAL.  Works in place on an HBM grid.                            */
int ocb_flood_relabel_device(ocb_engine *e, double *d_grid, const int sizes[3],
                             size_t index_start);
int ocb_flood_relabel_host(ocb_engine *e, double *grid, const int sizes[3],
                           size_t index_start);
/* whole pipeline with host buffers (what `computedistancefield` does after the
 * AABB sizing): occupancy -> flood/relabel -> SDF.  obs_out may be NULL.       */
int ocb_computedistancefield_host(ocb_engine *e, const ocb_prim *prims, int n_prims,
                                  const int sizes[3], const double lengths[3],
                                  double cube_extent, double *obs_out, double *sdf_out);

/* Device-resident variants: the field is built in HBM and stays there as SDF `*id` (no host
 * round trip); ocb_sdf_build_resident is cd_grid_double_bin_sdf of a host obstacle array
 * (addfield_fromobsarray, mod.cpp:592-722).  ocb_sdf_download copies a resident field out
 * (cache files, mod.cpp:571-580).  ocb_sdf_alias makes a second id over the SAME grid with
 * another world pose -- the per-run snapshot of a rooted field's pose (mod.cpp:2347-2369)
 * without copying the grid; remove aliases before the field they point at. */
int ocb_computedistancefield_resident(ocb_engine *e, const ocb_prim *prims, int n_prims,
                                      const int sizes[3], const double lengths[3], double cube_extent,
                                      const double pose_world_gsdf[7], int *id);
int ocb_sdf_build_resident(ocb_engine *e, const double *obs, const int sizes[3], const double lengths[3],
                           const double pose_world_gsdf[7], int *id);
int ocb_sdf_download(ocb_engine *e, int id, double *out);
int ocb_sdf_alias(ocb_engine *e, int id, const double pose_world_gsdf[7], int *alias_id);

/* --- batched CHOMP runs (replaces struct run + cd_chomp, mod.cpp:887-966) ---- */
/* R independent runs sharing robot, parameters and SDF set.  q_start/q_goal are
 * [R][n_dof]; seeds [R] (gsl_rng_set seed, mod.cpp:2303-2304; may be NULL = 0).
 * The straight-line initial trajectory is mod.cpp:2456-2458.                  */
int ocb_batch_create(ocb_engine *e, const ocb_robot *robot, const ocb_params *params,
                     int n_sdfs, const int *sdf_ids, int n_runs,
                     const double *q_start, const double *q_goal,
                     const unsigned int *seeds, ocb_batch **out);
/* Re-arm an existing batch as if it had just been created: new end points (host
 * pointers, [R][n_dof]; NULL = keep the resident ones), straight-line
 * trajectories, zero momentum, fresh rng (seeds NULL = keep the previous seeds).
 * Equivalent to destroy + create with the same robot / parameters / SDFs, without
 * re-allocating.  Asynchronous on the engine stream. */
int ocb_batch_reset(ocb_batch *b, const double *q_start, const double *q_goal,
                    const unsigned int *seeds);
/* starttraj variant (mod.cpp:2373-2415): traj is [R][n_points][n_dof] */
int ocb_batch_set_traj(ocb_batch *b, const double *traj);
/* n_iter CHOMP iterations per run (mod.cpp:2752-2828) followed by the cost-only
 * pass (mod.cpp:2830).  Outputs are [R], each may be NULL.  cost_* are
 * chomp.c:679-681 of the FINAL cost-only pass; status[r] is 0 or OCB_ERR_JLIMIT.
 * Returns OCB_OK even if individual runs hit OCB_ERR_JLIMIT.                  */
int ocb_batch_iterate(ocb_batch *b, int n_iter, double *cost_total, double *cost_obs,
                      double *cost_smooth, int *status);
/* asynchronous form: enqueue only (results stay on the device) */
int ocb_batch_iterate_async(ocb_batch *b, int n_iter);
/* One `iterate` command split over several calls (max_time, per-iteration trajectory dumps:
 * mod.cpp:2752-2828 runs r->iter = 0..n_iter-1 inside ONE command): the iterations of this call
 * are numbered first_iter, first_iter+1, ... for the HMC schedule (resample when r->iter ==
 * hmc_resample_iter, alpha = 100 exp(0.02 r->iter), mod.cpp:2755-2768).  ocb_batch_iterate is
 * first_iter = 0. */
int ocb_batch_iterate_from(ocb_batch *b, int first_iter, int n_iter, double *cost_total, double *cost_obs,
                           double *cost_smooth, int *status);
int ocb_batch_iterate_from_async(ocb_batch *b, int first_iter, int n_iter);
int ocb_batch_get_costs(ocb_batch *b, double *cost_total, double *cost_obs,
                        double *cost_smooth, int *status);
/* the public cd_chomp fields a caller may poke between iterations (chomp.h:40-41, 49-50,
 * 95-96): the momentum matrix AG [R][m][n] with its leapfrog_first flags [R] -- the
 * reference's module resamples AG itself for HMC (mod.cpp:2755-2768) -- and lambda. */
int ocb_batch_set_momentum(ocb_batch *b, const double *AG, const int *leapfrog_first);
int ocb_batch_get_momentum(ocb_batch *b, double *AG, int *leapfrog_first);
int ocb_batch_set_lambda(ocb_batch *b, double lambda);
/* iterations each run completed in the last iterate call: n_iter, or fewer for a run that left
 * the joint limits (the reference's r->iter when the exception is thrown, mod.cpp:2799-2803) */
int ocb_batch_get_iterations(ocb_batch *b, int *iterations);
/* the most projection steps of the joint-limit loop (chomp.c:608-655, at most 1000) any single iteration
 * of the last iterate call needed, per run: 0 = the run never left its limits.  The loop is the one
 * part of the algorithm that is not continuous in its own rounding, so this is the
 * diagnostic that tells a well-conditioned run from a chaotic one. */
int ocb_batch_get_limit_rounds(ocb_batch *b, int *rounds);
/* Hard constraints: how many waypoint systems (tridiagonal metric) or whole systems (wider metrics) of each
 * run were singular since the batch was created -- linearly dependent constraint rows; the reference prints
 * "constraint inversion error!" (chomp.c:582-590) and carries on, the engine skips those rows.  skips [n_runs]. */
int ocb_batch_get_constraint_skips(ocb_batch *b, int *skips);
/* per-iteration cost log of the last iterate call: [R][n_iter][3] (total, obs,
 * smooth) as RAVELOG_INFO prints them (mod.cpp:2798).  Enable before iterate. */
int ocb_batch_enable_trace(ocb_batch *b, int enable);
int ocb_batch_get_trace(ocb_batch *b, double *trace, int n_iter);
/* gettraj (numeric part of mod.cpp:2897-2903): [R][n_points][n_dof] */
int ocb_batch_get_traj(ocb_batch *b, double *traj);
/* Parity hook: capture, during subsequent iterate calls, the gradient of the most
 * recent iteration, [R][m][n_dof].  mode 1: G after chomp.c:515-522 (obstacle
 * gradient / m + A T + B); mode 2: the obstacle/self-collision part alone (sum of
 * the c_grad rows of sphere_cost scaled by 1/m, chomp.c:474-492); mode 0: off. */
int ocb_batch_capture_gradient(ocb_batch *b, int mode);
int ocb_batch_get_gradient(ocb_batch *b, double *G);
/* arg-min of cost_total over this batch's runs after iterate (device reduction) */
int ocb_batch_best(ocb_batch *b, int *best_run, double *best_cost);
int ocb_batch_destroy(ocb_batch *b);
/* sizes for callers that allocate outputs */
int ocb_batch_uses_jit(const ocb_batch *b);   /* 1 when this batch runs a run-time specialised kernel */
int ocb_batch_tile_width(const ocb_batch *b); /* waypoints per tile of the tiled large-robot path (32, 16, 8); 0 = persistent kernel */
int ocb_batch_dims(const ocb_batch *b, int *n_runs, int *n_points, int *n_dof);
/* device pointers (HBM) of the trajectory [R][n_points][n_dof] and costs [R][3],
 * for callers that keep everything resident (bench, NCCL gather)              */
int ocb_batch_device_ptrs(ocb_batch *b, void **d_traj, void **d_costs);
/* copy one run's trajectory [n_points][n_dof] to another HBM buffer (async, D2D) */
int ocb_batch_copy_run_traj_device(ocb_batch *b, int run, void *d_dst);
/* kernel launches issued by this engine since creation (bench bookkeeping)    */
long ocb_engine_launch_count(const ocb_engine *e);

/* ------------------------------------------------------------------------- */
/* Several GPUs in one process: G engines, one host thread each, driven in lock
 * step.  Runs are dealt round-robin (global run r lives on device slot r mod G), fields are
 * replicated, nothing is exchanged during the iterations; ocb_multi_batch_best is the one exchange
 * (arg-min of cost_total + the winner's trajectory, over NCCL when libnccl.so.2 is present and the
 * devices are distinct, else compared on the host and copied device to device).  Outputs are in
 * global run order, so results do not depend on G.  The reference has no counterpart: its module
 * advances runs one after the other (mod.cpp:2752-2828). */
typedef struct ocb_multi ocb_multi;
typedef struct ocb_multi_batch ocb_multi_batch;
const char *ocb_multi_last_error(void);
/* devices: [n_devices] CUDA ordinals (NULL = 0..n-1); an ordinal may repeat (several engines on one GPU) */
int ocb_multi_create(int n_devices, const int *devices, ocb_multi **out);
int ocb_multi_destroy(ocb_multi *m);
int ocb_multi_device_count(const ocb_multi *m);
int ocb_multi_uses_nccl(const ocb_multi *m);
int ocb_multi_engine(ocb_multi *m, int slot, ocb_engine **e);
int ocb_multi_enable_jit(ocb_multi *m, int on);
int ocb_multi_sdf_upload(ocb_multi *m, const ocb_sdf *sdf, int *id);
int ocb_multi_computedistancefield_resident(ocb_multi *m, const ocb_prim *prims, int n_prims, const int sizes[3],
                                            const double lengths[3], double cube_extent,
                                            const double pose_world_gsdf[7], int *id);
int ocb_multi_sdf_remove(ocb_multi *m, int id);
/* arguments as ocb_batch_create, [n_runs] rows in global run order */
int ocb_multi_batch_create(ocb_multi *m, const ocb_robot *robot, const ocb_params *params, int n_sdfs,
                           const int *sdf_ids, int n_runs, const double *q_start, const double *q_goal,
                           const unsigned int *seeds, ocb_multi_batch **out);
int ocb_multi_batch_dims(const ocb_multi_batch *b, int *n_runs, int *n_points, int *n_dof);
int ocb_multi_batch_iterate(ocb_multi_batch *b, int n_iter, double *cost_total, double *cost_obs,
                            double *cost_smooth, int *status);
int ocb_multi_batch_get_traj(ocb_multi_batch *b, double *traj);
/* best_run = global id or -1 when every run failed; traj: host [n_points][n_dof] or NULL */
int ocb_multi_batch_best(ocb_multi_batch *b, int *best_run, double *best_cost, double *traj);
int ocb_multi_best_traj_device(ocb_multi *m, int slot, void **d_traj);
int ocb_multi_batch_destroy(ocb_multi_batch *b);

#ifdef __cplusplus
}
#endif
#endif /* ORCDCHOMP_B200_H */
