/* ocb_internal.h -- types shared by the host engine and the sm_100a kernels.
 * Not part of the public ABI (that is include/orcdchomp_b200.h). */
#ifndef OCB_INTERNAL_H
#define OCB_INTERNAL_H

#ifdef __CUDACC_RTC__
/* run-time compilation of the CHOMP kernel (ocb_jit.cpp): no host headers */
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#else
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#endif

#define OCB_MAX_JOINTS 24   /* moving joints carried in kernel-parameter (constant) space */
#define OCB_MAX_SDFS 64  /* descriptors are staged in shared memory */
#define OCB_MAX_BW 8        /* half bandwidth of the smoothness metric (= derivative D) */
#define OCB_INLINE_SDFS 2    /* field descriptors carried in the kernel parameters themselves */

/* parent-transform source of a joint in the forward sweep */
#define OCB_LOAD_PREV (-1)
#define OCB_LOAD_BASE (-2)

/* One moving joint of the compiled kinematic tree.  The joint frame has its axis
 * on local z; X = fixed transform from the parent joint frame (or world, for
 * root-level joints; base pose folded in) to this joint's frame at value 0. */
struct OcbJointDev
{
   double XR[9];
   double Xt[3];
   double c0, c1;      /* joint value = c0 * q[dof] + c1 */
   int type;           /* OCB_JOINT_REVOLUTE / OCB_JOINT_PRISMATIC */
   int dof;
   int load;           /* OCB_LOAD_PREV / OCB_LOAD_BASE / slot index */
   int save;           /* slot index to save this joint's transform into, or -1 */
   int sph_begin, sph_end;   /* active spheres rigidly attached to this joint frame */
   int desc_begin, desc_end; /* sphere groups in this joint's subtree (incl. its own), in the desc table */
};

struct OcbSphereDev
{
   double pos[3];      /* in the joint frame */
   double radius;
   int link;           /* original robot link index (same-link pairs are skipped) */
   int group;          /* compact index of the joint frame among those that carry spheres */
};

struct OcbSdfDev
{
   const double *data;
   int size[3];
   int pad;
   double length[3];
   double scale[3];    /* size / length */
   double cell[3];     /* length / size */
   double Rgw[9];      /* rotation of pose_gsdf_world (world -> grid), row-major */
   double tgw[3];
   double Rwg[9];      /* rotation of pose_world_gsdf (grid -> world) */
   /* sdf_sample's fast path (y = p * scale in cells, u = y - floor(y) - 0.5): y > edge_hi is outside
    * beyond rounding; |u| < near (centre plane) or |u| > near_hi (cell face, ends of the axis) takes
    * the exact path */
   double edge_hi[3], near[3], near_hi[3];
};

/* One hard constraint (ocb_constraint) compiled for the kernel: the frame it holds is joint frame
 * `joint` times C (link offset, tool transform and inv(Twe) folded together), seen through A = inv(T0w). */
struct OcbConDev
{
   double AR[9], At[3];
   double CR[9], Ct[3];
   unsigned long long anc; /* bit j: joint j moves the frame */
   int joint;              /* -1: the frame is fixed in the world (C is then its world pose times inv(Twe)) */
   int where;              /* OCB_CON_* */
   int k;                  /* rows */
   int rows[6];            /* entry of [x y z yaw pitch roll] behind each row */
   int pad;
};

struct OcbChompArgs
{
   /* sizes */
   int R, P, m, n;
   int nj, nsa, nsi, nsdf;
   int bw, n_slots, Ppad, n_iter;
   int use_momentum, use_hmc, trace_on, grad_mode; /* grad_mode 0 none, 1 full G, 2 obstacle only */
   int tiled, ng;          /* tiled: large-robot path (chomp_tiled.cu); ng: joint frames that carry spheres */
   int n_desc, NAp;        /* NAp: padded active part of a cut2 row (>= nsa + 3); row = NAp + nsi */
   int tile_w, n_tiles;    /* tiled path: waypoints per tile (32, 16 or 8), tiles per run */
   int floating, iter_base; /* floating base: rows start with the base pose, joint 0 is the base frame;
                              iter_base: the reference's r->iter at the first iteration of this launch (HMC schedule) */
   /* robot */
   OcbJointDev joints[OCB_MAX_JOINTS];
   const OcbSphereDev *spheres;
   const int *desc;        /* [n_desc] group indices, see OcbJointDev::desc_begin */
   const int *ganc;        /* inverse of desc: [ng + 1] offsets into this same array, then the joints
                              whose subtree carries group g (tiled path) */
   const double *gbound;   /* [nj][4]: bounding sphere (centre in the joint frame, radius) of the spheres a joint
                              frame carries, radius < 0 when it carries none (tiled path: partner culling) */
   const double *inactive_pos;
   /* self-collision tables over NS = nsa + nsi spheres (active first):
    * cut2[s][o] = (r_s + r_o + epsilon_self)^2, or -1 when s and o sit on the same link
    * (mod.cpp:1256) so the pair never passes the range test.  Row layout: NAp entries for
    * active partners (padded with -1), then nsi entries for inactive ones.  radius[o]. */
   const double *cut2;
   const double *radius;
   const OcbSdfDev *sdfs; /* [nsdf] in HBM, staged to shared memory by the kernel */
   /* metric: A band [m][2bw+1], LDL^T factors, B coefficient vectors */
   const double *Aband;
   const double *Lband;   /* [m][bw] */
   const double *dinv;    /* [m]     */
   const double *bcoef_i; /* [m]     */
   const double *bcoef_f; /* [m]     */
   double trc_ss, trc_sg, trc_gg;
   /* band_toeplitz: every row of A holds the same band (true for derivative = 1: A = (m+1) tridiag(-1, 2, -1));
    * band_row is that band, read as instruction operands instead of m x (2 bw + 1) table loads */
   int band_toeplitz, band_121; /* band_121: A = c tridiag(-1, 2, -1) exactly -> closed-form inverse (band_solve_121_scan) */
   double band_row[2 * OCB_MAX_BW + 1];
   double band_121_scale;       /* 1 / ((m + 1) c) */
   /* parameters */
   double lambda, dt, eps, eps_self, obs_factor, obs_factor_self, hmc_lambda;
   const double *lim_lo;
   const double *lim_hi;
   /* per-run state in HBM */
   double *traj;          /* [R][P][n] */
   double *AG;            /* [R][m][n] */
   int *leapfrog_first;   /* [R] */
   int *hmc_next;         /* [R] */
   uint32_t *mt_state;    /* [R][625] */
   double *costs;         /* [R][3] */
   int *status;           /* [R] */
   int *iters_done;       /* [R] iterations completed by the last iterate call */
   int *limit_rounds;     /* [R] most joint-limit projection steps (chomp.c:608-655) one iteration of the last call took */
   double *trace;         /* [R][n_iter][3] */
   double *grad_out;      /* [R][m][n] */
   double *G_obs;         /* [R][m][n] obstacle + self-collision gradient, unscaled (tiled path) */
   double *tile_cost;     /* [R][n_tiles] cost partials (tiled path) */
   double *trig_cache;    /* [R][2 nj][Ppad]: sine / cosine of every joint angle, written by the forward sweep and read
                             back by the J^T sweep of the same iteration (compiled-robot kernel; stays in L2) */
   size_t ws_stride;      /* doubles of per-run workspace in shared memory (persistent kernel) */
   int robot_smem;        /* 1: kernel compiled with the robot as code (OCB_JIT_ROBOT) -- no pair / subtree tables in
                             shared memory, the tridiagonal factor staged there instead */
   int pad1;
   /* the first fields' descriptors by value: kernel parameters live in the constant bank, so a kernel
    * compiled for <= OCB_INLINE_SDFS fields reads them as instruction operands (no loads, no registers) */
   OcbSdfDev sdf_inline[OCB_INLINE_SDFS];
   /* hard constraints (chomp.c:553-600): con_K stacked rows, waypoint-major; library kernel only */
   int n_con, con_K;
   int free_start;            /* start_tsr: P - 1 moving waypoints, the first is the start point (see chomp_iterate_body) */
   int con_kmax;              /* most rows on one waypoint */
   int con_kuniform, pad4;    /* the number of rows when every moving waypoint carries the same (> 0), else 0 */
   /* con_fast: tridiagonal metric -> con_project_tridiag instead of the dense system.  Where its operands live:
    * con_jh_smem: J and h at the start of the run's shared workspace (over the sphere centres, which the
    * next forward sweep rewrites); con_rec_smem: the sweep's matrices behind them at con_rec_off; else in
    * con_scratch */
   int con_fast, con_jh_smem, con_rec_smem, con_rec_off;
   const OcbConDev *cons;     /* [n_con] */
   const int *con_row0;       /* [m + 1]: first row of each moving waypoint */
   const int *con_row_wp;     /* [con_K]: moving waypoint (0-based) of each row */
   const double *Ainv;        /* [m][m] dense inverse of the metric (only the entries between constrained waypoints are read) */
   double *con_scratch;       /* [R][con_stride]: J (con_K x n), h, saved h, S (con_K x con_K) */
   size_t con_stride;
   size_t con_slots_off;      /* tiled path: where, in a run's scratch, the constraint evaluation keeps its branch frames */
   int *con_singular;         /* [R] iterations whose constraint system had a zero pivot (the reference prints and goes on) */
};

/* doubles of scratch con_project_tridiag (chomp_constraints.cuh) needs */
static inline size_t ocb_con_tridiag_scratch(int m, int n, int kmax)
{
   /* N and v of every waypoint, two work areas (one per warp), one word for the second warp's count */
   return (size_t) m * n * n + (size_t) m * n + 2 * (2 * (size_t) n * n + 2 * (size_t) n * kmax + 2 * (size_t) kmax * kmax + n) + 1;
}

#if !defined(__CUDACC_RTC__) && defined(__cplusplus)
#include <mutex>
/* cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) property of a kernel:
 * engines on several GPUs in one process each need their own opt-in.  Remembers, per device,
 * the largest size a kernel has been configured for; safe across engine threads. */
struct OcbSmemOptIn
{
   std::mutex mu;
   size_t configured[64] = {0};
   template <class K>
   cudaError_t ensure(K kernel, size_t bytes)
   {
      int dev = 0;
      cudaError_t e = cudaGetDevice(&dev);
      if (e != cudaSuccess) return e;
      std::lock_guard<std::mutex> lock(mu);
      if (dev < 0 || dev >= 64 || bytes > configured[dev])
      {
         e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
         if (e != cudaSuccess) return e;
         if (dev >= 0 && dev < 64) configured[dev] = bytes;
      }
      return cudaSuccess;
   }
};
#endif

#ifndef __CUDACC_RTC__
#ifdef __cplusplus
extern "C" {
#endif
/* kernels' host launchers (defined in the .cu files) */
cudaError_t ocb_launch_chomp(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st);
size_t ocb_chomp_smem_bytes(const OcbChompArgs *args);
/* ocb_jit.cpp: the persistent kernel compiled at run time for one batch's sizes */
int ocb_jit_chomp_kernel(const OcbChompArgs *args, int device, int threads, int min_blocks, size_t smem,
                         const char *robot_header, void **kernel_out, char *err, size_t err_cap);
cudaError_t ocb_jit_launch(void *kernel, const OcbChompArgs *args, int threads, size_t smem, cudaStream_t st);
/* chomp_tiled.cu */
size_t ocb_tile_smem_bytes(const OcbChompArgs *args, int tile_w);
size_t ocb_run_update_smem_bytes(const OcbChompArgs *args);
cudaError_t ocb_launch_chomp_tiled(const OcbChompArgs *args, size_t tile_smem, size_t run_smem,
                                   int run_threads, cudaStream_t st, long *launches);
cudaError_t ocb_launch_init_traj(double *traj, const double *q_start, const double *q_goal,
                                 int R, int P, int n, int floating, cudaStream_t st);
cudaError_t ocb_launch_best(const double *costs, const int *status, int R, int *best_run,
                            double *best_cost, cudaStream_t st);
cudaError_t ocb_launch_sdf_sample(const OcbSdfDev *sdf, const double *d_points, int k, double *d_values,
                                  double *d_grads, int *d_errs, cudaStream_t st);

/* sdf_kernels.cu */
cudaError_t ocb_launch_dt_sqeuc(const double *d_func, double *d_out, const int sizes[3],
                                const double lengths[3], void *scratch, size_t scratch_bytes,
                                cudaStream_t st, long *launches);
size_t ocb_dt_scratch_bytes(const int sizes[3]);
cudaError_t ocb_launch_bin_sdf(const double *d_obs, double *d_sdf, const int sizes[3],
                               const double lengths[3], void *scratch, size_t scratch_bytes,
                               cudaStream_t st, long *launches);
size_t ocb_sdf_scratch_bytes(const int sizes[3], const double lengths[3]);
/* sdf_fast.cu: exact integer path for 0 / HUGE_VAL grids with cubic cells, axes <= 1024 */
int ocb_sdf_fast_eligible(const int sizes[3], const double lengths[3]);
size_t ocb_sdf_fast_scratch_bytes(const int sizes[3]);
cudaError_t ocb_launch_bin_sdf_fast(const double *d_obs, double *d_sdf, const int sizes[3],
                                    const double lengths[3], void *scratch, size_t scratch_bytes,
                                    cudaStream_t st, long *launches, int *used_fast);
cudaError_t ocb_launch_occupancy(const void *d_prims, int n_prims, const int sizes[3],
                                 const double lengths[3], double cube_extent, double *d_grid,
                                 int slices, cudaStream_t st);
cudaError_t ocb_launch_flood_relabel(double *d_grid, const int sizes[3], size_t index_start,
                                     void *scratch, size_t scratch_bytes, cudaStream_t st,
                                     long *launches);
size_t ocb_flood_scratch_bytes(const int sizes[3]);
#ifdef __cplusplus
}
#endif
#endif /* !__CUDACC_RTC__ */

#endif
