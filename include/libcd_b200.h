/* libcd_b200.h -- libcd-compatible entry points of the B200 SDF build.
 *
 * libcd_b200.so exports, under the reference's own names and with the reference's own
 * struct layout and ownership rules, the grid functions that sit on computedistancefield's
 * path ("lower boundary"):
 *
 *   cd_grid_double_bin_sdf    src/libcd/grid.h:93,  grid.c:637-687
 *   cd_grid_double_dt_sqeuc   src/libcd/grid.h:84,  grid.c:462-569
 *   cd_grid_double_sedt       src/libcd/grid.h:86   (legacy name of the same transform)
 *
 * and one fused replacement for the flood-fill block of the module, whose libcd entry point
 * takes a host callback per cell and so cannot be kept as it is:
 *
 *   cd_grid_b200_flood_relabel   = cd_grid_flood_fill(g, start, 0, replace_1_to_0, 0)
 *                                  + the 1.0 -> HUGE_VAL sweep   (grid_flood.c:30-111,
 *                                  orcdchomp_mod.cpp:143-151, 536-548)
 *
 * Grids are taken from and returned to HOST memory exactly as libcd does: the output grid
 * is a fresh malloc'ed `struct cd_grid` (struct, sizes, lengths and data each their own
 * malloc block, as grid.c:61-132) which the caller releases with libcd's cd_grid_destroy
 * (grid.c:134-143).  Inputs are not modified and not freed.  All work is done on the GPU;
 * without one every call fails (there is no CPU path behind these symbols).
 *
 * Return codes follow libcd: 0 ok, -1 allocation failure, -2 wrong cell type
 * (cell_size != sizeof(double)); additionally -2 for n != 3 (only 3-D grids are
 * accelerated) and -3 for a CUDA / no-device failure; cd_grid_b200_last_error() says which.
 *
 * To use it from the reference: drop the three functions from libcd's grid.c (or link this
 * library first) and link libcd_b200.so + liborcdchomp_b200.so; see INTEGRATION.md section 5.
 */
#ifndef LIBCD_B200_H
#define LIBCD_B200_H

#include <stddef.h>
#include <time.h>
#include "orcdchomp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* layout of src/libcd/grid.h:29-41; define LIBCD_B200_NO_STRUCT when libcd's grid.h is
 * included as well */
#ifndef LIBCD_B200_NO_STRUCT
struct cd_grid
{
   int n;            /* dimensionality */
   int *sizes;       /* [n] cells per axis */
   size_t ncells;
   int cell_size;    /* bytes per cell */
   char *data;       /* C order: index = (x * NY + y) * NZ + z */
   double *lengths;  /* [n] side lengths */
};
#endif

int cd_grid_double_bin_sdf(struct cd_grid **gp_dt, struct cd_grid *g_emp);
int cd_grid_double_dt_sqeuc(struct cd_grid **gp_dt, struct cd_grid *g_func);
int cd_grid_double_sedt(struct cd_grid **gp_dt, struct cd_grid *g_func);

/* in place on g (doubles): 6-connected fill 1.0 -> 0.0 from index_start, then every
 * remaining 1.0 -> HUGE_VAL */
int cd_grid_b200_flood_relabel(struct cd_grid *g, size_t index_start);

/* GPU used by the calls above (default: $OCB_DEVICE or 0); takes effect on the next call */
int cd_grid_b200_set_device(int device);
const char *cd_grid_b200_last_error(void);

/* ------------------------------------------------------------------------- *
 * The cd_chomp C interface (src/libcd/chomp.h:106-140) for ONE run, GPU-backed.
 *
 * Same four entry points, same public struct (chomp.h:38-101), same meaning of the fields
 * the reference's module pokes between calls (mod.cpp:2521-2660, 2752-2831):
 *   n, m, D, T / ldt (the caller's trajectory rows, not owned, chomp.h:45), lambda, dt,
 *   inits[0] / finals[0] (end points), jlimit_lower / jlimit_upper, use_momentum,
 *   leapfrog_first, AG (momentum; the module resamples it for HMC), T_points / AG_points.
 * What cannot cross to the device are the host callbacks cost_pre / cost / cost_extra: the
 * obstacle + self-collision cost of sphere_cost_pre / sphere_cost (mod.cpp:968-1327) is
 * built in instead and configured by cd_chomp_b200_set_sphere_cost() where the module would
 * assign c->cost_pre / c->cost (mod.cpp:2616-2618).  Constraints (cd_chomp_add_constraint)
 * are not offered.  Members the reference materialises only for its own dense algebra
 * (A, Ainv, B, Kvels, Evels, vels, cost_*, Gjlimit*, cons_*) are left NULL: the engine uses
 * the banded factor.  G is allocated and filled with the gradient of the last iteration.
 *
 * cd_chomp_iterate copies T (and AG) in, runs one update (do_iteration != 0) or the cost
 * evaluation alone, and copies T, AG, G back; costs as chomp.c:679-681.  Returns 0, or -1
 * when the joint-limit projection gives up (chomp.c:651-655), -2 for an unsupported set-up
 * (dt other than 1/(m+1), wds other than [0..0,1], free end points), -3 for CUDA failures.
 * For throughput use the batched API (orcdchomp_b200.h); this facade exists so that code
 * written against libcd keeps its shape.                                                   */
#ifndef LIBCD_B200_NO_STRUCT
struct cd_chomp_con;
struct cd_chomp
{
   int n;
   int m;
   double lambda;
   double dt;
   double *T;
   int ldt;
   double **T_points;
   double *G;
   double **G_points;
   double *AG;
   double **AG_points;
   int D;
   double *wds;
   double **inits;
   double **finals;
   double *initsfinals;
   double *A;
   double *Ainv;
   double *B;
   double trC;
   double *jlimit_lower;
   double *jlimit_upper;
   double *Kvels;
   double *Evels;
   double *vels;
   double *cost_nxn;
   double *cost_mxn;
   double *Gjlimit;
   double *GjlimitAinv;
   void *cptr;
   int (*cost_pre)(void *cptr, struct cd_chomp *c, int m, double **T_points);
   int (*cost)(void *cptr, struct cd_chomp *c, int ti, double *point, double *vel, double *costp, double *grad);
   int (*cost_extra)(void *cptr, struct cd_chomp *c, double *T, double *costp, double *G);
   int use_momentum;
   int leapfrog_first;
   struct cd_chomp_con *cons;
   int cons_k;
   double *cons_h;
   double *cons_Jcol;
   double *cons_JAJT;
   int *cons_ipiv;
   double *cons_delta;
   struct timespec ticks_vels;
   struct timespec ticks_callback_pre;
   struct timespec ticks_callbacks;
   struct timespec ticks_smoothgrad;
   struct timespec ticks_smoothcost;
};
#endif

int cd_chomp_create(struct cd_chomp **cp, int m, int n, int D, double *T, int ldt);
void cd_chomp_free(struct cd_chomp *c);
int cd_chomp_init(struct cd_chomp *c);
int cd_chomp_iterate(struct cd_chomp *c, int do_iteration, double *costp_total, double *costp_obs,
                     double *costp_smooth);
/* robot->limit_* are ignored (c->jlimit_* rule, as in the reference); of params only epsilon,
 * epsilon_self, obs_factor, obs_factor_self, floating_base (then c->n = 7 + robot->n_dof and
 * the quaternion re-normalisation of mod.cpp:2805-2808 happens inside cd_chomp_iterate) and the
 * hard constraints are read; sdfs[i].data are host grids, copied.  Constraints: what
 * cd_chomp_add_constraint + con_tsr (chomp.h:131-132, mod.cpp:1330-1497) would register is given as
 * data -- params->constraints[] with where = OCB_CON_START / END / ALL (copied); OCB_CON_START_TSR is
 * only offered by the batch interface. */
int cd_chomp_b200_set_sphere_cost(struct cd_chomp *c, const ocb_robot *robot, const ocb_params *params,
                                  int n_sdfs, const ocb_sdf *sdfs);

#ifdef __cplusplus
}
#endif
#endif
