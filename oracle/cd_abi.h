/* oracle/cd_abi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Binary-interface mirror of the parts of libcd the CHOMP hot path touches, so
 * that oracle/orcdchomp_port.c can be linked either against the restatement in
 * oracle/libcd_port.c or against the real reference objects compiled from
 * /root/reference/src/libcd (oracle/_ref).  Field order and types follow
 *    struct cd_grid   src/libcd/grid.h:29-41
 *    struct cd_chomp  src/libcd/chomp.h:38-101
 *    struct cd_chomp_con  src/libcd/chomp.h:118-129
 * and the prototypes follow grid.h:43-93, grid_flood.h:33-35, chomp.h:106-140,
 * kin.h (pose functions), mat.h.  When ORACLE_REF_HEADERS is defined the real
 * headers are included instead (the _ref build), which removes any chance of
 * layout drift on that side.
 */
#ifndef ORACLE_CD_ABI_H
#define ORACLE_CD_ABI_H

#include <stdlib.h>
#include <time.h>

#ifdef ORACLE_REF_HEADERS
#include <libcd/grid.h>
#include <libcd/grid_flood.h>
#include <libcd/chomp.h>
#include <libcd/kin.h>
#include <libcd/mat.h>
#include <libcd/spatial.h>
#else

struct cd_grid
{
   int n;
   int *sizes;
   size_t ncells;
   int cell_size;
   char *data;
   double *lengths;
};

int cd_grid_create_sizearray(struct cd_grid **gp, void *cell_init, int cell_size, int n, int *sizes);
int cd_grid_create_copy(struct cd_grid **gp, struct cd_grid *gsrc);
int cd_grid_destroy(struct cd_grid *g);
int cd_grid_index_to_subs(struct cd_grid *g, size_t index, int *subs);
int cd_grid_index_from_subs(struct cd_grid *g, size_t *index, int *subs);
int cd_grid_center_index(struct cd_grid *g, size_t index, double *center);
int cd_grid_lookup_index(struct cd_grid *g, double *p, size_t *index);
void *cd_grid_get_index(struct cd_grid *g, size_t index);
int cd_grid_double_interp(struct cd_grid *g, double *p, double *valuep);
int cd_grid_double_grad(struct cd_grid *g, double *p, double *grad);
int cd_grid_double_dt_sqeuc(struct cd_grid **gp_dt, struct cd_grid *g_func);
int cd_grid_double_bin_sdf(struct cd_grid **gp_dt, struct cd_grid *g_emp);
int cd_grid_flood_fill(struct cd_grid *g, size_t index_start, int *wrap_dim,
                       int (*replace)(void *, void *), void *rptr);

/* chomp.h:118-129 */
struct cd_chomp;
struct cd_chomp_con
{
   struct cd_chomp_con *next;
   int k;
   int i;
   void *cptr;
   int (*con_eval)(void *cptr, struct cd_chomp *c, int i, double *point, double *con_val, double *con_jacobian);
   double *h;
   double *J;
};
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
   int (*cost)(void *cptr, struct cd_chomp *c, int ti, double *point, double *vel,
               double *costp, double *grad);
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

int cd_chomp_create(struct cd_chomp **cp, int m, int n, int D, double *T, int ldt);
void cd_chomp_free(struct cd_chomp *c);
int cd_chomp_add_constraint(struct cd_chomp *c, int k, int i, void *cptr,
   int (*con_eval)(void *cptr, struct cd_chomp *c, int i, double *point, double *con_val, double *con_jacobian));
int cd_chomp_init(struct cd_chomp *c);
int cd_chomp_iterate(struct cd_chomp *c, int do_iteration, double *costp_total,
                     double *costp_obs, double *costp_smooth);

int cd_kin_pose_identity(double pose[7]);
int cd_kin_pose_normalize(double pose[7]);
int cd_kin_pose_compose(const double pose_ab[7], const double pose_bc[7], double pose_ac[7]);
int cd_kin_pose_compos(const double pose_ab[7], const double pos_bc[3], double pos_ac[3]);
int cd_kin_pose_compose_vec(const double pose_ab[7], const double vec_bc[3], double vec_ac[3]);
int cd_kin_pose_invert(const double pose_in[7], double pose_out[7]);
int cd_kin_quat_to_R(const double quat[4], double R[3][3]);
int cd_kin_pose_to_xyzypr(const double pose[7], double xyzypr[6]);
int cd_kin_pose_to_xyzypr_J(const double pose[7], double J[6][7]);
int cd_spatial_xm_from_pose(double xm[6][6], double pose[7]);
int cd_spatial_pose_jac(double pose[7], double jac[6][7]);
int cd_spatial_pose_jac_inverse(double pose[7], double jac_inverse[7][6]);

#endif /* ORACLE_REF_HEADERS */
#endif
