/* oracle/orcdchomp_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the numeric parts of or_cdchomp's OpenRAVE module that sit
 * on the CHOMP hot path (src/orcdchomp_mod.cpp), written against the libcd C
 * interface (oracle/cd_abi.h) so the same file links against
 *    - oracle/libcd_port.c            -> build/liboracle_port.so  ("port")
 *    - the unmodified reference libcd -> _ref/liboracle_ref.so    ("reference")
 * It is the checker / CPU baseline for the CUDA engine.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product never does.
 *
 * PARITY STATUS.  The reference has no tests or golden vectors for this path
 * (SURVEY.md section 4).  The libcd half is pinned by running the reference's own
 * sources (oracle/_ref).  Three third-party pieces are absent from
 * /root/reference and unpinned upstream (package.xml:13-21 names no versions):
 *   - OpenRAVE (SetActiveDOFValues / Link::GetTransform / CalculateJacobian,
 *     mod.cpp:1026,1033,1048): restated here as forward kinematics + geometric
 *     Jacobian over the explicit tree in `struct ocb_robot`
 *     (include/orcdchomp_b200.h) using libcd's own pose algebra.  "parity
 *     unpinned" for this piece: OpenRAVE's arithmetic order is unknown.
 *   - GSL (gsl_rng_mt19937 as gsl_rng_default, gsl_rng_set, gsl_ran_gaussian,
 *     gsl_rng_uniform; mod.cpp:2303-2304, 2763, 2767): restated from the
 *     published MT19937 algorithm (Matsumoto & Nishimura 1998, 2002 seeding) and
 *     GSL's documented polar Box-Muller; pinned against numpy's MT19937.
 *   - OpenRAVE CheckCollision(cube) (mod.cpp:520): replaced by the analytic
 *     cube-vs-primitive overlap defined in include/orcdchomp_b200.h.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "cd_abi.h"
#include "../include/orcdchomp_b200.h"

/* ------------------------------------------------------------------ MT19937 */
/* GSL rng/mt.c semantics: seed 0 -> 4357; mt[i] = 1812433253*(mt[i-1]^(mt[i-1]>>30))+i;
 * get_double = get()/2^32; uniform_pos redraws zeros. */
struct orc_mt
{
   unsigned int mt[624];
   int mti;
};

void orc_mt_seed(struct orc_mt *s, unsigned int seed)
{
   int i;
   if (seed == 0) seed = 4357;
   s->mt[0] = seed & 0xffffffffu;
   for (i = 1; i < 624; i++)
      s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (unsigned int) i;
   s->mti = 624;
}

unsigned int orc_mt_next(struct orc_mt *s)
{
   unsigned int y;
   if (s->mti >= 624)
   {
      int k;
      for (k = 0; k < 624; k++)
      {
         unsigned int a = s->mt[k], b = s->mt[(k + 1) % 624];
         y = (a & 0x80000000u) | (b & 0x7fffffffu);
         s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      s->mti = 0;
   }
   y = s->mt[s->mti++];
   y ^= (y >> 11);
   y ^= (y << 7) & 0x9d2c5680u;
   y ^= (y << 15) & 0xefc60000u;
   y ^= (y >> 18);
   return y;
}

double orc_mt_uniform(struct orc_mt *s) { return orc_mt_next(s) / 4294967296.0; }

static double mt_uniform_pos(struct orc_mt *s)
{
   double x;
   do { x = orc_mt_uniform(s); } while (x == 0.0);
   return x;
}

/* gsl_ran_gaussian (randist/gauss.c, polar method): one variate per accepted
 * pair, the second is discarded */
double orc_mt_gaussian(struct orc_mt *s, double sigma)
{
   double x, y, r2;
   do
   {
      x = -1.0 + 2.0 * mt_uniform_pos(s);
      y = -1.0 + 2.0 * mt_uniform_pos(s);
      r2 = x * x + y * y;
   } while (r2 > 1.0 || r2 == 0.0);
   return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

/* ------------------------------------------------------------ run structure */
/* mirrors struct run / run_sphere / run_rsdf (mod.cpp:850-966) with arrays in
 * place of the linked lists */
struct orc_rsdf
{
   double pose_world_gsdf[7];
   double pose_gsdf_world[7];
   struct cd_grid *grid;
};

struct orc_run
{
   const struct ocb_robot *robot; /* borrowed; must outlive the run */
   double *traj;
   int n_points;
   int n;
   double epsilon, epsilon_self, obs_factor, obs_factor_self;
   int n_spheres, n_spheres_active;
   int *sph_src;      /* [n_spheres] index into the robot sphere table, active first */
   double *sphere_poss_inactive;
   double *sphere_poss_all;
   double *sphere_poss;
   double *sphere_vels;
   double *sphere_accs;
   double *sphere_jacs;
   double *J2;
   double *Jadof;      /* scratch 3 x n_dof (floating base) */
   double *link_poses; /* scratch [n_links][7] */
   int n_rsdfs;
   struct orc_rsdf *rsdfs;
   int use_hmc;
   int hmc_resample_iter;
   double hmc_resample_lambda;
   struct orc_mt rng;
   struct cd_chomp *c;
   int iter;
   int floating; /* floating_base: rows are [x y z qx qy qz qw, adofs] (mod.cpp:991-1021) */
   int start_tsr; /* the start point is optimised too: c->m == n_points-1 (mod.cpp:2316) */
   int n_cons;
   struct orc_con *cons; /* struct run_contsr / start_tsr / everyn_tsr (mod.cpp:873-884, 926-940) */
};

/* one TSR constraint bound to a run: the frame, the two fixed TSR poses, which of
 * [x y z roll pitch yaw] are held at zero, and how many (k) */
struct orc_con
{
   struct orc_run *r;
   struct ocb_constraint spec;
   int enabled[6];
   int k;
};

/* ------------------------------------------------- kinematics (OpenRAVE side) */
/* forward kinematics of every link for active-dof vector q; stands in for
 * robot->SetActiveDOFValues (mod.cpp:1026) + Link::GetTransform (1033) */
static void orc_fk_base(const struct ocb_robot *rb, const double *base_pose, const double *q, double *link_poses);

void orc_fk(const struct ocb_robot *rb, const double *q, double *link_poses)
{
   orc_fk_base(rb, rb->base_pose, q, link_poses);
}

/* the same with the base at `base_pose`: robot->SetTransform(t) before SetActiveDOFValues in the
 * floating-base branch (mod.cpp:1009-1020) */
static void orc_fk_base(const struct ocb_robot *rb, const double *base_pose, const double *q, double *link_poses)
{
   int i;
   memcpy(link_poses, base_pose, 7 * sizeof(double));
   for (i = 1; i < rb->n_links; i++)
   {
      double motion[7], tmp[7];
      double value = rb->dof_coeff[2 * i + 1];
      const double *ax = &rb->axis[3 * i];
      if (rb->dof_index[i] >= 0) value += rb->dof_coeff[2 * i] * q[rb->dof_index[i]];
      cd_kin_pose_identity(motion);
      if (rb->joint_type[i] == OCB_JOINT_REVOLUTE)
      {
         double s = sin(0.5 * value), co = cos(0.5 * value);
         motion[3] = ax[0] * s;
         motion[4] = ax[1] * s;
         motion[5] = ax[2] * s;
         motion[6] = co;
      }
      else if (rb->joint_type[i] == OCB_JOINT_PRISMATIC)
      {
         motion[0] = ax[0] * value;
         motion[1] = ax[1] * value;
         motion[2] = ax[2] * value;
      }
      cd_kin_pose_compose(&link_poses[7 * rb->parent[i]], &rb->pose_parent[7 * i], tmp);
      cd_kin_pose_compose(tmp, motion, &link_poses[7 * i]);
   }
}

/* 3 x n linear Jacobian of world point p rigidly attached to `link`; stands in
 * for robot->CalculateJacobian(linkindex, v, J) with the active columns picked
 * (mod.cpp:1048, 1087-1093): revolute column = axis x (p - anchor), prismatic
 * column = axis, scaled by d(value)/d(dof) */
void orc_jacobian(const struct ocb_robot *rb, const double *link_poses, int link,
                  const double p[3], double *J /* 3 x n */)
{
   int n = rb->n_dof, a, k;
   for (k = 0; k < 3 * n; k++) J[k] = 0.0;
   for (a = link; a > 0; a = rb->parent[a])
   {
      double axw[3], col[3];
      int dof = rb->dof_index[a];
      if (dof < 0 || rb->joint_type[a] == OCB_JOINT_FIXED) continue;
      cd_kin_pose_compose_vec(&link_poses[7 * a], &rb->axis[3 * a], axw);
      if (rb->joint_type[a] == OCB_JOINT_REVOLUTE)
      {
         double r[3];
         r[0] = p[0] - link_poses[7 * a + 0];
         r[1] = p[1] - link_poses[7 * a + 1];
         r[2] = p[2] - link_poses[7 * a + 2];
         col[0] = axw[1] * r[2] - axw[2] * r[1];
         col[1] = axw[2] * r[0] - axw[0] * r[2];
         col[2] = axw[0] * r[1] - axw[1] * r[0];
      }
      else
      {
         col[0] = axw[0]; col[1] = axw[1]; col[2] = axw[2];
      }
      for (k = 0; k < 3; k++) J[k * n + dof] += rb->dof_coeff[2 * a] * col[k];
   }
}

/* cd_spatial_pose_jac (src/libcd/spatial.c:295-337): maps the derivatives of a pose
 * [x y z qx qy qz qw] to the world spatial velocity [omega; v of the point at the origin] */
static void orc_pose_jac(const double pose[7], double jac[6][7])
{
   double x = pose[0], y = pose[1], z = pose[2];
   double qxt2 = 2.0 * pose[3], qyt2 = 2.0 * pose[4], qzt2 = 2.0 * pose[5], qwt2 = 2.0 * pose[6];
   int i, j;
   for (i = 0; i < 6; i++) for (j = 0; j < 7; j++) jac[i][j] = 0.0;
   jac[3][0] = 1.0; jac[4][1] = 1.0; jac[5][2] = 1.0;
   jac[0][3] = qwt2;  jac[0][4] = -qzt2; jac[0][5] = qyt2;  jac[0][6] = -qxt2;
   jac[1][3] = qzt2;  jac[1][4] = qwt2;  jac[1][5] = -qxt2; jac[1][6] = -qyt2;
   jac[2][3] = -qyt2; jac[2][4] = qxt2;  jac[2][5] = qwt2;  jac[2][6] = -qzt2;
   jac[3][3] = -z * qzt2 - y * qyt2; jac[3][4] = -z * qwt2 + y * qxt2;
   jac[3][5] = z * qxt2 + y * qwt2;  jac[3][6] = z * qyt2 - y * qzt2;
   jac[4][3] = z * qwt2 + x * qyt2;  jac[4][4] = -z * qzt2 - x * qxt2;
   jac[4][5] = z * qyt2 - x * qwt2;  jac[4][6] = -z * qxt2 + x * qzt2;
   jac[5][3] = -y * qwt2 + x * qzt2; jac[5][4] = y * qzt2 + x * qwt2;
   jac[5][5] = -y * qyt2 - x * qxt2; jac[5][6] = y * qxt2 - x * qyt2;
}

/* the left 3 x 7 block of a sphere's Jacobian in floating-base mode (mod.cpp:1050-1080):
 * rows 3..5 of the motion transform to the frame at -v (identity rotation, so [rx] I | I with
 * r = -v, spatial.c:71-102) times Jsp, then the reference's 0.01 scaling.  J has row length n. */
static void orc_pose_columns(const double Jsp[6][7], const double v[3], double *J, int n)
{
   double Xm3[3][6];
   int i, k, l;
   for (i = 0; i < 3; i++) for (l = 0; l < 6; l++) Xm3[i][l] = 0.0;
   Xm3[0][1] = v[2];  Xm3[0][2] = -v[1];
   Xm3[1][0] = -v[2]; Xm3[1][2] = v[0];
   Xm3[2][0] = v[1];  Xm3[2][1] = -v[0];
   Xm3[0][3] = 1.0; Xm3[1][4] = 1.0; Xm3[2][5] = 1.0;
   for (i = 0; i < 3; i++)
      for (k = 0; k < 7; k++)
      {
         double acc = 0.0;
         for (l = 0; l < 6; l++) acc += Xm3[i][l] * Jsp[l][k];
         J[i * n + k] = acc * 0.01;
      }
}

/* test hook: the 3 x 7 pose block for a point v of a body at `pose` */
void orc_pose_block(const double pose[7], const double v[3], double J[21])
{
   double Jsp[6][7];
   orc_pose_jac(pose, Jsp);
   orc_pose_columns(Jsp, v, J, 7);
}

/* robot->DoesAffect(adof, linkindex) over all active dofs (mod.cpp:2270-2273) */
static int link_is_active(const struct ocb_robot *rb, int link)
{
   int a;
   for (a = link; a > 0; a = rb->parent[a])
      if (rb->dof_index[a] >= 0 && rb->joint_type[a] != OCB_JOINT_FIXED) return 1;
   return 0;
}

/* 3 x n angular-velocity Jacobian of `link`; stands in for
 * robot->CalculateAngularVelocityJacobian(linkindex, J) with the active columns picked
 * (mod.cpp:1456-1459): revolute column = world axis, prismatic column = 0 */
static void orc_angular_jacobian(const struct ocb_robot *rb, const double *link_poses, int link, double *J)
{
   int n = rb->n_dof, a, k;
   for (k = 0; k < 3 * n; k++) J[k] = 0.0;
   for (a = link; a > 0; a = rb->parent[a])
   {
      double axw[3];
      int dof = rb->dof_index[a];
      if (dof < 0 || rb->joint_type[a] != OCB_JOINT_REVOLUTE) continue;
      cd_kin_pose_compose_vec(&link_poses[7 * a], &rb->axis[3 * a], axw);
      for (k = 0; k < 3; k++) J[k * n + dof] += rb->dof_coeff[2 * a] * axw[k];
   }
}

/* con_tsr / con_everyn_tsr / con_start_tsr (mod.cpp:1330-1497, 1500-1657, 1659-1784; the three
 * differ only in where the frame and the TSR come from).  Value: the pose of the constrained frame
 * seen from the TSR, inv(T0w) * frame * inv(Twe), as [x y z yaw pitch roll], enabled entries in
 * x y z roll pitch yaw order.  Jacobian: the frame's world spatial Jacobian [angular; linear at
 * the world origin], moved to the TSR frame, turned into pose rates and then into xyzypr rates. */
static int orc_con_tsr(void *cptr, struct cd_chomp *c, int ti, double *point, double *con_val, double *con_jacobian)
{
   struct orc_con *con = (struct orc_con *) cptr;
   struct orc_run *r = con->r;
   const struct ocb_robot *rb = r->robot;
   double pose_ee[7], pose_ee_obj[7], pose_obj[7], pose_table_world[7], pose_table_obj[7], xyzypr[6];
   int tsri, ki, i, j, l, n = c->n;
   (void) ti;
   if (r->floating)
      orc_fk_base(rb, point, point + 7, r->link_poses);
   else
      orc_fk(rb, point, r->link_poses);
   cd_kin_pose_compose(&r->link_poses[7 * con->spec.link], con->spec.pose_link_ee, pose_ee);
   cd_kin_pose_invert(con->spec.Twe, pose_ee_obj);
   cd_kin_pose_compose(pose_ee, pose_ee_obj, pose_obj);
   cd_kin_pose_invert(con->spec.T0w, pose_table_world);
   cd_kin_pose_compose(pose_table_world, pose_obj, pose_table_obj);
   cd_kin_pose_to_xyzypr(pose_table_obj, xyzypr);
   for (ki = 0, tsri = 0; tsri < 6; tsri++)
      if (con->enabled[tsri]) con_val[ki++] = xyzypr[tsri < 3 ? tsri : 8 - tsri];
   if (con_jacobian)
   {
      double *spajac = (double *) calloc((size_t) 6 * n, sizeof(double));
      double *Jpart = (double *) malloc((size_t) 3 * (rb->n_dof ? rb->n_dof : 1) * sizeof(double));
      double *full = (double *) malloc((size_t) 6 * n * sizeof(double));
      double xm[6][6], ji[7][6], Jx[6][7], t1[6][6], t2[6][6];
      const double origin[3] = {0.0, 0.0, 0.0};
      int off = r->floating ? 7 : 0, na = rb->n_dof;
      if (r->floating)
      {
         double Jsp[6][7];
         cd_spatial_pose_jac(point, Jsp);
         for (i = 0; i < 6; i++) for (j = 0; j < 7; j++) spajac[i * n + j] = Jsp[i][j];
      }
      orc_angular_jacobian(rb, r->link_poses, con->spec.link, Jpart);
      for (i = 0; i < 3; i++) for (j = 0; j < na; j++) spajac[i * n + off + j] = Jpart[i * na + j];
      orc_jacobian(rb, r->link_poses, con->spec.link, origin, Jpart);
      for (i = 0; i < 3; i++) for (j = 0; j < na; j++) spajac[(3 + i) * n + off + j] = Jpart[i * na + j];
      cd_spatial_xm_from_pose(xm, pose_table_world);
      cd_spatial_pose_jac_inverse(pose_table_obj, ji);
      cd_kin_pose_to_xyzypr_J(pose_table_obj, Jx);
      for (i = 0; i < 6; i++)
         for (j = 0; j < 6; j++)
         {
            double acc = 0.0;
            for (l = 0; l < 7; l++) acc += Jx[i][l] * ji[l][j];
            t1[i][j] = acc;
         }
      for (i = 0; i < 6; i++)
         for (j = 0; j < 6; j++)
         {
            double acc = 0.0;
            for (l = 0; l < 6; l++) acc += t1[i][l] * xm[l][j];
            t2[i][j] = acc;
         }
      for (i = 0; i < 6; i++)
         for (j = 0; j < n; j++)
         {
            double acc = 0.0;
            for (l = 0; l < 6; l++) acc += t2[i][l] * spajac[l * n + j];
            full[i * n + j] = acc;
         }
      for (ki = 0, tsri = 0; tsri < 6; tsri++)
         if (con->enabled[tsri])
         {
            memcpy(con_jacobian + (size_t) ki * n, full + (size_t) (tsri < 3 ? tsri : 8 - tsri) * n, n * sizeof(double));
            ki++;
         }
      free(spajac);
      free(Jpart);
      free(full);
   }
   return 0;
}

/* test hook: value and Jacobian of one constraint at configuration `point` */
int orc_run_constraint_eval(struct orc_run *r, int index, double *point, double *con_val, double *con_jacobian)
{
   if (index < 0 || index >= r->n_cons) return -2;
   orc_con_tsr(&r->cons[index], r->c, 0, point, con_val, con_jacobian);
   return r->cons[index].k;
}

/* ------------------------------------------------ callbacks (mod.cpp:968-1327) */

/* sphere_cost_pre, non-floating-base branch (mod.cpp:968-1132, 1022-1028,
 * 1031-1049, 1087-1093): sphere positions at all P waypoints, Jacobians at the
 * moving ones, then central-difference velocities and accelerations (1099-1127) */
static int orc_sphere_cost_pre(void *cptr, struct cd_chomp *c, int m, double **T_points)
{
   struct orc_run *r = (struct orc_run *) cptr;
   const struct ocb_robot *rb = r->robot;
   int sa = r->n_spheres_active;
   int ti, s, k;
   size_t row = (size_t) sa * 3;
   (void) T_points;
   for (ti = 0; ti < r->n_points; ti++)
   {
      int ti_mov = (c->m == r->n_points - 2) ? ti - 1 : ti; /* mod.cpp:1040-1043 */
      double Jsp[6][7];
      if (r->floating)
      {
         orc_pose_jac(&r->traj[ti * c->n], Jsp);
         orc_fk_base(rb, &r->traj[ti * c->n], &r->traj[ti * c->n + 7], r->link_poses);
      }
      else
         orc_fk(rb, &r->traj[ti * c->n], r->link_poses);
      for (s = 0; s < sa; s++)
      {
         int src = r->sph_src[s];
         int link = rb->sphere_link[src];
         double v[3];
         double *J;
         cd_kin_pose_compos(&r->link_poses[7 * link], &rb->sphere_pos[3 * src], v);
         for (k = 0; k < 3; k++) r->sphere_poss_all[ti * row + s * 3 + k] = v[k];
         if (ti_mov < 0 || m <= ti_mov) continue;
         J = &r->sphere_jacs[((size_t) ti_mov * sa + s) * 3 * c->n];
         if (r->floating)
         {
            int i, j, na = rb->n_dof;
            orc_pose_columns(Jsp, v, J, c->n);
            orc_jacobian(rb, r->link_poses, link, v, r->Jadof);
            for (i = 0; i < 3; i++)
               for (j = 0; j < na; j++) J[i * c->n + 7 + j] = r->Jadof[i * na + j]; /* mod.cpp:1083-1085 */
         }
         else
            orc_jacobian(rb, r->link_poses, link, v, J);
      }
   }
   {
   /* with start_tsr the interior rows start one further down and row 0 belongs to the start point */
   size_t shift = (c->m != r->n_points - 2) ? row : 0;
   /* velocities: (p[t+1] - p[t-1]) * (1/(2 dt))   (mod.cpp:1101-1107) */
   for (ti = 0; ti < r->n_points - 2; ti++)
      for (k = 0; k < (int) row; k++)
      {
         double x = r->sphere_poss_all[(ti + 2) * row + k];
         x -= r->sphere_poss_all[ti * row + k];
         x *= 1.0 / (2.0 * c->dt);
         r->sphere_vels[shift + ti * row + k] = x;
      }
   /* the start point: one-sided difference (mod.cpp:1108-1114) */
   if (shift)
      for (k = 0; k < (int) row; k++)
      {
         double x = r->sphere_poss_all[row + k];
         x -= r->sphere_poss_all[k];
         x *= 1.0 / (c->dt);
         r->sphere_vels[k] = x;
      }
   /* accelerations: (-2 p[t] + p[t-1] + p[t+1]) * (1/dt^2)   (mod.cpp:1118-1125) */
   for (ti = 0; ti < r->n_points - 2; ti++)
      for (k = 0; k < (int) row; k++)
      {
         double x = r->sphere_poss_all[(ti + 1) * row + k];
         x *= -2.0;
         x += r->sphere_poss_all[ti * row + k];
         x += r->sphere_poss_all[(ti + 2) * row + k];
         x *= 1.0 / (c->dt * c->dt);
         r->sphere_accs[shift + ti * row + k] = x;
      }
   /* the start point takes the acceleration of its neighbour (mod.cpp:1126-1128) */
   if (shift)
      for (k = 0; k < (int) row; k++) r->sphere_accs[k] = r->sphere_accs[row + k];
   }
   return 0;
}

static double nrm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* sphere_cost (mod.cpp:1134-1327): obstacle + self-collision cost of moving
 * waypoint ti and its configuration-space gradient */
static int orc_sphere_cost(void *cptr, struct cd_chomp *c, int ti, double *c_point,
                           double *c_vel, double *costp, double *c_grad)
{
   struct orc_run *r = (struct orc_run *) cptr;
   const struct ocb_robot *rb = r->robot;
   int n = c->n, sa = r->n_spheres_active;
   size_t row = (size_t) sa * 3;
   double cost = 0.0;
   int s, s2, i, k, j;
   (void) c_point; (void) c_vel;
   if (c_grad) for (j = 0; j < n; j++) c_grad[j] = 0.0;

   for (s = 0; s < sa; s++)
   {
      const double *x_pos = &r->sphere_poss[ti * row + s * 3];
      const double *x_vel = &r->sphere_vels[ti * row + s * 3];
      const double *Js = &r->sphere_jacs[((size_t) ti * sa + s) * 3 * n];
      double radius = rb->sphere_radius[r->sph_src[s]];
      int link = rb->sphere_link[r->sph_src[s]];
      double x_vel_norm = nrm3(x_vel);
      double cost_sphere = 0.0;
      double g_point[3], dist, best_dist = HUGE_VAL;
      int best = -1;

      /* closest field = smallest interpolated value; first wins ties (1169-1189) */
      for (i = 0; i < r->n_rsdfs; i++)
      {
         cd_kin_pose_compos(r->rsdfs[i].pose_gsdf_world, x_pos, g_point);
         if (cd_grid_double_interp(r->rsdfs[i].grid, g_point, &dist)) continue;
         if (dist < best_dist) { best_dist = dist; best = i; }
      }
      if (best != -1)
      {
         cd_kin_pose_compos(r->rsdfs[best].pose_gsdf_world, x_pos, g_point);
         cd_grid_double_interp(r->rsdfs[best].grid, g_point, &dist);
         dist -= radius;
         if (dist < 0.0)
            cost_sphere += x_vel_norm * r->obs_factor * (0.5 * r->epsilon - dist);
         else if (dist < r->epsilon)
            cost_sphere += x_vel_norm * r->obs_factor * (0.5 / r->epsilon) * (dist - r->epsilon) * (dist - r->epsilon);
         if (c_grad)
         {
            double g_grad[3], x_grad[3], x_curv[3], proj;
            cd_grid_double_grad(r->rsdfs[best].grid, g_point, g_grad);
            cd_kin_pose_compose_vec(r->rsdfs[best].pose_world_gsdf, g_grad, g_grad);
            /* scale by -1 inside, (d/eps - 1) in the margin, zero outside (1216-1223) */
            for (k = 0; k < 3; k++) x_grad[k] = g_grad[k];
            if (dist < 0.0)
               for (k = 0; k < 3; k++) x_grad[k] *= -1.0;
            else if (dist < r->epsilon)
               for (k = 0; k < 3; k++) x_grad[k] *= dist / r->epsilon - 1.0;
            else
               for (k = 0; k < 3; k++) x_grad[k] = 0.0;
            for (k = 0; k < 3; k++) x_grad[k] *= x_vel_norm * r->obs_factor;
            if (x_vel_norm > 0.000001)
            {
               proj = dot3(x_grad, x_vel) / (x_vel_norm * x_vel_norm);
               for (k = 0; k < 3; k++) x_grad[k] += -proj * x_vel[k];
            }
            for (k = 0; k < 3; k++) x_curv[k] = r->sphere_accs[ti * row + s * 3 + k];
            if (x_vel_norm > 0.000001)
            {
               proj = dot3(x_curv, x_vel) / (x_vel_norm * x_vel_norm);
               for (k = 0; k < 3; k++) x_curv[k] += -proj * x_vel[k];
            }
            for (k = 0; k < 3; k++) x_curv[k] *= 1.0 / (x_vel_norm * x_vel_norm);
            /* cblas_daxpy returns at once when alpha == 0 (1241): inf / NaN curvature of a
             * sphere at rest is never read */
            if (cost_sphere != 0.0)
               for (k = 0; k < 3; k++) x_grad[k] += -cost_sphere * x_curv[k];
            /* c_grad += x_vel_norm * J^T x_grad   (dgemv, 1244-1245); alpha == 0 with beta == 1
             * is BLAS's quick return as well */
            if (x_vel_norm != 0.0)
               for (j = 0; j < n; j++)
               {
                  double acc = 0.0;
                  for (k = 0; k < 3; k++) acc += Js[k * n + j] * x_grad[k];
                  c_grad[j] += x_vel_norm * acc;
               }
         }
      }

      /* self collision against every sphere on another link (1251-1317) */
      for (s2 = 0; s2 < r->n_spheres; s2++)
      {
         double radius2 = rb->sphere_radius[r->sph_src[s2]];
         double from_other[3], g_grad[3];
         if (link == rb->sphere_link[r->sph_src[s2]]) continue;
         for (k = 0; k < 3; k++) from_other[k] = x_pos[k];
         if (s2 < sa)
            for (k = 0; k < 3; k++) from_other[k] -= r->sphere_poss[ti * row + s2 * 3 + k];
         else
            for (k = 0; k < 3; k++) from_other[k] -= r->sphere_poss_inactive[(s2 - sa) * 3 + k];
         dist = nrm3(from_other);
         if (dist > radius + radius2 + r->epsilon_self) continue;
         if (c_grad)
            for (k = 0; k < 3; k++) g_grad[k] = from_other[k] / dist;
         dist -= radius + radius2;
         if (costp)
         {
            if (dist < 0.0)
               cost_sphere += x_vel_norm * r->obs_factor_self * (0.5 * r->epsilon_self - dist);
            else
               cost_sphere += x_vel_norm * r->obs_factor_self * (0.5 / r->epsilon_self) * (dist - r->epsilon_self) * (dist - r->epsilon_self);
         }
         if (c_grad)
         {
            double x_grad[3], proj;
            for (k = 0; k < 3; k++) x_grad[k] = g_grad[k];
            if (dist < 0.0)
               for (k = 0; k < 3; k++) x_grad[k] *= -1.0;
            else if (dist < r->epsilon_self)
               for (k = 0; k < 3; k++) x_grad[k] *= dist / r->epsilon_self - 1.0;
            for (k = 0; k < 3; k++) x_grad[k] *= x_vel_norm * r->obs_factor_self;
            if (x_vel_norm > 0.000001)
            {
               proj = dot3(x_grad, x_vel) / (x_vel_norm * x_vel_norm);
               for (k = 0; k < 3; k++) x_grad[k] += -proj * x_vel[k];
            }
            for (k = 0; k < 3 * n; k++) r->J2[k] = Js[k];
            if (s2 < sa)
            {
               const double *Jo = &r->sphere_jacs[((size_t) ti * sa + s2) * 3 * n];
               for (k = 0; k < 3 * n; k++) r->J2[k] -= Jo[k];
            }
            for (j = 0; j < n; j++)
            {
               double acc = 0.0;
               for (k = 0; k < 3; k++) acc += r->J2[k * n + j] * x_grad[k];
               c_grad[j] += acc;
            }
         }
      }
      cost += cost_sphere;
   }
   if (costp) *costp = cost;
   return 0;
}

/* ------------------------------------------------------ create / iterate / ... */

void orc_run_destroy(struct orc_run *r)
{
   int i;
   if (!r) return;
   if (r->c) cd_chomp_free(r->c);
   for (i = 0; i < r->n_rsdfs; i++) cd_grid_destroy(r->rsdfs[i].grid);
   free(r->rsdfs);
   free(r->traj);
   free(r->sph_src);
   free(r->sphere_poss_inactive);
   free(r->sphere_poss_all);
   free(r->sphere_vels);
   free(r->sphere_accs);
   free(r->sphere_jacs);
   free(r->J2);
   free(r->Jadof);
   free(r->link_poses);
   free(r->cons);
   free(r);
}

/* numeric part of mod::create (mod.cpp:2266-2299 sphere split, 2315-2345 buffers,
 * 2347-2369 rooted sdfs, 2417-2464 straight line, 2521 cd_chomp_create, 2567-2580
 * dt / inits / finals, 2617-2664 callbacks, lambda, momentum, limits, init) */
int orc_run_create(const struct ocb_robot *rb, const struct ocb_params *pr, int n_sdfs,
                   const struct ocb_sdf *sdfs, const double *q_start, const double *q_goal,
                   unsigned int seed, struct orc_run **out)
{
   struct orc_run *r;
   struct cd_chomp *c = 0;
   int fl = pr->floating_base ? 1 : 0;
   int n = rb->n_dof + (fl ? 7 : 0), P = pr->n_points, m = P - 2;
   int i, j, s, na = 0, ni = 0, start_tsr = 0;
   if (pr->lambda < 0.01 || P < 3 || n_sdfs < 1) return -2;
   for (i = 0; i < pr->n_constraints; i++)
      if (pr->constraints[i].where == OCB_CON_START_TSR) start_tsr++;
   if (start_tsr > 1 || (start_tsr && fl)) return -2; /* mod.cpp:2100 */
   if (start_tsr) m++;                                /* mod.cpp:2316 */
   r = (struct orc_run *) calloc(1, sizeof(struct orc_run));
   if (!r) return -1;
   r->floating = fl;
   r->robot = rb;
   r->n_points = P;
   r->n = n;
   r->epsilon = pr->epsilon;
   r->epsilon_self = pr->epsilon_self;
   r->obs_factor = pr->obs_factor;
   r->obs_factor_self = pr->obs_factor_self;
   r->use_hmc = pr->use_hmc;
   r->hmc_resample_lambda = pr->hmc_resample_lambda;
   r->hmc_resample_iter = 0;
   orc_mt_seed(&r->rng, seed);

   /* active spheres first (XML order), then inactive (XML order): SURVEY A.6 */
   r->n_spheres = rb->n_spheres;
   r->sph_src = (int *) malloc(rb->n_spheres * sizeof(int));
   /* "if we're floating, call all spheres active" (mod.cpp:2274) */
   for (s = 0; s < rb->n_spheres; s++)
      if (fl || link_is_active(rb, rb->sphere_link[s])) r->sph_src[na++] = s;
   for (s = 0; s < rb->n_spheres; s++)
      if (!(fl || link_is_active(rb, rb->sphere_link[s]))) r->sph_src[na + ni++] = s;
   r->n_spheres_active = na;
   if (!na) { orc_run_destroy(r); return -2; }

   r->J2 = (double *) malloc(3 * n * sizeof(double));
   r->Jadof = (double *) malloc(3 * (rb->n_dof ? rb->n_dof : 1) * sizeof(double));
   r->link_poses = (double *) malloc((size_t) rb->n_links * 7 * sizeof(double));
   r->sphere_poss_all = (double *) malloc((size_t) P * na * 3 * sizeof(double));
   r->start_tsr = start_tsr;
   r->sphere_poss = r->sphere_poss_all + (start_tsr ? 0 : (size_t) na * 3); /* mod.cpp:2320-2323 */
   r->sphere_vels = (double *) malloc((size_t) m * na * 3 * sizeof(double));
   r->sphere_accs = (double *) malloc((size_t) m * na * 3 * sizeof(double));
   r->sphere_jacs = (double *) malloc((size_t) m * na * 3 * n * sizeof(double));
   r->sphere_poss_inactive = (double *) malloc((size_t) (ni ? ni : 1) * 3 * sizeof(double));

   /* inactive spheres are frozen at the robot's current configuration (none when floating) */
   if (!fl) orc_fk(rb, q_start, r->link_poses);
   for (s = 0; s < ni; s++)
   {
      int src = r->sph_src[na + s];
      cd_kin_pose_compos(&r->link_poses[7 * rb->sphere_link[src]], &rb->sphere_pos[3 * src],
                         &r->sphere_poss_inactive[3 * s]);
   }

   r->n_rsdfs = n_sdfs;
   r->rsdfs = (struct orc_rsdf *) calloc(n_sdfs, sizeof(struct orc_rsdf));
   for (i = 0; i < n_sdfs; i++)
   {
      double zero = 0.0;
      int sizes[3];
      for (j = 0; j < 3; j++) sizes[j] = sdfs[i].sizes[j];
      if (cd_grid_create_sizearray(&r->rsdfs[i].grid, &zero, sizeof(double), 3, sizes))
      { orc_run_destroy(r); return -1; }
      memcpy(r->rsdfs[i].grid->data, sdfs[i].data, r->rsdfs[i].grid->ncells * sizeof(double));
      for (j = 0; j < 3; j++) r->rsdfs[i].grid->lengths[j] = sdfs[i].lengths[j];
      memcpy(r->rsdfs[i].pose_world_gsdf, sdfs[i].pose_world_gsdf, 7 * sizeof(double));
      cd_kin_pose_invert(r->rsdfs[i].pose_world_gsdf, r->rsdfs[i].pose_gsdf_world);
   }

   /* straight line, evaluated in place exactly as mod.cpp:2456-2458 */
   r->traj = (double *) malloc((size_t) P * n * sizeof(double));
   for (i = 0; i < P * n; i++) r->traj[i] = 0.0;
   for (j = 0; j < n; j++) r->traj[j] = q_start[j];
   for (j = 0; j < n; j++) r->traj[(P - 1) * n + j] = q_goal[j];
   for (i = 0; i < P; i++)
      for (j = 0; j < n; j++)
         r->traj[i * n + j] = r->traj[j] + (r->traj[(P - 1) * n + j] - r->traj[j]) * i / (P - 1);
   if (fl) /* mod.cpp:2461-2464 */
      for (i = 0; i < P; i++) cd_kin_pose_normalize(&r->traj[i * n]);

   if (cd_chomp_create(&c, m, n, pr->derivative, &r->traj[start_tsr ? 0 : n], n)) { orc_run_destroy(r); return -1; } /* mod.cpp:2521 */
   r->c = c;
   c->dt = 1.0 / (P - 1);
   c->inits[0] = start_tsr ? 0 : &r->traj[0]; /* mod.cpp:2571-2578 */
   c->finals[0] = &r->traj[(P - 1) * n];
   c->cptr = r;
   c->cost_pre = orc_sphere_cost_pre;
   c->cost = orc_sphere_cost;
   c->lambda = pr->lambda;
   if (pr->use_momentum) c->use_momentum = 1;
   for (j = 0; j < n; j++)
   {
      /* mod.cpp:2639-2660: the pose entries are unbounded */
      c->jlimit_lower[j] = (fl && j < 7) ? -HUGE_VAL : rb->limit_lower[j - (fl ? 7 : 0)];
      c->jlimit_upper[j] = (fl && j < 7) ? HUGE_VAL : rb->limit_upper[j - (fl ? 7 : 0)];
   }
   /* constraints (mod.cpp:2466-2519 enabled masks, 2582-2613 registration: everyn_tsr and
    * con_tsr 'all' on every moving point, 'start' on the first, 'end' on the last) */
   if (pr->n_constraints > 0)
   {
      r->cons = (struct orc_con *) calloc(pr->n_constraints, sizeof(struct orc_con));
      if (!r->cons) { orc_run_destroy(r); return -1; }
      r->n_cons = pr->n_constraints;
      for (i = 0; i < pr->n_constraints; i++)
      {
         struct orc_con *con = &r->cons[i];
         con->r = r;
         con->spec = pr->constraints[i];
         con->k = 0;
         for (j = 0; j < 6; j++)
         {
            con->enabled[j] = (con->spec.Bw[j][0] == 0.0 && con->spec.Bw[j][1] == 0.0);
            con->k += con->enabled[j];
         }
         if (con->spec.link < 0 || con->spec.link >= rb->n_links) { orc_run_destroy(r); return -2; }
         switch (con->spec.where)
         {
         case OCB_CON_START_TSR: /* con_start_tsr on the start point itself (mod.cpp:2574) */
         case OCB_CON_START: cd_chomp_add_constraint(c, con->k, 0, con, orc_con_tsr); break;
         case OCB_CON_END: cd_chomp_add_constraint(c, con->k, m - 1, con, orc_con_tsr); break;
         case OCB_CON_ALL:
            for (j = 0; j < m; j++) cd_chomp_add_constraint(c, con->k, j, con, orc_con_tsr);
            break;
         default: orc_run_destroy(r); return -2;
         }
      }
   }
   if (cd_chomp_init(c)) { orc_run_destroy(r); return -2; }
   *out = r;
   return 0;
}

/* starttraj variant (mod.cpp:2373-2415 after sampling): overwrite all P rows */
void orc_run_set_traj(struct orc_run *r, const double *traj)
{
   memcpy(r->traj, traj, (size_t) r->n_points * r->n * sizeof(double));
}

/* mod::iterate loop (mod.cpp:2752-2831): optional HMC momentum resample, one
 * cd_chomp_iterate per iteration, then the cost-only pass.
 *   trace   [n_iter][3]    per-iteration (total, obs, smooth) as logged at 2798
 *   grads   [n_iter][m][n] c->G after each iteration (G/m + A T + B)
 *   costs   [3]            the [FINAL] triple (2830-2831)
 * returns 0, or -1 when the joint-limit loop gives up (exception at 2799-2803) */
int orc_run_iterate(struct orc_run *r, int n_iter, double *costs, double *trace, double *grads)
{
   struct cd_chomp *c = r->c;
   double total = 0, obs = 0, smooth = 0;
   int i, j;
   for (r->iter = 0; r->iter < n_iter; r->iter++)
   {
      int ret;
      if (r->use_hmc && r->iter == r->hmc_resample_iter)
      {
         double hmc_alpha = 100.0 * exp(0.02 * r->iter);
         for (i = 0; i < c->m; i++)
            for (j = 0; j < c->n; j++)
               c->AG[i * c->n + j] = orc_mt_gaussian(&r->rng, 1.0 / sqrt(hmc_alpha));
         c->leapfrog_first = 1;
         r->hmc_resample_iter += 1 + (int) (-log(orc_mt_uniform(&r->rng)) / r->hmc_resample_lambda);
      }
      ret = cd_chomp_iterate(c, 1, &total, &obs, &smooth);
      if (trace)
      {
         trace[3 * r->iter + 0] = total;
         trace[3 * r->iter + 1] = obs;
         trace[3 * r->iter + 2] = smooth;
      }
      if (grads) memcpy(&grads[(size_t) r->iter * c->m * c->n], c->G, (size_t) c->m * c->n * sizeof(double));
      if (ret == -1) return -1;
      if (r->floating) /* mod.cpp:2805-2808 */
         for (i = 0; i < r->n_points; i++) cd_kin_pose_normalize(&r->traj[i * c->n]);
   }
   cd_chomp_iterate(c, 0, &total, &obs, &smooth);
   if (costs) { costs[0] = total; costs[1] = obs; costs[2] = smooth; }
   return 0;
}

/* numeric part of gettraj (mod.cpp:2897-2903): the P x n waypoint rows */
void orc_run_get_traj(struct orc_run *r, double *traj)
{
   memcpy(traj, r->traj, (size_t) r->n_points * r->n * sizeof(double));
}

void orc_run_get_momentum(struct orc_run *r, double *AG)
{
   memcpy(AG, r->c->AG, (size_t) r->c->m * r->c->n * sizeof(double));
}

int orc_run_hmc_next(struct orc_run *r) { return r->hmc_resample_iter; }

/* raw obstacle gradient and cost of the current trajectory, without changing it:
 * the rows cost() writes (before the 1/m scaling and the smoothness term) */
int orc_run_obstacle_gradient(struct orc_run *r, double *grad /* m x n */, double *cost_rows /* m */)
{
   struct cd_chomp *c = r->c;
   int i;
   c->cost_pre(c->cptr, c, c->m, c->T_points);
   for (i = 0; i < c->m; i++)
      c->cost(c->cptr, c, i, c->T_points[i], 0, &cost_rows[i], &grad[i * c->n]);
   return 0;
}

/* sphere positions of the current trajectory: [P][n_active][3] */
int orc_run_sphere_positions(struct orc_run *r, double *out, int *n_active)
{
   struct cd_chomp *c = r->c;
   c->cost_pre(c->cptr, c, c->m, c->T_points);
   memcpy(out, r->sphere_poss_all, (size_t) r->n_points * r->n_spheres_active * 3 * sizeof(double));
   if (n_active) *n_active = r->n_spheres_active;
   return 0;
}

/* --------------------------------------------------------------- SDF commands */

/* core of addfield_fromobsarray (mod.cpp:693-714): wrap the array in a grid with
 * the given lengths and run cd_grid_double_bin_sdf.  obs is not consumed here. */
int orc_sdf_from_obsarray(const double *obs, const int sizes[3], const double lengths[3], double *sdf)
{
   struct cd_grid *g_obs = 0, *g_sdf = 0;
   double zero = 0.0;
   int sz[3], i, err;
   for (i = 0; i < 3; i++) sz[i] = sizes[i];
   if (cd_grid_create_sizearray(&g_obs, &zero, sizeof(double), 3, sz)) return -1;
   memcpy(g_obs->data, obs, g_obs->ncells * sizeof(double));
   for (i = 0; i < 3; i++) g_obs->lengths[i] = lengths[i];
   err = cd_grid_double_bin_sdf(&g_sdf, g_obs);
   if (!err)
   {
      memcpy(sdf, g_sdf->data, g_sdf->ncells * sizeof(double));
      cd_grid_destroy(g_sdf);
   }
   cd_grid_destroy(g_obs);
   return err;
}

int orc_dt_sqeuc(const double *func, const int sizes[3], const double lengths[3], double *out)
{
   struct cd_grid *g = 0, *g_dt = 0;
   double zero = 0.0;
   int sz[3], i, err;
   for (i = 0; i < 3; i++) sz[i] = sizes[i];
   if (cd_grid_create_sizearray(&g, &zero, sizeof(double), 3, sz)) return -1;
   memcpy(g->data, func, g->ncells * sizeof(double));
   for (i = 0; i < 3; i++) g->lengths[i] = lengths[i];
   err = cd_grid_double_dt_sqeuc(&g_dt, g);
   if (!err)
   {
      memcpy(out, g_dt->data, g_dt->ncells * sizeof(double));
      cd_grid_destroy(g_dt);
   }
   cd_grid_destroy(g);
   return err;
}

/* rotation matrix of a pose quaternion, row-major, each entry written as the
 * explicit product expansion used by kin.c:204-210 */
static void quat_to_rows(const double pose[7], double R[9])
{
   double qx = pose[3], qy = pose[4], qz = pose[5], qw = pose[6];
   double qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz, qw2 = qw * qw;
   double qxqy = qx * qy, qxqz = qx * qz, qxqw = qx * qw;
   double qyqz = qy * qz, qyqw = qy * qw, qzqw = qz * qw;
   R[0] = qx2 - qy2 - qz2 + qw2; R[1] = 2 * (qxqy - qzqw);      R[2] = 2 * (qxqz + qyqw);
   R[3] = 2 * (qxqy + qzqw);     R[4] = -qx2 + qy2 - qz2 + qw2; R[5] = 2 * (qyqz - qxqw);
   R[6] = 2 * (qxqz - qyqw);     R[7] = 2 * (qyqz + qxqw);      R[8] = -qx2 - qy2 + qz2 + qw2;
}

/* analytic stand-in for CheckCollision(cube) (mod.cpp:520); the definition is
 * in include/orcdchomp_b200.h.  c = cube centre (grid frame), h = half extent.
 * Compiled with -ffp-contract=off so every product/sum rounds once, matching
 * the device code's explicit __dmul_rn/__dadd_rn sequence bit for bit. */
/* cube (centre c, half extent h) against a triangle: separating-axis test over the 9 edge
 * cross products, the 3 cube face normals and the triangle plane (Akenine-Moller 2001);
 * strict inequalities, so touching is a hit.  Straight-line arithmetic, same order as the
 * device code. */
static int cube_hits_triangle(const double c[3], double h, const double V[9])
{
   double v[3][3], e[3][3], n[3], vmin[3], vmax[3];
   int i, k;
   for (i = 0; i < 3; i++)
      for (k = 0; k < 3; k++) v[i][k] = V[3 * i + k] - c[k];
   for (k = 0; k < 3; k++)
   {
      e[0][k] = v[1][k] - v[0][k];
      e[1][k] = v[2][k] - v[1][k];
      e[2][k] = v[0][k] - v[2][k];
   }
   for (i = 0; i < 3; i++)
   {
      const double ex = e[i][0], ey = e[i][1], ez = e[i][2];
      const double fx = fabs(ex), fy = fabs(ey), fz = fabs(ez);
      double p0, p1, p2, lo, hi, rad;
      /* axis x cross e */
      p0 = ey * v[0][2] - ez * v[0][1]; p1 = ey * v[1][2] - ez * v[1][1]; p2 = ey * v[2][2] - ez * v[2][1];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fz * h + fy * h;
      if (lo > rad || hi < -rad) return 0;
      /* axis y cross e */
      p0 = ez * v[0][0] - ex * v[0][2]; p1 = ez * v[1][0] - ex * v[1][2]; p2 = ez * v[2][0] - ex * v[2][2];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fz * h + fx * h;
      if (lo > rad || hi < -rad) return 0;
      /* axis z cross e */
      p0 = ex * v[0][1] - ey * v[0][0]; p1 = ex * v[1][1] - ey * v[1][0]; p2 = ex * v[2][1] - ey * v[2][0];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fy * h + fx * h;
      if (lo > rad || hi < -rad) return 0;
   }
   for (k = 0; k < 3; k++)
   {
      double lo = v[0][k], hi = v[0][k];
      if (v[1][k] < lo) lo = v[1][k]; if (v[1][k] > hi) hi = v[1][k];
      if (v[2][k] < lo) lo = v[2][k]; if (v[2][k] > hi) hi = v[2][k];
      if (lo > h || hi < -h) return 0;
   }
   n[0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
   n[1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
   n[2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
   for (k = 0; k < 3; k++)
   {
      if (n[k] > 0.0) { vmin[k] = -h - v[0][k]; vmax[k] = h - v[0][k]; }
      else { vmin[k] = h - v[0][k]; vmax[k] = -h - v[0][k]; }
   }
   if (n[0] * vmin[0] + n[1] * vmin[1] + n[2] * vmin[2] > 0.0) return 0;
   if (n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2] >= 0.0) return 1;
   return 0;
}

static int cube_hits_prim(const double c[3], double h, const struct ocb_prim *p)
{
   if (p->type == OCB_PRIM_TRIANGLE)
   {
      double V[9];
      int k;
      for (k = 0; k < 7; k++) V[k] = p->pose[k];
      V[7] = p->extents[0];
      V[8] = p->extents[1];
      return cube_hits_triangle(c, h, V);
   }
   if (p->type == OCB_PRIM_SPHERE)
   {
      double d2 = 0.0, r = p->extents[0];
      int k;
      for (k = 0; k < 3; k++)
      {
         double d = fabs(p->pose[k] - c[k]) - h;
         if (d < 0.0) d = 0.0;
         d2 = d2 + d * d;
      }
      return d2 <= r * r;
   }
   else
   {
      /* separating-axis test: cube (axes = grid axes) against an oriented box */
      double R[9], A[9], t[3], ra, rb, tl;
      const double *e = p->extents;
      int i, j;
      quat_to_rows(p->pose, R); /* columns of R = box axes in the grid frame */
      for (i = 0; i < 9; i++) A[i] = fabs(R[i]) + 1e-12;
      for (i = 0; i < 3; i++) t[i] = p->pose[i] - c[i];
      /* cube face normals */
      for (i = 0; i < 3; i++)
      {
         ra = h;
         rb = e[0] * A[3 * i + 0] + e[1] * A[3 * i + 1] + e[2] * A[3 * i + 2];
         if (fabs(t[i]) > ra + rb) return 0;
      }
      /* box face normals */
      for (j = 0; j < 3; j++)
      {
         ra = h * A[0 + j] + h * A[3 + j] + h * A[6 + j];
         rb = e[j];
         tl = t[0] * R[0 + j] + t[1] * R[3 + j] + t[2] * R[6 + j];
         if (fabs(tl) > ra + rb) return 0;
      }
      /* edge x edge */
      for (i = 0; i < 3; i++)
      {
         int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
         for (j = 0; j < 3; j++)
         {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            ra = h * A[3 * i1 + j] + h * A[3 * i2 + j];
            rb = e[j1] * A[3 * i + j2] + e[j2] * A[3 * i + j1];
            tl = t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j];
            if (fabs(tl) > ra + rb) return 0;
         }
      }
      return 1;
   }
}

static int replace_one_by_zero(void *cell, void *rptr)
{
   double *v = (double *) cell;
   (void) rptr;
   if (*v == 1.0) { *v = 0.0; return 1; }
   return 0;
}

/* computedistancefield after the AABB sizing (mod.cpp:398-403 grid of 1.0 with
 * lengths, 498-525 occupancy loop, 543-548 flood fill from voxel 0 + relabel,
 * 560 bin_sdf).  obs_out (may be NULL) receives the grid handed to bin_sdf. */
int orc_computedistancefield(const struct ocb_prim *prims, int n_prims, const int sizes[3],
                             const double lengths[3], double cube_extent,
                             double *obs_out, double *sdf_out)
{
   struct cd_grid *g_obs = 0, *g_sdf = 0;
   double one = 1.0;
   int sz[3], i, err = 0;
   size_t idx;
   for (i = 0; i < 3; i++) sz[i] = sizes[i];
   if (cd_grid_create_sizearray(&g_obs, &one, sizeof(double), 3, sz)) return -1;
   for (i = 0; i < 3; i++) g_obs->lengths[i] = lengths[i];
   for (idx = 0; idx < g_obs->ncells; idx++)
   {
      double center[3];
      cd_grid_center_index(g_obs, idx, center);
      for (i = 0; i < n_prims; i++)
         if (cube_hits_prim(center, cube_extent, &prims[i]))
         {
            *(double *) cd_grid_get_index(g_obs, idx) = HUGE_VAL;
            break;
         }
   }
   cd_grid_flood_fill(g_obs, 0, 0, replace_one_by_zero, 0);
   for (idx = 0; idx < g_obs->ncells; idx++)
      if (*(double *) cd_grid_get_index(g_obs, idx) == 1.0)
         *(double *) cd_grid_get_index(g_obs, idx) = HUGE_VAL;
   if (obs_out) memcpy(obs_out, g_obs->data, g_obs->ncells * sizeof(double));
   if (sdf_out)
   {
      err = cd_grid_double_bin_sdf(&g_sdf, g_obs);
      if (!err)
      {
         memcpy(sdf_out, g_sdf->data, g_sdf->ncells * sizeof(double));
         cd_grid_destroy(g_sdf);
      }
   }
   cd_grid_destroy(g_obs);
   return err;
}

/* occupancy only (before the flood fill): 1.0 free / HUGE_VAL hit */
int orc_occupancy(const struct ocb_prim *prims, int n_prims, const int sizes[3],
                  const double lengths[3], double cube_extent, double *grid_out)
{
   struct cd_grid g;
   int sz[3], i;
   size_t idx;
   double len[3];
   for (i = 0; i < 3; i++) { sz[i] = sizes[i]; len[i] = lengths[i]; }
   g.n = 3; g.sizes = sz; g.lengths = len; g.cell_size = sizeof(double);
   g.ncells = (size_t) sz[0] * sz[1] * sz[2];
   g.data = (char *) grid_out;
   for (idx = 0; idx < g.ncells; idx++)
   {
      double center[3];
      grid_out[idx] = 1.0;
      cd_grid_center_index(&g, idx, center);
      for (i = 0; i < n_prims; i++)
         if (cube_hits_prim(center, cube_extent, &prims[i])) { grid_out[idx] = HUGE_VAL; break; }
   }
   return 0;
}

/* SDF sampling hooks for known-answer tests of interp / grad */
int orc_sdf_sample(const double *data, const int sizes[3], const double lengths[3],
                   const double *points, int n_points, double *values, double *grads, int *errs)
{
   struct cd_grid g;
   int sz[3], i;
   double len[3];
   for (i = 0; i < 3; i++) { sz[i] = sizes[i]; len[i] = lengths[i]; }
   g.n = 3; g.sizes = sz; g.lengths = len; g.cell_size = sizeof(double);
   g.ncells = (size_t) sz[0] * sz[1] * sz[2];
   g.data = (char *) data;
   for (i = 0; i < n_points; i++)
   {
      double p[3];
      p[0] = points[3 * i]; p[1] = points[3 * i + 1]; p[2] = points[3 * i + 2];
      values[i] = 0.0;
      grads[3 * i] = grads[3 * i + 1] = grads[3 * i + 2] = 0.0;
      errs[i] = cd_grid_double_interp(&g, p, &values[i]);
      if (!errs[i]) cd_grid_double_grad(&g, p, &grads[3 * i]);
   }
   return 0;
}

const char *orc_flavour(void)
{
#ifdef ORACLE_REF_HEADERS
   return "reference";
#else
   return "port";
#endif
}
