/* libcd_compat.cpp -- libcd-named, libcd-laid-out entry points over the engine's SDF build
 * (include/libcd_b200.h).  Host code only: every numeric step is a call into
 * liborcdchomp_b200.so, which runs it on the GPU or fails. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>

#include "../../include/libcd_b200.h"
#include "../../include/orcdchomp_b200.h"

namespace
{
std::mutex g_lock;
ocb_engine *g_engine = nullptr;
int g_engine_device = -1;
int g_device = -1;
thread_local char g_err[512] = "";

int set_err(int code, const char *what)
{
   snprintf(g_err, sizeof(g_err), "%s", what);
   return code;
}

/* the process-wide engine behind the handle-less libcd signatures */
ocb_engine *engine()
{
   std::lock_guard<std::mutex> guard(g_lock);
   int want = g_device;
   if (want < 0)
   {
      const char *env = getenv("OCB_DEVICE");
      want = env ? atoi(env) : 0;
   }
   if (g_engine && g_engine_device != want)
   {
      ocb_engine_destroy(g_engine);
      g_engine = nullptr;
   }
   if (!g_engine)
   {
      if (ocb_engine_create(want, &g_engine) != OCB_OK)
      {
         g_engine = nullptr;
         return nullptr;
      }
      g_engine_device = want;
   }
   return g_engine;
}

/* a new grid shaped like src with uninitialised cells; four malloc blocks as grid.c:99-132
 * so that libcd's cd_grid_destroy releases it */
struct cd_grid *grid_like(const struct cd_grid *src)
{
   struct cd_grid *g = (struct cd_grid *) malloc(sizeof(struct cd_grid));
   if (!g) return nullptr;
   g->n = src->n;
   g->ncells = src->ncells;
   g->cell_size = src->cell_size;
   g->sizes = (int *) malloc(src->n * sizeof(int));
   g->lengths = (double *) malloc(src->n * sizeof(double));
   g->data = (char *) malloc(src->ncells * (size_t) src->cell_size);
   if (!g->sizes || !g->lengths || !g->data)
   {
      free(g->sizes);
      free(g->lengths);
      free(g->data);
      free(g);
      return nullptr;
   }
   memcpy(g->sizes, src->sizes, src->n * sizeof(int));
   memcpy(g->lengths, src->lengths, src->n * sizeof(double));
   return g;
}

void grid_free(struct cd_grid *g)
{
   if (!g) return;
   free(g->data);
   free(g->sizes);
   free(g->lengths);
   free(g);
}

int check(const struct cd_grid *g)
{
   if (!g || !g->sizes || !g->lengths || !g->data) return set_err(-2, "null grid");
   if (g->cell_size != (int) sizeof(double)) return set_err(-2, "cell type is not double");
   if (g->n != 3) return set_err(-2, "only 3-dimensional grids are accelerated");
   if (g->ncells != (size_t) g->sizes[0] * g->sizes[1] * g->sizes[2]) return set_err(-2, "ncells does not match sizes");
   return 0;
}

int from_ocb(int rc)
{
   if (rc == OCB_OK) return 0;
   snprintf(g_err, sizeof(g_err), "%s", ocb_last_error());
   return rc == OCB_ERR_ALLOC ? -1 : (rc == OCB_ERR_ARG ? -2 : -3);
}

typedef int (*host_transform)(ocb_engine *, const double *, const int *, const double *, double *);

int transform(struct cd_grid **gp_out, struct cd_grid *g_in, host_transform fn)
{
   int rc = check(g_in);
   if (rc) return rc;
   if (!gp_out) return set_err(-2, "null output pointer");
   ocb_engine *e = engine();
   if (!e) return from_ocb(OCB_ERR_NODEVICE);
   struct cd_grid *out = grid_like(g_in);
   if (!out) return set_err(-1, "out of host memory");
   rc = from_ocb(fn(e, (const double *) g_in->data, g_in->sizes, g_in->lengths, (double *) out->data));
   if (rc)
   {
      grid_free(out);
      return rc;
   }
   *gp_out = out;
   return 0;
}
} /* namespace */

extern "C" int cd_grid_double_bin_sdf(struct cd_grid **gp_dt, struct cd_grid *g_emp)
{
   return transform(gp_dt, g_emp, ocb_sdf_build_host);
}

extern "C" int cd_grid_double_dt_sqeuc(struct cd_grid **gp_dt, struct cd_grid *g_func)
{
   return transform(gp_dt, g_func, ocb_dt_sqeuc_host);
}

extern "C" int cd_grid_double_sedt(struct cd_grid **gp_dt, struct cd_grid *g_func)
{
   return transform(gp_dt, g_func, ocb_dt_sqeuc_host);
}

extern "C" int cd_grid_b200_flood_relabel(struct cd_grid *g, size_t index_start)
{
   int rc = check(g);
   if (rc) return rc;
   ocb_engine *e = engine();
   if (!e) return from_ocb(OCB_ERR_NODEVICE);
   return from_ocb(ocb_flood_relabel_host(e, (double *) g->data, g->sizes, index_start));
}

extern "C" int cd_grid_b200_set_device(int device)
{
   if (device < 0) return set_err(-2, "bad device");
   std::lock_guard<std::mutex> guard(g_lock);
   g_device = device;
   return 0;
}

extern "C" const char *cd_grid_b200_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------- */
/* cd_chomp facade: one run of the batched engine behind libcd's struct       */
#include <vector>

namespace
{
struct ChompPrivate
{
   /* deep copy of what cd_chomp_b200_set_sphere_cost was given */
   bool have_cost = false;
   ocb_robot robot;
   ocb_params params;
   std::vector<int> parent, joint_type, dof_index, sphere_link;
   std::vector<double> pose_parent, axis, dof_coeff, sphere_pos, sphere_radius;
   std::vector<ocb_sdf> sdfs;
   std::vector<std::vector<double>> sdf_data;
   std::vector<ocb_constraint> constraints;
   /* engine side, made by cd_chomp_init */
   ocb_engine *e = nullptr;
   ocb_batch *b = nullptr;
   std::vector<int> sdf_ids;
   std::vector<double> traj; /* [m + 2][n] staging */
};

struct ChompHolder
{
   struct cd_chomp c; /* first member: a cd_chomp* is a ChompHolder* */
   ChompPrivate *p;
};

ChompPrivate *priv(struct cd_chomp *c) { return reinterpret_cast<ChompHolder *>(c)->p; }

void release_engine_side(ChompPrivate *p)
{
   if (p->b) ocb_batch_destroy(p->b);
   p->b = nullptr;
   if (p->e)
      for (int id : p->sdf_ids) ocb_sdf_remove(p->e, id);
   p->sdf_ids.clear();
}
} /* namespace */

extern "C" int cd_chomp_create(struct cd_chomp **cp, int m, int n, int D, double *T, int ldt)
{
   if (!cp || m < 1 || n < 1 || D < 1 || !T || ldt < n) return set_err(-2, "bad argument");
   ChompHolder *h = (ChompHolder *) calloc(1, sizeof(ChompHolder));
   if (!h) return set_err(-1, "out of host memory");
   h->p = new (std::nothrow) ChompPrivate();
   struct cd_chomp *c = &h->c;
   /* defaults of chomp.c:40-178 */
   c->n = n;
   c->m = m;
   c->D = D;
   c->lambda = 1.0;
   c->dt = 1.0 / (m + 1);
   c->T = T;
   c->ldt = ldt;
   c->T_points = (double **) malloc(m * sizeof(double *));
   c->G = (double *) calloc((size_t) m * n, sizeof(double));
   c->G_points = (double **) malloc(m * sizeof(double *));
   c->AG = (double *) calloc((size_t) m * n, sizeof(double)); /* chomp.c:114-115 */
   c->AG_points = (double **) malloc(m * sizeof(double *));
   c->wds = (double *) malloc(D * sizeof(double));
   c->initsfinals = (double *) calloc((size_t) 2 * D * n, sizeof(double));
   c->inits = (double **) malloc(D * sizeof(double *));
   c->finals = (double **) malloc(D * sizeof(double *));
   c->jlimit_lower = (double *) malloc(n * sizeof(double));
   c->jlimit_upper = (double *) malloc(n * sizeof(double));
   if (!h->p || !c->T_points || !c->G || !c->G_points || !c->AG || !c->AG_points || !c->wds || !c->initsfinals ||
       !c->inits || !c->finals || !c->jlimit_lower || !c->jlimit_upper)
   {
      cd_chomp_free(c);
      return set_err(-1, "out of host memory");
   }
   for (int i = 0; i < m; i++)
   {
      c->T_points[i] = T + (size_t) i * ldt;
      c->G_points[i] = c->G + (size_t) i * n;
      c->AG_points[i] = c->AG + (size_t) i * n;
   }
   for (int d = 0; d < D; d++)
   {
      c->wds[d] = (d < D - 1) ? 0.0 : 1.0;               /* chomp.c:127-128 */
      c->inits[d] = c->initsfinals + (size_t) (2 * d) * n;      /* zero vectors, interleaved as chomp.c:131-141 */
      c->finals[d] = c->initsfinals + (size_t) (2 * d + 1) * n;
   }
   for (int j = 0; j < n; j++)
   {
      c->jlimit_lower[j] = -HUGE_VAL;                     /* chomp.c:165-169 */
      c->jlimit_upper[j] = HUGE_VAL;
   }
   c->leapfrog_first = 1;                                 /* chomp.c:88 */
   *cp = c;
   return 0;
}

extern "C" void cd_chomp_free(struct cd_chomp *c)
{
   if (!c) return;
   ChompPrivate *p = priv(c);
   if (p)
   {
      release_engine_side(p);
      delete p;
   }
   free(c->T_points);
   free(c->G);
   free(c->G_points);
   free(c->AG);
   free(c->AG_points);
   free(c->wds);
   free(c->initsfinals);
   free(c->inits);
   free(c->finals);
   free(c->jlimit_lower);
   free(c->jlimit_upper);
   free(c);
}

extern "C" int cd_chomp_b200_set_sphere_cost(struct cd_chomp *c, const ocb_robot *robot, const ocb_params *params,
                                             int n_sdfs, const ocb_sdf *sdfs)
{
   if (!c || !robot || !params || n_sdfs < 1 || !sdfs) return set_err(-2, "bad argument");
   if (robot->n_dof + (params->floating_base ? 7 : 0) != c->n) return set_err(-2, "robot active dofs differ from the run's n");
   ChompPrivate *p = priv(c);
   const int nl = robot->n_links, ns = robot->n_spheres;
   p->parent.assign(robot->parent, robot->parent + nl);
   p->pose_parent.assign(robot->pose_parent, robot->pose_parent + 7 * (size_t) nl);
   p->joint_type.assign(robot->joint_type, robot->joint_type + nl);
   p->axis.assign(robot->axis, robot->axis + 3 * (size_t) nl);
   p->dof_index.assign(robot->dof_index, robot->dof_index + nl);
   p->dof_coeff.assign(robot->dof_coeff, robot->dof_coeff + 2 * (size_t) nl);
   p->sphere_link.assign(robot->sphere_link, robot->sphere_link + ns);
   p->sphere_pos.assign(robot->sphere_pos, robot->sphere_pos + 3 * (size_t) ns);
   p->sphere_radius.assign(robot->sphere_radius, robot->sphere_radius + ns);
   p->robot = *robot;
   p->robot.parent = p->parent.data();
   p->robot.pose_parent = p->pose_parent.data();
   p->robot.joint_type = p->joint_type.data();
   p->robot.axis = p->axis.data();
   p->robot.dof_index = p->dof_index.data();
   p->robot.dof_coeff = p->dof_coeff.data();
   p->robot.sphere_link = p->sphere_link.data();
   p->robot.sphere_pos = p->sphere_pos.data();
   p->robot.sphere_radius = p->sphere_radius.data();
   p->params = *params;
   /* hard constraints travel as data (ocb_constraint), not as con_eval callbacks; start_tsr would change
    * which rows T holds and is only offered by the batch interface */
   p->constraints.clear();
   for (int i = 0; i < params->n_constraints; i++)
   {
      if (!params->constraints) return set_err(-2, "constraints is null");
      if (params->constraints[i].where == OCB_CON_START_TSR) return set_err(-2, "start_tsr is not offered by the cd_chomp facade");
      p->constraints.push_back(params->constraints[i]);
   }
   p->params.n_constraints = (int) p->constraints.size();
   p->params.constraints = p->constraints.empty() ? nullptr : p->constraints.data();
   p->sdfs.assign(sdfs, sdfs + n_sdfs);
   p->sdf_data.resize(n_sdfs);
   for (int i = 0; i < n_sdfs; i++)
   {
      const size_t cells = (size_t) sdfs[i].sizes[0] * sdfs[i].sizes[1] * sdfs[i].sizes[2];
      if (!sdfs[i].data || !cells) return set_err(-2, "empty signed distance field");
      p->sdf_data[i].assign(sdfs[i].data, sdfs[i].data + cells);
      p->sdfs[i].data = p->sdf_data[i].data();
   }
   p->have_cost = true;
   return 0;
}

extern "C" int cd_chomp_init(struct cd_chomp *c)
{
   if (!c) return set_err(-2, "null run");
   ChompPrivate *p = priv(c);
   if (!p->have_cost) return set_err(-2, "no cost attached: call cd_chomp_b200_set_sphere_cost before cd_chomp_init");
   if (c->cost_pre || c->cost || c->cost_extra || c->cons)
      return set_err(-2, "host callbacks (cost, con_eval) cannot run on the device: pass constraints as ocb_constraint in the params of cd_chomp_b200_set_sphere_cost");
   /* the engine's metric is the module's: D-th derivative only, fixed end points, uniform dt */
   for (int d = 0; d < c->D; d++)
      if (c->wds[d] != ((d < c->D - 1) ? 0.0 : 1.0)) return set_err(-2, "wds other than [0..0,1]");
   if (fabs(c->dt * (c->m + 1) - 1.0) > 1e-12) return set_err(-2, "dt other than 1/(m+1)");
   if (!c->inits[0] || !c->finals[0]) return set_err(-2, "free end points are not supported");
   for (int d = 1; d < c->D; d++)
      for (int j = 0; j < c->n; j++)
         if ((c->inits[d] && c->inits[d][j] != 0.0) || (c->finals[d] && c->finals[d][j] != 0.0))
            return set_err(-2, "end-point derivatives other than zero are not supported");
   release_engine_side(p);
   p->e = engine();
   if (!p->e) return from_ocb(OCB_ERR_NODEVICE);
   for (size_t i = 0; i < p->sdfs.size(); i++)
   {
      int id = -1;
      int rc = ocb_sdf_upload(p->e, &p->sdfs[i], &id);
      if (rc)
      {
         release_engine_side(p);
         return from_ocb(rc);
      }
      p->sdf_ids.push_back(id);
   }
   ocb_robot rb = p->robot;
   /* with a floating base the first seven entries of a row are the pose (unbounded, mod.cpp:2640-2652) */
   rb.limit_lower = c->jlimit_lower + (p->params.floating_base ? 7 : 0);
   rb.limit_upper = c->jlimit_upper + (p->params.floating_base ? 7 : 0);
   ocb_params pr = p->params;
   pr.n_points = c->m + 2;
   pr.derivative = c->D;
   pr.lambda = c->lambda;
   pr.use_momentum = c->use_momentum ? 1 : 0;
   pr.use_hmc = 0; /* the caller resamples AG itself, as the module does (mod.cpp:2755-2768) */
   int rc = ocb_batch_create(p->e, &rb, &pr, (int) p->sdf_ids.size(), p->sdf_ids.data(), 1, c->inits[0],
                             c->finals[0], nullptr, &p->b);
   if (rc)
   {
      p->b = nullptr;
      release_engine_side(p);
      return from_ocb(rc);
   }
   ocb_batch_enable_trace(p->b, 1);
   ocb_batch_capture_gradient(p->b, 1);
   p->traj.assign((size_t) (c->m + 2) * c->n, 0.0);
   return 0;
}

extern "C" int cd_chomp_iterate(struct cd_chomp *c, int do_iteration, double *costp_total, double *costp_obs,
                                double *costp_smooth)
{
   if (!c) return set_err(-2, "null run");
   ChompPrivate *p = priv(c);
   if (!p->b) return set_err(-2, "cd_chomp_init has not succeeded");
   const int m = c->m, n = c->n;
   /* the caller's rows (leading dimension ldt) between the fixed end points */
   memcpy(&p->traj[0], c->inits[0], n * sizeof(double));
   for (int i = 0; i < m; i++) memcpy(&p->traj[(size_t) (i + 1) * n], c->T + (size_t) i * c->ldt, n * sizeof(double));
   memcpy(&p->traj[(size_t) (m + 1) * n], c->finals[0], n * sizeof(double));
   int rc = ocb_batch_set_traj(p->b, p->traj.data());
   if (rc == OCB_OK) rc = ocb_batch_set_lambda(p->b, c->lambda);
   if (rc == OCB_OK && c->use_momentum) rc = ocb_batch_set_momentum(p->b, c->AG, &c->leapfrog_first);
   if (rc) return from_ocb(rc);
   double total = 0.0, obs = 0.0, smooth = 0.0;
   int status = 0;
   rc = ocb_batch_iterate(p->b, do_iteration ? 1 : 0, &total, &obs, &smooth, &status);
   if (rc) return from_ocb(rc);
   if (do_iteration)
   {
      rc = ocb_batch_get_traj(p->b, p->traj.data());
      if (rc == OCB_OK) rc = ocb_batch_get_gradient(p->b, c->G);
      if (rc == OCB_OK && c->use_momentum) rc = ocb_batch_get_momentum(p->b, c->AG, &c->leapfrog_first);
      if (rc) return from_ocb(rc);
      for (int i = 0; i < m; i++) memcpy(c->T + (size_t) i * c->ldt, &p->traj[(size_t) (i + 1) * n], n * sizeof(double));
      if (status == OCB_ERR_JLIMIT) return set_err(-1, "joint-limit projection did not converge"); /* chomp.c:651-655 */
      /* obstacle cost of the trajectory before the step, smoothness after it (chomp.c:493-511, 660-681) */
      double tr[3];
      rc = ocb_batch_get_trace(p->b, tr, 1);
      if (rc) return from_ocb(rc);
      total = tr[0]; obs = tr[1]; smooth = tr[2];
   }
   if (costp_total) *costp_total = total;
   if (costp_obs) *costp_obs = obs;
   if (costp_smooth) *costp_smooth = smooth;
   return 0;
}
