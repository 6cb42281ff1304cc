/* libcd_compat.cpp -- libcd-named, libcd-laid-out entry points over the engine's SDF build
 * (include/libcd_b200.h).  Host code only: every numeric step is a call into
 * liborcdchomp_b200.so, which runs it on the GPU or fails. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

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
