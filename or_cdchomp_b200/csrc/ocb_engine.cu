/* ocb_engine.cu -- host side of the C ABI (include/orcdchomp_b200.h).
 *
 * Owns device memory, compiles the robot description into the joint-frame form
 * the kernels consume, builds the smoothness metric (band of A, its LDL^T
 * factor, the B / trC coefficients) and launches the sm_100a kernels.  No numeric
 * hot-path work happens on the host: everything per-iteration is in
 * chomp_kernel.cu, everything per-voxel in sdf_kernels.cu.
 *
 * Reference counterparts (paths relative to the reference root):
 *   numeric part of mod::create          src/orcdchomp_mod.cpp:2266-2299, 2315-2369,
 *                                         2417-2464, 2521, 2567-2580, 2617-2664
 *   cd_chomp_create / add_KEs / init     src/libcd/chomp.c:40-178, 239-340, 342-428
 *   mod::iterate / gettraj / destroy     src/orcdchomp_mod.cpp:2690-2852, 2897-2903, 3039-3066
 *   computedistancefield / addfield      src/orcdchomp_mod.cpp:297-589, 592-722
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/orcdchomp_b200.h"
#include "ocb_internal.h"

/* ------------------------------------------------------------------ errors */
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof(g_err), fmt, ap);
   va_end(ap);
   return code;
}

#define CU(call)                                                                              \
   do                                                                                         \
   {                                                                                          \
      cudaError_t e__ = (call);                                                               \
      if (e__ != cudaSuccess)                                                                 \
         return fail(OCB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
   } while (0)

extern "C" const char *ocb_last_error(void) { return g_err; }
extern "C" const char *ocb_version(void) { return "or_cdchomp_b200 0.1 (sm_100a)"; }

extern "C" void ocb_params_default(ocb_params *p)
{
   /* src/orcdchomp_mod.cpp:1818-1848, 1875 */
   p->n_points = 101;
   p->derivative = 1;
   p->lambda = 10.0;
   p->use_momentum = 0;
   p->use_hmc = 0;
   p->hmc_resample_lambda = 0.02;
   p->epsilon = 0.1;
   p->epsilon_self = 0.04;
   p->obs_factor = 200.0;
   p->obs_factor_self = 10.0;
   p->floating_base = 0;
   p->n_constraints = 0;
   p->constraints = nullptr;
}

/* ------------------------------------------------------------ small algebra */
namespace
{
struct Xf
{
   double R[9]; /* row-major */
   double t[3];
};

Xf xf_identity()
{
   Xf x;
   for (int i = 0; i < 9; i++) x.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
   x.t[0] = x.t[1] = x.t[2] = 0.0;
   return x;
}

/* rotation matrix of a libcd pose quaternion; entries are the expansion terms of
 * cd_kin_pose_compos (src/libcd/kin.c:192-210) so R p reproduces its products */
void quat_to_R(const double *q, double *R)
{
   const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
   const double qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz, qw2 = qw * qw;
   const double qxqy = qx * qy, qxqz = qx * qz, qxqw = qx * qw;
   const double qyqz = qy * qz, qyqw = qy * qw, qzqw = qz * qw;
   R[0] = qx2 - qy2 - qz2 + qw2; R[1] = 2 * (qxqy - qzqw);      R[2] = 2 * (qxqz + qyqw);
   R[3] = 2 * (qxqy + qzqw);     R[4] = -qx2 + qy2 - qz2 + qw2; R[5] = 2 * (qyqz - qxqw);
   R[6] = 2 * (qxqz - qyqw);     R[7] = 2 * (qyqz + qxqw);      R[8] = -qx2 - qy2 + qz2 + qw2;
}

Xf xf_from_pose(const double *pose)
{
   Xf x;
   quat_to_R(pose + 3, x.R);
   x.t[0] = pose[0]; x.t[1] = pose[1]; x.t[2] = pose[2];
   return x;
}

Xf xf_mul(const Xf &a, const Xf &b)
{
   Xf c;
   for (int r = 0; r < 3; r++)
   {
      for (int k = 0; k < 3; k++)
         c.R[3 * r + k] = a.R[3 * r] * b.R[k] + a.R[3 * r + 1] * b.R[3 + k] + a.R[3 * r + 2] * b.R[6 + k];
      c.t[r] = a.R[3 * r] * b.t[0] + a.R[3 * r + 1] * b.t[1] + a.R[3 * r + 2] * b.t[2] + a.t[r];
   }
   return c;
}

void xf_apply(const Xf &a, const double *p, double *out)
{
   for (int r = 0; r < 3; r++) out[r] = a.R[3 * r] * p[0] + a.R[3 * r + 1] * p[1] + a.R[3 * r + 2] * p[2] + a.t[r];
}

/* rotation Q with Q e_z = axis (unit) */
Xf xf_align_z(const double *axis)
{
   double a[3] = {axis[0], axis[1], axis[2]};
   const double len = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
   for (int i = 0; i < 3; i++) a[i] /= len;
   Xf q = xf_identity();
   if (fabs(a[0]) < 1e-15 && fabs(a[1]) < 1e-15 && a[2] > 0) return q;
   double h[3] = {0, 0, 0};
   int smallest = 0;
   if (fabs(a[1]) < fabs(a[smallest])) smallest = 1;
   if (fabs(a[2]) < fabs(a[smallest])) smallest = 2;
   h[smallest] = 1.0;
   double u[3] = {h[1] * a[2] - h[2] * a[1], h[2] * a[0] - h[0] * a[2], h[0] * a[1] - h[1] * a[0]};
   const double ul = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
   for (int i = 0; i < 3; i++) u[i] /= ul;
   double w[3] = {a[1] * u[2] - a[2] * u[1], a[2] * u[0] - a[0] * u[2], a[0] * u[1] - a[1] * u[0]};
   for (int r = 0; r < 3; r++)
   {
      q.R[3 * r] = u[r];
      q.R[3 * r + 1] = w[r];
      q.R[3 * r + 2] = a[r];
   }
   return q;
}

Xf xf_transpose_rot(const Xf &a)
{
   Xf c = xf_identity();
   for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) c.R[3 * r + k] = a.R[3 * k + r];
   return c;
}

Xf xf_motion(int type, const double *axis, double value)
{
   Xf x = xf_identity();
   if (type == OCB_JOINT_REVOLUTE)
   {
      const double s = sin(0.5 * value), c = cos(0.5 * value);
      const double q[4] = {axis[0] * s, axis[1] * s, axis[2] * s, c};
      quat_to_R(q, x.R);
   }
   else if (type == OCB_JOINT_PRISMATIC)
      for (int i = 0; i < 3; i++) x.t[i] = axis[i] * value;
   return x;
}

template <class T>
int dev_upload(T **dptr, const std::vector<T> &h)
{
   *dptr = nullptr;
   const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
   if (cudaMalloc((void **) dptr, bytes) != cudaSuccess) return -1;
   if (!h.empty() && cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
      return -1;
   return 0;
}
} /* namespace */

/* ------------------------------------------------------------------ engine */
struct SdfSlot
{
   bool used = false;
   bool owned = false;
   double *d_data = nullptr;
   int sizes[3] = {0, 0, 0};
   double lengths[3] = {0, 0, 0};
   double pose[7] = {0, 0, 0, 0, 0, 0, 1};
};

struct ocb_engine
{
   int device = 0;
   cudaStream_t own_stream = nullptr;
   cudaStream_t stream = nullptr;
   std::vector<SdfSlot> sdfs;
   void *scratch = nullptr;
   size_t scratch_bytes = 0;
   long launches = 0;
   int smem_optin = 0;
   int smem_per_sm = 0;
   int sm_count = 0;
   int force_general_sdf = 0; /* test hook: always take the general fp64 distance transform */
   int jit = 0;               /* compile the persistent kernel per batch configuration (ocb_jit.cpp) */
   /* every device buffer of batches, resident SDFs and host-call temporaries comes from this
    * stream-ordered pool, which keeps what it is given back: create / destroy of a batch costs
    * microseconds and never synchronises the device (cudaMalloc / cudaFree took 1 - 300 ms per
    * batch in the end-to-end path) */
   cudaMemPool_t pool = nullptr;
};

static cudaError_t pool_alloc(ocb_engine *e, void **p, size_t bytes)
{
   *p = nullptr;
   return cudaMallocFromPoolAsync(p, std::max<size_t>(bytes, 1), e->pool, e->stream);
}

static void pool_free(ocb_engine *e, void *p)
{
   if (p) cudaFreeAsync(p, e->stream);
}

static int engine_scratch(ocb_engine *e, size_t bytes)
{
   if (bytes <= e->scratch_bytes) return OCB_OK;
   if (e->scratch) cudaFree(e->scratch);
   e->scratch = nullptr;
   e->scratch_bytes = 0;
   if (cudaMalloc(&e->scratch, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      return fail(OCB_ERR_ALLOC, "cudaMalloc of %zu scratch bytes failed", bytes);
   }
   e->scratch_bytes = bytes;
   return OCB_OK;
}

extern "C" int ocb_engine_create(int device, ocb_engine **out)
{
   if (!out) return fail(OCB_ERR_ARG, "null out pointer");
   *out = nullptr;
   int count = 0;
   cudaError_t err = cudaGetDeviceCount(&count);
   if (err != cudaSuccess || count == 0)
   {
      cudaGetLastError();
      return fail(OCB_ERR_NODEVICE, "no CUDA device (%s); this engine has no CPU path",
                  err == cudaSuccess ? "count = 0" : cudaGetErrorString(err));
   }
   if (device < 0 || device >= count) return fail(OCB_ERR_ARG, "device %d out of range (%d)", device, count);
   cudaDeviceProp prop;
   CU(cudaGetDeviceProperties(&prop, device));
   if (prop.major < 10)
      return fail(OCB_ERR_NODEVICE, "device %d is sm_%d%d; this build carries sm_100a code only", device,
                  prop.major, prop.minor);
   CU(cudaSetDevice(device));
   ocb_engine *e = new (std::nothrow) ocb_engine();
   if (!e) return fail(OCB_ERR_ALLOC, "out of host memory");
   e->device = device;
   e->smem_optin = (int) prop.sharedMemPerBlockOptin;
   e->smem_per_sm = (int) prop.sharedMemPerMultiprocessor;
   e->sm_count = prop.multiProcessorCount;
   if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess)
   {
      delete e;
      return fail(OCB_ERR_CUDA, "cudaStreamCreate failed");
   }
   e->stream = e->own_stream;
   cudaMemPoolProps pp;
   memset(&pp, 0, sizeof(pp));
   pp.allocType = cudaMemAllocationTypePinned;
   pp.handleTypes = cudaMemHandleTypeNone;
   pp.location.type = cudaMemLocationTypeDevice;
   pp.location.id = device;
   if (cudaMemPoolCreate(&e->pool, &pp) != cudaSuccess)
   {
      cudaStreamDestroy(e->own_stream);
      delete e;
      return fail(OCB_ERR_CUDA, "cudaMemPoolCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
   }
   {
      const char *env = getenv("OCB_JIT"); /* OCB_JIT=1: run-time specialisation on for every engine */
      e->jit = (env && env[0] && env[0] != '0') ? 1 : 0;
   }
   unsigned long long keep = ~0ull; /* never hand memory back to the driver between batches */
   cudaMemPoolSetAttribute(e->pool, cudaMemPoolAttrReleaseThreshold, &keep);
   *out = e;
   return OCB_OK;
}

struct ocb_batch;
static ocb_engine *batch_engine(const ocb_batch *b);

/* Handles of live batches, process wide: a batch outliving its engine (a caller that destroys the
 * engine first, a garbage collector that runs late) must not touch the freed engine.  Destroying an
 * engine destroys the batches it still owns; destroying a batch that is no longer registered is a
 * harmless no-op. */
static std::mutex g_batches_lock;
static std::set<ocb_batch *> g_batches;
static void batch_free(ocb_batch *b);

extern "C" int ocb_engine_destroy(ocb_engine *e)
{
   if (!e) return OCB_OK;
   cudaSetDevice(e->device);
   cudaStreamSynchronize(e->stream);
   {
      std::vector<ocb_batch *> mine;
      {
         std::lock_guard<std::mutex> lock(g_batches_lock);
         for (ocb_batch *b : g_batches)
            if (batch_engine(b) == e) mine.push_back(b);
         for (ocb_batch *b : mine) g_batches.erase(b);
      }
      for (ocb_batch *b : mine) batch_free(b);
   }
   for (auto &s : e->sdfs)
      if (s.used && s.owned) pool_free(e, s.d_data);
   if (e->scratch) cudaFree(e->scratch);
   cudaStreamSynchronize(e->stream);
   if (e->pool) cudaMemPoolDestroy(e->pool);
   if (e->own_stream) cudaStreamDestroy(e->own_stream);
   delete e;
   return OCB_OK;
}

extern "C" int ocb_engine_set_stream(ocb_engine *e, void *cuda_stream)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   cudaStreamSynchronize(e->stream);
   e->stream = cuda_stream ? (cudaStream_t) cuda_stream : e->own_stream;
   return OCB_OK;
}

extern "C" int ocb_engine_sync(ocb_engine *e)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->stream));
   return OCB_OK;
}

extern "C" long ocb_engine_launch_count(const ocb_engine *e) { return e ? e->launches : 0; }

extern "C" int ocb_engine_trim(ocb_engine *e)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->stream));
   CU(cudaMemPoolTrimTo(e->pool, 0));
   return OCB_OK;
}

extern "C" int ocb_engine_enable_jit(ocb_engine *e, int on)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   e->jit = on ? 1 : 0;
   return OCB_OK;
}

extern "C" int ocb_engine_force_general_sdf(ocb_engine *e, int on)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   e->force_general_sdf = on ? 1 : 0;
   return OCB_OK;
}

/* --------------------------------------------------------------------- SDFs */
static int sdf_slot_new(ocb_engine *e)
{
   for (size_t i = 0; i < e->sdfs.size(); i++)
      if (!e->sdfs[i].used) return (int) i;
   e->sdfs.push_back(SdfSlot());
   return (int) e->sdfs.size() - 1;
}

static int check_grid(const int sizes[3], const double lengths[3])
{
   for (int i = 0; i < 3; i++)
   {
      if (sizes[i] < 2) return fail(OCB_ERR_ARG, "grid size %d on axis %d (need >= 2; grid.c:352 indexes a neighbour)", sizes[i], i);
      if (!(lengths[i] > 0.0)) return fail(OCB_ERR_ARG, "grid length must be positive");
      if (sizes[i] > 46340) return fail(OCB_ERR_ARG, "grid axis longer than 46340 cells (q*q overflows int, grid.c:298)");
   }
   return OCB_OK;
}

extern "C" int ocb_sdf_upload(ocb_engine *e, const ocb_sdf *sdf, int *id)
{
   if (!e || !sdf || !id || !sdf->data) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sdf->sizes, sdf->lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t cells = (size_t) sdf->sizes[0] * sdf->sizes[1] * sdf->sizes[2];
   double *d = nullptr;
   if (pool_alloc(e, (void **) &d, cells * sizeof(double)) != cudaSuccess)
   {
      cudaGetLastError();
      return fail(OCB_ERR_ALLOC, "device allocation of %zu SDF bytes failed", cells * sizeof(double));
   }
   cudaError_t err = cudaMemcpyAsync(d, sdf->data, cells * sizeof(double), cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
   if (err != cudaSuccess)
   {
      pool_free(e, d);
      return fail(OCB_ERR_CUDA, "SDF upload: %s", cudaGetErrorString(err));
   }
   const int slot = sdf_slot_new(e);
   SdfSlot &s = e->sdfs[slot];
   s.used = true;
   s.owned = true;
   s.d_data = d;
   for (int i = 0; i < 3; i++) { s.sizes[i] = sdf->sizes[i]; s.lengths[i] = sdf->lengths[i]; }
   memcpy(s.pose, sdf->pose_world_gsdf, sizeof(s.pose));
   *id = slot;
   return OCB_OK;
}

extern "C" int ocb_sdf_adopt_device(ocb_engine *e, const int sizes[3], const double lengths[3],
                                    const double pose_world_gsdf[7], const double *d_data, int *id)
{
   if (!e || !d_data || !id) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   const int slot = sdf_slot_new(e);
   SdfSlot &s = e->sdfs[slot];
   s.used = true;
   s.owned = false;
   s.d_data = const_cast<double *>(d_data);
   for (int i = 0; i < 3; i++) { s.sizes[i] = sizes[i]; s.lengths[i] = lengths[i]; }
   memcpy(s.pose, pose_world_gsdf, sizeof(s.pose));
   *id = slot;
   return OCB_OK;
}

extern "C" int ocb_sdf_remove(ocb_engine *e, int id)
{
   if (!e || id < 0 || id >= (int) e->sdfs.size() || !e->sdfs[id].used) return fail(OCB_ERR_ARG, "bad sdf id %d", id);
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->stream));
   if (e->sdfs[id].owned) pool_free(e, e->sdfs[id].d_data);
   e->sdfs[id] = SdfSlot();
   return OCB_OK;
}

extern "C" int ocb_dt_sqeuc_device(ocb_engine *e, const double *d_func, const int sizes[3],
                                   const double lengths[3], double *d_out)
{
   if (!e || !d_func || !d_out) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   rc = engine_scratch(e, ocb_dt_scratch_bytes(sizes));
   if (rc) return rc;
   CU(ocb_launch_dt_sqeuc(d_func, d_out, sizes, lengths, e->scratch, e->scratch_bytes, e->stream, &e->launches));
   return OCB_OK;
}

extern "C" int ocb_sdf_build_device(ocb_engine *e, const double *d_obs, const int sizes[3],
                                    const double lengths[3], double *d_sdf)
{
   if (!e || !d_obs || !d_sdf) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   if (ocb_sdf_fast_eligible(sizes, lengths) && !e->force_general_sdf)
   {
      rc = engine_scratch(e, ocb_sdf_fast_scratch_bytes(sizes));
      if (rc) return rc;
      int used_fast = 0;
      CU(ocb_launch_bin_sdf_fast(d_obs, d_sdf, sizes, lengths, e->scratch, e->scratch_bytes, e->stream,
                                 &e->launches, &used_fast));
      if (used_fast) return OCB_OK;
   }
   rc = engine_scratch(e, ocb_sdf_scratch_bytes(sizes, lengths));
   if (rc) return rc;
   CU(ocb_launch_bin_sdf(d_obs, d_sdf, sizes, lengths, e->scratch, e->scratch_bytes, e->stream, &e->launches));
   return OCB_OK;
}

extern "C" int ocb_sdf_build_host(ocb_engine *e, const double *obs, const int sizes[3],
                                  const double lengths[3], double *sdf)
{
   if (!e || !obs || !sdf) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d_in = nullptr, *d_out = nullptr;
   if (pool_alloc(e, (void **) &d_in, bytes) != cudaSuccess || pool_alloc(e, (void **) &d_out, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      pool_free(e, d_in);
      return fail(OCB_ERR_ALLOC, "device allocation of 2 x %zu bytes failed", bytes);
   }
   cudaError_t err = cudaMemcpyAsync(d_in, obs, bytes, cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess)
   {
      rc = ocb_sdf_build_device(e, d_in, sizes, lengths, d_out);
      if (rc == OCB_OK)
      {
         err = cudaMemcpyAsync(sdf, d_out, bytes, cudaMemcpyDeviceToHost, e->stream);
         if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
      }
   }
   pool_free(e, d_in);
   pool_free(e, d_out);
   if (rc) return rc;
   if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "sdf_build_host: %s", cudaGetErrorString(err));
   return OCB_OK;
}

extern "C" int ocb_dt_sqeuc_host(ocb_engine *e, const double *func, const int sizes[3],
                                 const double lengths[3], double *out)
{
   if (!e || !func || !out) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d_in = nullptr, *d_out = nullptr;
   if (pool_alloc(e, (void **) &d_in, bytes) != cudaSuccess || pool_alloc(e, (void **) &d_out, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      pool_free(e, d_in);
      return fail(OCB_ERR_ALLOC, "device allocation of 2 x %zu bytes failed", bytes);
   }
   cudaError_t err = cudaMemcpyAsync(d_in, func, bytes, cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess)
   {
      rc = ocb_dt_sqeuc_device(e, d_in, sizes, lengths, d_out);
      if (rc == OCB_OK)
      {
         err = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, e->stream);
         if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
      }
   }
   pool_free(e, d_in);
   pool_free(e, d_out);
   if (rc) return rc;
   if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "dt_sqeuc_host: %s", cudaGetErrorString(err));
   return OCB_OK;
}

extern "C" int ocb_occupancy_device(ocb_engine *e, const ocb_prim *prims, int n_prims, const int sizes[3],
                                    const double lengths[3], double cube_extent, double *d_grid)
{
   if (!e || !d_grid || (n_prims > 0 && !prims)) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   rc = engine_scratch(e, std::max<size_t>(1, n_prims) * sizeof(ocb_prim));
   if (rc) return rc;
   if (n_prims > 0)
      CU(cudaMemcpyAsync(e->scratch, prims, n_prims * sizeof(ocb_prim), cudaMemcpyHostToDevice, e->stream));
   /* blocks per primitive from the mean size of the voxel ranges (host estimate, bounds only) */
   int slices = 1;
   if (n_prims > 0)
   {
      double vox = 0.0;
      for (int i = 0; i < n_prims; i++)
      {
         double v = 1.0;
         for (int k = 0; k < 3; k++)
         {
            double w;
            if (prims[i].type == OCB_PRIM_TRIANGLE)
            {
               const double a = prims[i].pose[k], b = prims[i].pose[3 + k];
               const double c = (k == 0) ? prims[i].pose[6] : prims[i].extents[k - 1];
               w = std::max(a, std::max(b, c)) - std::min(a, std::min(b, c));
            }
            else if (prims[i].type == OCB_PRIM_SPHERE) w = 2.0 * prims[i].extents[0];
            else w = 2.0 * sqrt(prims[i].extents[0] * prims[i].extents[0] + prims[i].extents[1] * prims[i].extents[1] +
                                prims[i].extents[2] * prims[i].extents[2]);
            v *= w / (lengths[k] / sizes[k]) + 3.0;
         }
         vox += v;
      }
      slices = (int) std::min(64.0, std::max(1.0, vox / n_prims / 4096.0));
   }
   CU(ocb_launch_occupancy(e->scratch, n_prims, sizes, lengths, cube_extent, d_grid, slices, e->stream));
   e->launches += 3;
   /* the primitives live in the shared scratch: finish before anyone reuses it */
   CU(cudaStreamSynchronize(e->stream));
   return OCB_OK;
}

extern "C" int ocb_flood_relabel_device(ocb_engine *e, double *d_grid, const int sizes[3], size_t index_start)
{
   if (!e || !d_grid) return fail(OCB_ERR_ARG, "null argument");
   const size_t cells = (size_t) sizes[0] * sizes[1] * sizes[2];
   if (index_start >= cells) return fail(OCB_ERR_ARG, "flood start outside the grid");
   CU(cudaSetDevice(e->device));
   int rc = engine_scratch(e, ocb_flood_scratch_bytes(sizes));
   if (rc) return rc;
   CU(ocb_launch_flood_relabel(d_grid, sizes, index_start, e->scratch, e->scratch_bytes, e->stream, &e->launches));
   return OCB_OK;
}

extern "C" int ocb_flood_relabel_host(ocb_engine *e, double *grid, const int sizes[3], size_t index_start)
{
   if (!e || !grid) return fail(OCB_ERR_ARG, "null argument");
   if (sizes[0] < 1 || sizes[1] < 1 || sizes[2] < 1) return fail(OCB_ERR_ARG, "bad grid sizes");
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d = nullptr;
   if (pool_alloc(e, (void **) &d, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      return fail(OCB_ERR_ALLOC, "device allocation of %zu bytes failed", bytes);
   }
   int rc = OCB_OK;
   cudaError_t err = cudaMemcpyAsync(d, grid, bytes, cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess)
   {
      rc = ocb_flood_relabel_device(e, d, sizes, index_start);
      if (rc == OCB_OK)
      {
         err = cudaMemcpyAsync(grid, d, bytes, cudaMemcpyDeviceToHost, e->stream);
         if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
      }
   }
   pool_free(e, d);
   if (rc) return rc;
   if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "flood_relabel_host: %s", cudaGetErrorString(err));
   return OCB_OK;
}

extern "C" int ocb_computedistancefield_host(ocb_engine *e, const ocb_prim *prims, int n_prims,
                                             const int sizes[3], const double lengths[3],
                                             double cube_extent, double *obs_out, double *sdf_out)
{
   if (!e) return fail(OCB_ERR_ARG, "null engine");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d_obs = nullptr, *d_sdf = nullptr;
   if (pool_alloc(e, (void **) &d_obs, bytes) != cudaSuccess || pool_alloc(e, (void **) &d_sdf, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      pool_free(e, d_obs);
      return fail(OCB_ERR_ALLOC, "device allocation of 2 x %zu bytes failed", bytes);
   }
   rc = ocb_occupancy_device(e, prims, n_prims, sizes, lengths, cube_extent, d_obs);
   if (rc == OCB_OK) rc = ocb_flood_relabel_device(e, d_obs, sizes, 0);
   cudaError_t err = cudaSuccess;
   if (rc == OCB_OK && obs_out) err = cudaMemcpyAsync(obs_out, d_obs, bytes, cudaMemcpyDeviceToHost, e->stream);
   if (rc == OCB_OK && err == cudaSuccess && sdf_out)
   {
      rc = ocb_sdf_build_device(e, d_obs, sizes, lengths, d_sdf);
      if (rc == OCB_OK) err = cudaMemcpyAsync(sdf_out, d_sdf, bytes, cudaMemcpyDeviceToHost, e->stream);
   }
   if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
   pool_free(e, d_obs);
   pool_free(e, d_sdf);
   if (rc) return rc;
   if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "computedistancefield_host: %s", cudaGetErrorString(err));
   return OCB_OK;
}

/* ---- fields that never leave the device ------------------------------------------------- */
static int sdf_slot_own(ocb_engine *e, double *d, const int sizes[3], const double lengths[3],
                        const double pose[7], bool owned, int *id)
{
   const int slot = sdf_slot_new(e);
   SdfSlot &s = e->sdfs[slot];
   s.used = true;
   s.owned = owned;
   s.d_data = d;
   for (int i = 0; i < 3; i++) { s.sizes[i] = sizes[i]; s.lengths[i] = lengths[i]; }
   memcpy(s.pose, pose, sizeof(s.pose));
   *id = slot;
   return OCB_OK;
}

extern "C" int ocb_computedistancefield_resident(ocb_engine *e, const ocb_prim *prims, int n_prims,
                                                 const int sizes[3], const double lengths[3], double cube_extent,
                                                 const double pose_world_gsdf[7], int *id)
{
   if (!e || !id || !pose_world_gsdf) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d_obs = nullptr, *d_sdf = nullptr;
   if (pool_alloc(e, (void **) &d_obs, bytes) != cudaSuccess || pool_alloc(e, (void **) &d_sdf, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      pool_free(e, d_obs);
      return fail(OCB_ERR_ALLOC, "device allocation of 2 x %zu bytes failed", bytes);
   }
   rc = ocb_occupancy_device(e, prims, n_prims, sizes, lengths, cube_extent, d_obs);
   if (rc == OCB_OK) rc = ocb_flood_relabel_device(e, d_obs, sizes, 0);
   if (rc == OCB_OK) rc = ocb_sdf_build_device(e, d_obs, sizes, lengths, d_sdf);
   pool_free(e, d_obs);
   if (rc)
   {
      pool_free(e, d_sdf);
      return rc;
   }
   return sdf_slot_own(e, d_sdf, sizes, lengths, pose_world_gsdf, true, id);
}

extern "C" int ocb_sdf_build_resident(ocb_engine *e, const double *obs, const int sizes[3], const double lengths[3],
                                      const double pose_world_gsdf[7], int *id)
{
   if (!e || !obs || !id || !pose_world_gsdf) return fail(OCB_ERR_ARG, "null argument");
   int rc = check_grid(sizes, lengths);
   if (rc) return rc;
   CU(cudaSetDevice(e->device));
   const size_t bytes = (size_t) sizes[0] * sizes[1] * sizes[2] * sizeof(double);
   double *d_obs = nullptr, *d_sdf = nullptr;
   if (pool_alloc(e, (void **) &d_obs, bytes) != cudaSuccess || pool_alloc(e, (void **) &d_sdf, bytes) != cudaSuccess)
   {
      cudaGetLastError();
      pool_free(e, d_obs);
      return fail(OCB_ERR_ALLOC, "device allocation of 2 x %zu bytes failed", bytes);
   }
   cudaError_t err = cudaMemcpyAsync(d_obs, obs, bytes, cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess) rc = ocb_sdf_build_device(e, d_obs, sizes, lengths, d_sdf);
   if (err == cudaSuccess && rc == OCB_OK) err = cudaStreamSynchronize(e->stream); /* obs is the caller's */
   pool_free(e, d_obs);
   if (rc || err != cudaSuccess)
   {
      pool_free(e, d_sdf);
      return rc ? rc : fail(OCB_ERR_CUDA, "sdf_build_resident: %s", cudaGetErrorString(err));
   }
   return sdf_slot_own(e, d_sdf, sizes, lengths, pose_world_gsdf, true, id);
}

extern "C" int ocb_sdf_download(ocb_engine *e, int id, double *out)
{
   if (!e || !out || id < 0 || id >= (int) e->sdfs.size() || !e->sdfs[id].used) return fail(OCB_ERR_ARG, "bad sdf id %d", id);
   CU(cudaSetDevice(e->device));
   const SdfSlot &s = e->sdfs[id];
   const size_t bytes = (size_t) s.sizes[0] * s.sizes[1] * s.sizes[2] * sizeof(double);
   CU(cudaMemcpyAsync(out, s.d_data, bytes, cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   return OCB_OK;
}

extern "C" int ocb_sdf_alias(ocb_engine *e, int id, const double pose_world_gsdf[7], int *alias_id)
{
   if (!e || !alias_id || !pose_world_gsdf || id < 0 || id >= (int) e->sdfs.size() || !e->sdfs[id].used)
      return fail(OCB_ERR_ARG, "bad sdf id %d", id);
   const SdfSlot src = e->sdfs[id]; /* copy: sdf_slot_new may grow the table */
   return sdf_slot_own(e, src.d_data, src.sizes, src.lengths, pose_world_gsdf, false, alias_id);
}

/* -------------------------------------------------------------------- batch */
/* device-side descriptor of a resident field (what the kernels stage in shared memory) */
static void fill_sdf_dev(const SdfSlot &s, OcbSdfDev &d)
{
   memset(&d, 0, sizeof(d));
   d.data = s.d_data;
   for (int k = 0; k < 3; k++)
   {
      d.size[k] = s.sizes[k];
      d.length[k] = s.lengths[k];
      d.scale[k] = s.sizes[k] / s.lengths[k];
      d.cell[k] = s.lengths[k] / s.sizes[k];
      /* sdf_sample's exact-path thresholds (chomp_device.cuh): relative 1e-10 around the decisions */
      d.edge_hi[k] = s.sizes[k] * (1.0 + 1e-10);
      d.near[k] = s.sizes[k] * 1e-10;
      d.near_hi[k] = 0.5 - d.near[k];
   }
   /* pose_gsdf_world = inverse of the snapshot pose (cd_kin_pose_invert, mod.cpp:2368) */
   const double *p = s.pose;
   const double qi[4] = {-p[3], -p[4], -p[5], p[6]};
   double Ri[9];
   quat_to_R(qi, Ri);
   for (int r = 0; r < 3; r++) d.tgw[r] = -(Ri[3 * r] * p[0] + Ri[3 * r + 1] * p[1] + Ri[3 * r + 2] * p[2]);
   memcpy(d.Rgw, Ri, sizeof(Ri));
   quat_to_R(p + 3, d.Rwg);
}

/* cd_grid_double_interp + cd_grid_double_grad (grid.c:331-454) of a resident field at k points given
 * in the GRID frame -- the device function the CHOMP kernels sample with, exposed for parity tests */
extern "C" int ocb_sdf_sample_host(ocb_engine *e, int id, const double *points, int k, double *values,
                                   double *grads, int *errs)
{
   if (!e || !points || !values || !grads || !errs || k < 0) return fail(OCB_ERR_ARG, "bad argument");
   if (id < 0 || id >= (int) e->sdfs.size() || !e->sdfs[id].used) return fail(OCB_ERR_ARG, "bad sdf id %d", id);
   if (k == 0) return OCB_OK;
   CU(cudaSetDevice(e->device));
   OcbSdfDev d;
   fill_sdf_dev(e->sdfs[id], d);
   double *dp = nullptr, *dv = nullptr, *dg = nullptr;
   int *de = nullptr;
   cudaError_t err = pool_alloc(e, (void **) &dp, (size_t) k * 3 * sizeof(double));
   if (err == cudaSuccess) err = pool_alloc(e, (void **) &dv, (size_t) k * sizeof(double));
   if (err == cudaSuccess) err = pool_alloc(e, (void **) &dg, (size_t) k * 3 * sizeof(double));
   if (err == cudaSuccess) err = pool_alloc(e, (void **) &de, (size_t) k * sizeof(int));
   if (err == cudaSuccess) err = cudaMemcpyAsync(dp, points, (size_t) k * 3 * sizeof(double), cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess) err = ocb_launch_sdf_sample(&d, dp, k, dv, dg, de, e->stream);
   if (err == cudaSuccess) e->launches++;
   if (err == cudaSuccess) err = cudaMemcpyAsync(values, dv, (size_t) k * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
   if (err == cudaSuccess) err = cudaMemcpyAsync(grads, dg, (size_t) k * 3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
   if (err == cudaSuccess) err = cudaMemcpyAsync(errs, de, (size_t) k * sizeof(int), cudaMemcpyDeviceToHost, e->stream);
   if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
   pool_free(e, dp); pool_free(e, dv); pool_free(e, dg); pool_free(e, de);
   if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "sdf_sample: %s", cudaGetErrorString(err));
   return OCB_OK;
}

struct ocb_batch
{
   ocb_engine *e = nullptr;
   OcbChompArgs args;
   int threads = 128;
   size_t smem = 0;
   size_t tile_smem = 0, run_smem = 0; /* tiled path (args.tiled) */
   void *jit_kernel = nullptr;         /* run-time specialised persistent kernel, when the engine asks for it */
   int trace_cap = 0; /* iterations the trace buffer can hold */
   int last_n_iter = 0;
   std::vector<void *> owned; /* device allocations to free */
   double *d_start = nullptr, *d_goal = nullptr;
   std::vector<uint32_t> mt_host; /* seeded generator states, [R][625] (hmc only) */
};

namespace
{
/* sparse row of a finite-difference operator */
typedef std::vector<std::pair<int, double>> SRow;

void srow_axpy(SRow &dst, double alpha, const SRow &src)
{
   for (const auto &e : src)
   {
      bool found = false;
      for (auto &d : dst)
         if (d.first == e.first) { d.second += alpha * e.second; found = true; break; }
      if (!found) dst.push_back(std::make_pair(e.first, alpha * e.second));
   }
}

/* Smoothness metric of cd_chomp_add_KEs (chomp.c:239-340) for the module's
 * set-up (inits[d] / finals[d] all present, only d = 0 non-zero: mod.cpp:2578-2580,
 * chomp.c:131-141), kept in banded / coefficient form:
 *   A    = sum_d w_d/N_d K_d^T K_d                    -> band [m][2D+1]
 *   B    = sum_d w_d/N_d K_d^T E_d = bi (x) q_start + bf (x) q_goal
 *   trC  = 1/2 (ss |q_s|^2 + 2 sg q_s.q_g + gg |q_g|^2)                       */
struct Metric
{
   int m, bw;
   std::vector<double> Aband, Lband, dinv, bi, bf;
   double ss, sg, gg;
};

int build_metric(int m, int D, double dt, Metric &M, bool free_start = false)
{
   M.m = m;
   M.bw = D;
   const int bw = D;
   M.Aband.assign((size_t) m * (2 * bw + 1), 0.0);
   M.bi.assign(m, 0.0);
   M.bf.assign(m, 0.0);
   M.ss = M.sg = M.gg = 0.0;
   std::vector<double> wds(D);
   for (int d = 0; d < D; d++) wds[d] = (d < D - 1) ? 0.0 : 1.0; /* chomp.c:127-128 */

   std::vector<SRow> Kprev; /* identity on the m moving points (N_{-1} = m) */
   std::vector<double> ciprev, cfprev;
   int prev = m;
   for (int d = 0; d < D; d++)
   {
      /* start_tsr leaves inits[0] unset (mod.cpp:2571-2576): level 0 then has no row for the start
       * boundary (chomp.c:262-264, 278-283); the higher levels keep theirs (zero vectors, chomp.c:131-141) */
      const int has_i = (d == 0 && free_start) ? 0 : 1;
      const int cur = prev - 1 + has_i + 1; /* interior rows + init row + final row */
      std::vector<SRow> K(cur);
      std::vector<double> ci(cur, 0.0), cf(cur, 0.0);
      /* rows of the differencing matrix: (col, value) pairs over the previous level */
      for (int r = 0; r < cur; r++)
      {
         SRow diff;
         if (has_i && r == 0) diff.push_back(std::make_pair(0, 1.0 / dt));
         else if (r == cur - 1) diff.push_back(std::make_pair(prev - 1, -1.0 / dt));
         else
         {
            diff.push_back(std::make_pair(r - has_i, -1.0 / dt));
            diff.push_back(std::make_pair(r - has_i + 1, 1.0 / dt));
         }
         for (const auto &e : diff)
         {
            if (d == 0)
               K[r].push_back(std::make_pair(e.first, e.second));
            else
            {
               srow_axpy(K[r], e.second, Kprev[e.first]);
               ci[r] += e.second * ciprev[e.first];
               cf[r] += e.second * cfprev[e.first];
            }
         }
      }
      if (d == 0)
      {
         if (has_i) ci[0] += -1.0 / dt; /* E row 0   = -inits[0]/dt  (chomp.c:281) */
         cf[cur - 1] += 1.0 / dt; /* E row N-1 = +finals[0]/dt (chomp.c:295) */
      }
      const double w = wds[d] / cur;
      /* S = K^T K accumulated over rows, then A += w S (dgemm alpha, chomp.c:318-320) */
      std::vector<double> S((size_t) m * (2 * bw + 1), 0.0), sbi(m, 0.0), sbf(m, 0.0);
      double sss = 0, ssg = 0, sgg = 0;
      for (int r = 0; r < cur; r++)
      {
         for (const auto &e1 : K[r])
         {
            for (const auto &e2 : K[r])
            {
               const int off = e2.first - e1.first;
               if (off < -bw || off > bw) return -1;
               S[(size_t) e1.first * (2 * bw + 1) + off + bw] += e1.second * e2.second;
            }
            sbi[e1.first] += e1.second * ci[r];
            sbf[e1.first] += e1.second * cf[r];
         }
         sss += ci[r] * ci[r];
         ssg += ci[r] * cf[r];
         sgg += cf[r] * cf[r];
      }
      for (size_t k = 0; k < S.size(); k++) M.Aband[k] += w * S[k];
      for (int i = 0; i < m; i++) { M.bi[i] += w * sbi[i]; M.bf[i] += w * sbf[i]; }
      M.ss += w * sss;
      M.sg += w * ssg;
      M.gg += w * sgg;
      Kprev.swap(K);
      ciprev.swap(ci);
      cfprev.swap(cf);
      prev = cur;
   }
   /* banded LDL^T of A (stands in for dgetrf/dgetri, chomp.c:393-403) */
   M.Lband.assign((size_t) m * bw, 0.0);
   M.dinv.assign(m, 0.0);
   std::vector<double> dd(m, 0.0);
   auto A = [&](int i, int k) { return M.Aband[(size_t) i * (2 * bw + 1) + (k - i) + bw]; };
   auto L = [&](int i, int k) -> double & { return M.Lband[(size_t) i * bw + (i - k - 1)]; };
   for (int i = 0; i < m; i++)
   {
      for (int j = std::max(0, i - bw); j < i; j++)
      {
         double acc = A(i, j);
         for (int p = std::max(0, i - bw); p < j; p++)
            if (j - p <= bw) acc -= L(i, p) * dd[p] * L(j, p);
         L(i, j) = acc / dd[j];
      }
      double acc = A(i, i);
      for (int p = std::max(0, i - bw); p < i; p++) acc -= L(i, p) * L(i, p) * dd[p];
      if (!(acc > 0.0)) return -2;
      dd[i] = acc;
      M.dinv[i] = 1.0 / acc;
   }
   return 0;
}

struct CompiledRobot
{
   std::vector<OcbJointDev> joints;
   std::vector<OcbSphereDev> spheres;
   std::vector<double> cut2, radius; /* [nsa][NS], [NS] */
   int n_groups = 0;
   std::vector<int> desc;
   std::vector<int> ganc; /* [n_groups + 1] offsets, then joints above each group (see OcbChompArgs) */
   std::vector<double> gbound; /* [nj][4] bounding sphere of each joint frame's spheres (see OcbChompArgs) */
   std::vector<double> inactive_pos;
   std::vector<double> inactive_radius;
   std::vector<int> inactive_link;
   int n_slots = 0;
   /* where every robot link sits: its joint frame (index in `joints`, -1: not moved by an active dof) and
    * the link frame in that joint frame (or in the world); the joint above each joint (-1: none) */
   std::vector<int> link_joint, joint_parent;
   std::vector<Xf> link_rel;
};

/* Fold fixed / frozen links into joint frames with the moving axis on local z.
 * Replaces the OpenRAVE-side bookkeeping of mod::create (mod.cpp:2148-2300) and
 * prepares what sphere_cost_pre asks OpenRAVE for on every waypoint. */
int compile_robot(const ocb_robot *rb, double eps_self, bool floating, CompiledRobot &C)
{
   const int nl = rb->n_links;
   if (nl < 1 || rb->n_dof < 1) return fail(OCB_ERR_ARG, "robot needs links and active dofs");
   std::vector<int> anchor(nl, -1);  /* nearest moving ancestor-or-self (index in `raw` joints) */
   std::vector<Xf> rel(nl);          /* link frame expressed in its anchor's joint frame (or world) */
   struct RawJoint { int parent; Xf X; int type, dof; double c0, c1; int link; };
   std::vector<RawJoint> raw;
   const int dof0 = floating ? 7 : 0; /* trajectory column of active dof 0 */
   if (floating)
   {
      /* floating base (mod.cpp:991-1021): the base frame is a joint frame of its own whose
       * transform the kernel takes from the waypoint's pose entries; it has no dof (c0 = 0),
       * carries the spheres of every link that no active dof moves (all spheres are active,
       * mod.cpp:2274) and is the parent of the root-level joints */
      const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
      RawJoint B;
      B.parent = -1;
      B.X = xf_from_pose(ident);
      B.type = OCB_JOINT_PRISMATIC;
      B.dof = 0;
      B.c0 = 0.0;
      B.c1 = 0.0;
      B.link = 0;
      raw.push_back(B);
      anchor[0] = 0;
      rel[0] = xf_from_pose(ident);
   }
   else
      rel[0] = xf_from_pose(rb->base_pose);
   for (int i = 1; i < nl; i++)
   {
      const int p = rb->parent[i];
      if (p < 0 || p >= i) return fail(OCB_ERR_ARG, "link %d: parent must precede it", i);
      const Xf Xp = xf_mul(rel[p], xf_from_pose(rb->pose_parent + 7 * i));
      const int type = rb->joint_type[i];
      const bool moving = (type != OCB_JOINT_FIXED) && rb->dof_index[i] >= 0;
      if (moving)
      {
         if (rb->dof_index[i] >= rb->n_dof) return fail(OCB_ERR_ARG, "link %d: dof index out of range", i);
         const Xf Q = xf_align_z(rb->axis + 3 * i);
         RawJoint J;
         J.parent = anchor[p];
         J.X = xf_mul(Xp, Q);
         J.type = type;
         J.dof = dof0 + rb->dof_index[i];
         J.c0 = rb->dof_coeff[2 * i];
         J.c1 = rb->dof_coeff[2 * i + 1];
         J.link = i;
         raw.push_back(J);
         anchor[i] = (int) raw.size() - 1;
         rel[i] = xf_transpose_rot(Q);
      }
      else
      {
         anchor[i] = anchor[p];
         rel[i] = xf_mul(Xp, xf_motion(type, rb->axis + 3 * i, rb->dof_coeff[2 * i + 1]));
      }
   }
   const int nj = (int) raw.size();
   if (nj == 0) return fail(OCB_ERR_ARG, "no moving joints");
   if (nj > OCB_MAX_JOINTS) return fail(OCB_ERR_ARG, "%d moving joints (max %d)", nj, OCB_MAX_JOINTS);

   /* depth-first order so a child usually follows its parent directly */
   std::vector<int> order, newidx(nj, -1);
   {
      std::vector<std::vector<int>> kids(nj);
      std::vector<int> roots;
      for (int j = 0; j < nj; j++)
         if (raw[j].parent < 0) roots.push_back(j);
         else kids[raw[j].parent].push_back(j);
      std::vector<int> stack(roots.rbegin(), roots.rend());
      while (!stack.empty())
      {
         const int j = stack.back();
         stack.pop_back();
         newidx[j] = (int) order.size();
         order.push_back(j);
         for (auto it = kids[j].rbegin(); it != kids[j].rend(); ++it) stack.push_back(*it);
      }
   }
   C.joints.assign(nj, OcbJointDev());
   std::vector<int> slot_of(nj, -1);
   C.n_slots = 0;
   for (int k = 0; k < nj; k++)
   {
      const RawJoint &J = raw[order[k]];
      OcbJointDev &D = C.joints[k];
      memcpy(D.XR, J.X.R, sizeof(D.XR));
      memcpy(D.Xt, J.X.t, sizeof(D.Xt));
      D.c0 = J.c0;
      D.c1 = J.c1;
      D.type = J.type;
      D.dof = J.dof;
      D.save = -1;
      const int par = (J.parent < 0) ? -1 : newidx[J.parent];
      if (par < 0) D.load = OCB_LOAD_BASE;
      else if (par == k - 1) D.load = OCB_LOAD_PREV;
      else
      {
         if (slot_of[par] < 0) { slot_of[par] = C.n_slots++; C.joints[par].save = slot_of[par]; }
         D.load = slot_of[par];
      }
   }
   C.link_joint.assign(nl, -1);
   C.link_rel = rel;
   for (int i = 0; i < nl; i++)
      if (anchor[i] >= 0) C.link_joint[i] = newidx[anchor[i]];
   C.joint_parent.assign(nj, -1);
   for (int k = 0; k < nj; k++)
      if (raw[order[k]].parent >= 0) C.joint_parent[k] = newidx[raw[order[k]].parent];
   /* spheres: active ones grouped by joint (stable), inactive ones frozen in the world */
   struct Tmp { int joint; OcbSphereDev s; };
   std::vector<Tmp> act;
   for (int s = 0; s < rb->n_spheres; s++)
   {
      const int link = rb->sphere_link[s];
      if (link < 0 || link >= nl) return fail(OCB_ERR_ARG, "sphere %d: bad link", s);
      double p[3];
      xf_apply(rel[link], rb->sphere_pos + 3 * s, p);
      if (anchor[link] >= 0)
      {
         Tmp t;
         t.joint = newidx[anchor[link]];
         t.s.pos[0] = p[0]; t.s.pos[1] = p[1]; t.s.pos[2] = p[2];
         t.s.radius = rb->sphere_radius[s];
         t.s.link = link;
         t.s.group = 0;
         act.push_back(t);
      }
      else
      {
         C.inactive_pos.push_back(p[0]);
         C.inactive_pos.push_back(p[1]);
         C.inactive_pos.push_back(p[2]);
         C.inactive_radius.push_back(rb->sphere_radius[s]);
         C.inactive_link.push_back(link);
      }
   }
   if (act.empty()) return fail(OCB_ERR_ARG, "robot active dofs must have at least one sphere!");
   std::stable_sort(act.begin(), act.end(), [](const Tmp &a, const Tmp &b) { return a.joint < b.joint; });
   for (int k = 0; k < nj; k++) C.joints[k].sph_begin = C.joints[k].sph_end = 0;
   for (size_t i = 0; i < act.size(); i++) C.spheres.push_back(act[i].s);
   for (int k = 0, i = 0; k < nj; k++)
   {
      C.joints[k].sph_begin = i;
      while (i < (int) act.size() && act[i].joint == k) i++;
      C.joints[k].sph_end = i;
   }
   const int nsa = (int) C.spheres.size(), nsi = (int) C.inactive_radius.size(), NS = nsa + nsi;
   /* compact group index per sphere-carrying joint frame */
   C.n_groups = 0;
   for (int k = 0; k < nj; k++)
      if (C.joints[k].sph_end > C.joints[k].sph_begin)
      {
         for (int s = C.joints[k].sph_begin; s < C.joints[k].sph_end; s++) C.spheres[s].group = C.n_groups;
         C.n_groups++;
      }
   C.radius.resize(NS);
   for (int o = 0; o < NS; o++) C.radius[o] = (o < nsa) ? C.spheres[o].radius : C.inactive_radius[o - nsa];
   const int NAp = nsa + 3; /* the 4-wide range test starts at s+1: up to 3 entries past nsa */
   C.cut2.assign((size_t) nsa * (NAp + nsi), -1.0);
   for (int s = 0; s < nsa; s++)
      for (int o = 0; o < NS; o++)
      {
         const int link2 = (o < nsa) ? C.spheres[o].link : C.inactive_link[o - nsa];
         if (o == s || link2 == C.spheres[s].link) continue; /* mod.cpp:1256 */
         const double cut = C.spheres[s].radius + C.radius[o] + eps_self; /* mod.cpp:1268 */
         C.cut2[(size_t) s * (NAp + nsi) + (o < nsa ? o : NAp + (o - nsa))] = cut * cut;
      }
   /* for every joint: the sphere groups carried by its subtree (its J^T sees their wrenches) */
   for (int k = 0; k < nj; k++)
   {
      C.joints[k].desc_begin = (int) C.desc.size();
      for (int k2 = 0; k2 < nj; k2++)
      {
         if (C.joints[k2].sph_end == C.joints[k2].sph_begin) continue;
         bool below = false;
         for (int j = order[k2]; j >= 0; j = raw[j].parent)
            if (j == order[k]) { below = true; break; }
         if (below) C.desc.push_back(C.spheres[C.joints[k2].sph_begin].group);
      }
      C.joints[k].desc_end = (int) C.desc.size();
   }
   /* bounding sphere of the spheres each joint frame carries (centre of their box, slightly inflated) */
   C.gbound.assign((size_t) nj * 4, -1.0);
   for (int k = 0; k < nj; k++)
   {
      const int sb = C.joints[k].sph_begin, se = C.joints[k].sph_end;
      if (se <= sb) continue;
      double lo[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, hi[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
      for (int s = sb; s < se; s++)
         for (int c = 0; c < 3; c++)
         {
            lo[c] = std::min(lo[c], C.spheres[s].pos[c]);
            hi[c] = std::max(hi[c], C.spheres[s].pos[c]);
         }
      double bc[3], br = 0.0;
      for (int c = 0; c < 3; c++) bc[c] = 0.5 * (lo[c] + hi[c]);
      for (int s = sb; s < se; s++)
      {
         double d2 = 0.0;
         for (int c = 0; c < 3; c++) d2 += (C.spheres[s].pos[c] - bc[c]) * (C.spheres[s].pos[c] - bc[c]);
         br = std::max(br, sqrt(d2) + C.spheres[s].radius);
      }
      for (int c = 0; c < 3; c++) C.gbound[4 * k + c] = bc[c];
      C.gbound[4 * k + 3] = br * (1.0 + 1e-9) + 1e-9; /* the frames are orthonormal only up to rounding */
   }
   /* the same relation by group: the joints whose subtree carries it */
   C.ganc.assign(C.n_groups + 1, 0);
   for (int g = 0; g < C.n_groups; g++)
   {
      C.ganc[g] = (int) C.ganc.size();
      for (int k = 0; k < nj; k++)
         for (int di = C.joints[k].desc_begin; di < C.joints[k].desc_end; di++)
            if (C.desc[di] == g) C.ganc.push_back(k);
   }
   C.ganc[C.n_groups] = (int) C.ganc.size();
   return OCB_OK;
}

/* The compiled robot as constexpr tables for the run-time compiler (csrc/chomp_jit_robot.cuh turns
 * them into straight-line code).  Returns an empty string when the robot does not suit that form
 * (more pairs than the hit set has bits); the generic run-time specialised kernel is used then. */
std::string jit_robot_header(const CompiledRobot &C, double eps_self)
{
   const int nj = (int) C.joints.size(), nsa = (int) C.spheres.size(), nsi = (int) C.inactive_radius.size();
   const int NAp = nsa + 3, row = NAp + nsi;
   /* static pair list: own sphere ascending, partner ascending, pairs on the same link left out (mod.cpp:1256) */
   std::vector<int> pair_begin(nsa + 1, 0), pair_o, kind;
   std::vector<double> cut2;
   for (int s = 0; s < nsa; s++)
   {
      pair_begin[s] = (int) pair_o.size();
      for (int o = s + 1; o < nsa; o++)
      {
         const double c2 = C.cut2[(size_t) s * row + o];
         if (c2 < 0.0) continue;
         int k = 0;
         if (C.spheres[s].group == C.spheres[o].group)
         {
            /* rigidly attached to the same joint frame: the distance never changes (up to the rounding of the
             * frames, ~1e-15); decided here unless it is within 1e-9 of the cut-off */
            double d2 = 0.0;
            for (int c = 0; c < 3; c++) d2 += (C.spheres[s].pos[c] - C.spheres[o].pos[c]) * (C.spheres[s].pos[c] - C.spheres[o].pos[c]);
            if (d2 <= c2 * (1.0 - 1e-9)) k = 1;
            else if (d2 >= c2 * (1.0 + 1e-9)) k = 2;
         }
         pair_o.push_back(o);
         kind.push_back(k);
         cut2.push_back(c2);
      }
   }
   pair_begin[nsa] = (int) pair_o.size();
   const int npa = (int) pair_o.size();
   for (int s = 0; s < nsa; s++)
      for (int i = 0; i < nsi; i++)
      {
         const double c2 = C.cut2[(size_t) s * row + NAp + i];
         kind.push_back(c2 < 0.0 ? 2 : 0);
         cut2.push_back(c2);
      }
   const int nbits = npa + nsa * nsi;
   if (nbits > 128 || nbits == 0 || nsa + nsi > 255 || C.n_groups > 255) return std::string();
   for (int s = 0; s < nsa; s++)
      if (pair_begin[s + 1] - pair_begin[s] > 32 || nsi > 32) return std::string();
   const int words = (nbits + 31) / 32;
   std::vector<unsigned> always(words, 0u);
   for (int k = 0; k < nbits; k++)
      if (kind[k] == 1) always[k >> 5] |= 1u << (k & 31);
   std::vector<int> own_tests(nsa, 0);
   for (int s = 0; s < nsa; s++)
   {
      for (int k = pair_begin[s]; k < pair_begin[s + 1]; k++) own_tests[s] += (kind[k] == 0);
      for (int i = 0; i < nsi; i++) own_tests[s] += (kind[npa + s * nsi + i] == 0);
   }

   std::string h;
   char buf[128];
   auto def = [&](const char *name, int v) { snprintf(buf, sizeof(buf), "#define %s %d\n", name, v); h += buf; };
   auto dbl = [&](double v)
   {
      if (v == HUGE_VAL) return std::string("(1.0 / 0.0)");
      if (v == -HUGE_VAL) return std::string("(-1.0 / 0.0)");
      snprintf(buf, sizeof(buf), "%a", v); /* hexadecimal floating literal: exact */
      return std::string(buf);
   };
   auto iarr = [&](const char *name, const std::vector<int> &v)
   {
      h += std::string("__device__ constexpr int ") + name + "[" + std::to_string(std::max<size_t>(v.size(), 1)) + "] = {";
      for (size_t i = 0; i < std::max<size_t>(v.size(), 1); i++) h += (i ? ", " : "") + std::to_string(i < v.size() ? v[i] : 0);
      h += "};\n";
   };
   auto darr2 = [&](const char *name, const std::vector<double> &v, int cols)
   {
      const size_t rows = std::max<size_t>(v.size() / cols, 1);
      h += std::string("__device__ constexpr double ") + name + "[" + std::to_string(rows) + "][" + std::to_string(cols) + "] = {";
      for (size_t r = 0; r < rows; r++)
      {
         h += r ? ", {" : "{";
         for (int c = 0; c < cols; c++) h += (c ? ", " : "") + dbl(r * cols + c < v.size() ? v[r * cols + c] : 0.0);
         h += "}";
      }
      h += "};\n";
   };
   auto darr = [&](const char *name, const std::vector<double> &v)
   {
      h += std::string("__device__ constexpr double ") + name + "[" + std::to_string(std::max<size_t>(v.size(), 1)) + "] = {";
      for (size_t i = 0; i < std::max<size_t>(v.size(), 1); i++) h += (i ? ", " : "") + dbl(i < v.size() ? v[i] : 0.0);
      h += "};\n";
   };
   h += "/* generated by ocb_engine.cu (jit_robot_header): one compiled robot */\n";
   (void) eps_self;
   def("JR_NJ", nj); def("JR_NSA", nsa); def("JR_NSI", nsi); def("JR_NSLOTS", C.n_slots); def("JR_NG", C.n_groups);
   def("JR_NPA", npa); def("JR_HIT_WORDS", words);
   std::vector<double> XR, Xt, c0, c1, sp, ip;
   std::vector<int> type, dof, load, save, sb, se, db, de;
   for (int j = 0; j < nj; j++)
   {
      const OcbJointDev &J = C.joints[j];
      XR.insert(XR.end(), J.XR, J.XR + 9);
      Xt.insert(Xt.end(), J.Xt, J.Xt + 3);
      c0.push_back(J.c0); c1.push_back(J.c1);
      type.push_back(J.type); dof.push_back(J.dof); load.push_back(J.load); save.push_back(J.save);
      sb.push_back(J.sph_begin); se.push_back(J.sph_end); db.push_back(J.desc_begin); de.push_back(J.desc_end);
   }
   for (int s = 0; s < nsa; s++) sp.insert(sp.end(), C.spheres[s].pos, C.spheres[s].pos + 3);
   darr2("jr_XR", XR, 9); darr2("jr_Xt", Xt, 3); darr("jr_c0", c0); darr("jr_c1", c1);
   iarr("jr_type", type); iarr("jr_dof", dof); iarr("jr_load", load); iarr("jr_save", save);
   iarr("jr_sph_begin", sb); iarr("jr_sph_end", se); iarr("jr_desc_begin", db); iarr("jr_desc_end", de);
   iarr("jr_desc", C.desc);
   darr2("jr_sph_pos", sp, 3);
   darr2("jr_inactive_pos", C.inactive_pos, 3);
   iarr("jr_pair_begin", pair_begin); iarr("jr_pair_o", pair_o); iarr("jr_pair_kind", kind); darr("jr_pair_cut2", cut2);
   iarr("jr_own_tests", own_tests);
   {
      /* per active pair: the two spheres and their joint frames in one word, and the sum of the radii */
      std::vector<double> rsum;
      h += "__device__ constexpr unsigned jr_pair_info[" + std::to_string(std::max(npa, 1)) + "] = {";
      int k = 0;
      for (int s = 0; s < nsa; s++)
         for (int e = pair_begin[s]; e < pair_begin[s + 1]; e++, k++)
         {
            const int o = pair_o[e];
            const unsigned info = (unsigned) s | ((unsigned) o << 8) | ((unsigned) C.spheres[s].group << 16) | ((unsigned) C.spheres[o].group << 24);
            h += (k ? ", " : "") + std::to_string(info) + "u";
            rsum.push_back(C.spheres[s].radius + C.radius[o]);
         }
      if (npa == 0) h += "0u";
      h += "};\n";
      darr("jr_pair_rsum", rsum);
   }
   {
      std::vector<int> grp;
      for (int s = 0; s < nsa; s++) grp.push_back(C.spheres[s].group);
      iarr("jr_group", grp);
      darr("jr_radius", C.radius);
   }
   h += "__device__ constexpr unsigned jr_hits_always[" + std::to_string(words) + "] = {";
   for (int k = 0; k < words; k++) h += (k ? ", " : "") + std::to_string(always[k]) + "u";
   h += "};\n";
   return h;
}

/* dense inverse of the metric from its banded factor, column by column (stands in for dgetri,
 * chomp.c:393-403; read by the constraint system only) */
void metric_inverse(const Metric &M, std::vector<double> &Ainv)
{
   const int m = M.m, bw = M.bw;
   Ainv.assign((size_t) m * m, 0.0);
   std::vector<double> x(m);
   for (int c = 0; c < m; c++)
   {
      for (int i = 0; i < m; i++) x[i] = (i == c) ? 1.0 : 0.0;
      for (int i = 0; i < m; i++)
      {
         double acc = x[i];
         for (int k = 1; k <= bw && k <= i; k++) acc -= M.Lband[(size_t) i * bw + (k - 1)] * x[i - k];
         x[i] = acc;
      }
      for (int i = m - 1; i >= 0; i--)
      {
         double acc = x[i] * M.dinv[i];
         for (int k = 1; k <= bw && i + k < m; k++) acc -= M.Lband[(size_t) (i + k) * bw + (k - 1)] * x[i + k];
         x[i] = acc;
      }
      for (int i = 0; i < m; i++) Ainv[(size_t) i * m + c] = x[i];
   }
}

Xf xf_invert(const Xf &a)
{
   Xf c = xf_transpose_rot(a);
   for (int r = 0; r < 3; r++) c.t[r] = -(c.R[3 * r] * a.t[0] + c.R[3 * r + 1] * a.t[1] + c.R[3 * r + 2] * a.t[2]);
   return c;
}

/* the `create` command's constraints (mod.cpp:2466-2519 enabled masks, 2571-2613 registration) for the
 * kernel: frames folded onto joint frames, rows stacked waypoint by waypoint */
int compile_constraints(const ocb_params *params, const ocb_robot *rb, const CompiledRobot &C, int m,
                        std::vector<OcbConDev> &cons, std::vector<int> &row0, std::vector<int> &row_wp)
{
   cons.clear();
   for (int ci = 0; ci < params->n_constraints; ci++)
   {
      const ocb_constraint &s = params->constraints[ci];
      if (s.link < 0 || s.link >= rb->n_links) return fail(OCB_ERR_ARG, "constraint %d: bad link", ci);
      if (s.where != OCB_CON_START && s.where != OCB_CON_END && s.where != OCB_CON_ALL && s.where != OCB_CON_START_TSR)
         return fail(OCB_ERR_ARG, "constraint %d: where must be one of OCB_CON_*", ci);
      OcbConDev d;
      memset(&d, 0, sizeof(d));
      const Xf A = xf_invert(xf_from_pose(s.T0w));
      const Xf Cx = xf_mul(xf_mul(C.link_rel[s.link], xf_from_pose(s.pose_link_ee)), xf_invert(xf_from_pose(s.Twe)));
      memcpy(d.AR, A.R, sizeof(d.AR)); memcpy(d.At, A.t, sizeof(d.At));
      memcpy(d.CR, Cx.R, sizeof(d.CR)); memcpy(d.Ct, Cx.t, sizeof(d.Ct));
      d.joint = C.link_joint[s.link];
      d.anc = 0;
      for (int j = d.joint; j >= 0; j = C.joint_parent[j]) d.anc |= 1ull << j;
      d.where = s.where;
      d.k = 0;
      for (int i = 0; i < 6; i++)
         if (s.Bw[i][0] == 0.0 && s.Bw[i][1] == 0.0) d.rows[d.k++] = (i < 3) ? i : 8 - i; /* mod.cpp:1413 */
      if (d.k > 0) cons.push_back(d); /* a constraint without rows adds nothing to the system */
   }
   row0.assign(m + 1, 0);
   row_wp.clear();
   for (int i = 0; i < m; i++)
   {
      row0[i] = (int) row_wp.size();
      for (const OcbConDev &d : cons)
         if (d.where == OCB_CON_ALL || ((d.where == OCB_CON_START || d.where == OCB_CON_START_TSR) && i == 0) ||
             (d.where == OCB_CON_END && i == m - 1))
            for (int r = 0; r < d.k; r++) row_wp.push_back(i);
   }
   row0[m] = (int) row_wp.size();
   return OCB_OK;
}

void seed_mt(uint32_t *mt, unsigned int seed)
{
   /* gsl_rng_set on mt19937: seed 0 -> 4357, 2002 initialisation */
   if (seed == 0) seed = 4357;
   mt[0] = seed;
   for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t) i;
   mt[624] = 624;
}
} /* namespace */

template <class T>
static int batch_alloc(ocb_batch *b, T **ptr, size_t count)
{
   *ptr = nullptr;
   if (pool_alloc(b->e, (void **) ptr, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess)
   {
      cudaGetLastError();
      return fail(OCB_ERR_ALLOC, "device allocation of %zu bytes failed", count * sizeof(T));
   }
   b->owned.push_back(*ptr);
   return OCB_OK;
}

template <class T>
static int batch_upload(ocb_batch *b, const T **ptr, const std::vector<T> &h)
{
   T *d = nullptr;
   int rc = batch_alloc(b, &d, h.size());
   if (rc) return rc;
   if (!h.empty())
      CU(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, b->e->stream));
   *ptr = d;
   return OCB_OK;
}

static ocb_engine *batch_engine(const ocb_batch *b) { return b->e; }

static void batch_free(ocb_batch *b)
{
   cudaSetDevice(b->e->device);
   cudaStreamSynchronize(b->e->stream);
   for (void *p : b->owned) pool_free(b->e, p);
   delete b;
}

extern "C" int ocb_batch_destroy(ocb_batch *b)
{
   if (!b) return OCB_OK;
   {
      std::lock_guard<std::mutex> lock(g_batches_lock);
      if (!g_batches.erase(b)) return OCB_OK; /* already gone with its engine */
   }
   batch_free(b);
   return OCB_OK;
}

extern "C" int ocb_batch_create(ocb_engine *e, const ocb_robot *robot, const ocb_params *params,
                                int n_sdfs, const int *sdf_ids, int n_runs, const double *q_start,
                                const double *q_goal, const unsigned int *seeds, ocb_batch **out)
{
   if (!e || !robot || !params || !out || !q_start || !q_goal) return fail(OCB_ERR_ARG, "null argument");
   *out = nullptr;
   /* argument checks of mod::create (mod.cpp:2091-2101) */
   if (n_sdfs < 1 || !sdf_ids) return fail(OCB_ERR_ARG, "No signed distance fields have yet been computed!");
   if (n_sdfs > OCB_MAX_SDFS) return fail(OCB_ERR_ARG, "more than %d signed distance fields", OCB_MAX_SDFS);
   if (params->lambda < 0.01) return fail(OCB_ERR_ARG, "lambda must be >=0.01!");
   if (params->n_points < 3) return fail(OCB_ERR_ARG, "n_points must be >=3!");
   if (params->derivative < 1 || params->derivative > OCB_MAX_BW)
      return fail(OCB_ERR_ARG, "derivative must be in 1..%d", OCB_MAX_BW);
   if (n_runs < 1) return fail(OCB_ERR_ARG, "need at least one run");
   /* use_hmc without use_momentum: the reference resamples AG and then overwrites it with
    * Ainv G (beta = 0, chomp.c:529-530), so nothing observable changes -> same as no hmc */
   const int use_hmc = (params->use_hmc && params->use_momentum) ? (params->use_hmc == 2 ? 2 : 1) : 0;
   for (int i = 0; i < n_sdfs; i++)
      if (sdf_ids[i] < 0 || sdf_ids[i] >= (int) e->sdfs.size() || !e->sdfs[sdf_ids[i]].used)
         return fail(OCB_ERR_ARG, "bad sdf id %d", sdf_ids[i]);
   CU(cudaSetDevice(e->device));

   CompiledRobot C;
   const bool floating = params->floating_base != 0;
   int rc = compile_robot(robot, params->epsilon_self, floating, C);
   if (rc) return rc;
   /* start_tsr: the start point joins the optimised rows (mod.cpp:2316) */
   int n_start_tsr = 0;
   if (params->n_constraints > 0 && !params->constraints) return fail(OCB_ERR_ARG, "constraints is null");
   for (int i = 0; i < params->n_constraints; i++)
      if (params->constraints[i].where == OCB_CON_START_TSR) n_start_tsr++;
   if (n_start_tsr > 1) return fail(OCB_ERR_ARG, "at most one start_tsr");
   if (n_start_tsr && floating) return fail(OCB_ERR_ARG, "floating_base and start_tsr together is not yet implemented!"); /* mod.cpp:2100 */
   const bool free_start = n_start_tsr > 0;
   const int P = params->n_points, m = P - 2 + (free_start ? 1 : 0), n = robot->n_dof + (floating ? 7 : 0);
   Metric M;
   const double dt = 1.0 / (P - 1); /* mod.cpp:2567 */
   rc = build_metric(m, params->derivative, dt, M, free_start);
   if (rc) return fail(OCB_ERR_ARG, "smoothness metric is not positive definite (code %d)", rc);

   ocb_batch *b = new (std::nothrow) ocb_batch();
   if (!b) return fail(OCB_ERR_ALLOC, "out of host memory");
   b->e = e;
   {
      std::lock_guard<std::mutex> lock(g_batches_lock);
      g_batches.insert(b);
   }
   OcbChompArgs &a = b->args;
   memset(&a, 0, sizeof(a));
   a.R = n_runs; a.P = P; a.m = m; a.n = n;
   a.nj = (int) C.joints.size();
   a.nsa = (int) C.spheres.size();
   a.nsi = (int) C.inactive_radius.size();
   a.nsdf = n_sdfs;
   a.bw = params->derivative;
   a.n_slots = C.n_slots;
   a.ng = C.n_groups;
   a.n_desc = (int) C.desc.size();
   a.NAp = a.nsa + 3;
   a.Ppad = P + (free_start ? 1 : 0); /* the kernel's extra column, see chomp_iterate_body */
   a.free_start = free_start ? 1 : 0;
   a.use_momentum = params->use_momentum ? 1 : 0;
   a.use_hmc = use_hmc;
   a.floating = floating ? 1 : 0;
   for (int j = 0; j < a.nj; j++) a.joints[j] = C.joints[j];
   a.trc_ss = M.ss; a.trc_sg = M.sg; a.trc_gg = M.gg;
   {
      /* the same band in every row?  (entries outside the matrix are never read: band_AT skips them) */
      const int bw = params->derivative, W = 2 * bw + 1, mid = m / 2;
      bool same = m > 2 * bw + 1;
      for (int i = 0; i < m && same; i++)
         for (int k = -bw; k <= bw; k++)
         {
            if (i + k < 0 || i + k >= m) continue;
            if (M.Aband[(size_t) i * W + k + bw] != M.Aband[(size_t) mid * W + k + bw]) { same = false; break; }
         }
      a.band_toeplitz = same ? 1 : 0;
      for (int k = 0; k < W; k++) a.band_row[k] = same ? M.Aband[(size_t) mid * W + k] : 0.0;
      /* c tridiag(-1, 2, -1): the solve is two weighted running sums (chomp_device.cuh: band_solve_121_scan);
       * OCB_SOLVE_LDL=1 in the environment keeps the factorised form (development knob) */
      const char *ldl = getenv("OCB_SOLVE_LDL");
      const double c = same && bw == 1 ? -a.band_row[0] : 0.0;
      a.band_121 = (same && bw == 1 && c > 0.0 && a.band_row[2] == -c && a.band_row[1] == 2.0 * c && !(ldl && ldl[0] == '1')) ? 1 : 0;
      a.band_121_scale = a.band_121 ? 1.0 / ((double) (m + 1) * c) : 0.0;
      if (a.band_121)
      {
         /* and B = -c (q_start e_1 + q_goal e_m): the smoothness part of the update in closed form (bit 1; chomp_iterate_body) */
         bool ends = M.bi[0] == -c && M.bf[m - 1] == -c;
         for (int i = 0; i < m && ends; i++)
            if ((i > 0 && M.bi[i] != 0.0) || (i < m - 1 && M.bf[i] != 0.0)) ends = false;
         const char *lf = getenv("OCB_LINE_FORM");
         if (ends && !(lf && lf[0] == '0')) a.band_121 |= 2;
      }
   }
   a.lambda = params->lambda;
   a.dt = dt;
   a.eps = params->epsilon;
   a.eps_self = params->epsilon_self;
   a.obs_factor = params->obs_factor;
   a.obs_factor_self = params->obs_factor_self;
   a.hmc_lambda = params->hmc_resample_lambda;

   std::vector<OcbSdfDev> sd(n_sdfs);
   for (int i = 0; i < n_sdfs; i++) fill_sdf_dev(e->sdfs[sdf_ids[i]], sd[i]);
   for (int i = 0; i < n_sdfs && i < OCB_INLINE_SDFS; i++) a.sdf_inline[i] = sd[i];

#define TRY(x) do { rc = (x); if (rc) { ocb_batch_destroy(b); return rc; } } while (0)
   TRY(batch_upload(b, &a.spheres, C.spheres));
   TRY(batch_upload(b, &a.cut2, C.cut2));
   TRY(batch_upload(b, &a.radius, C.radius));
   TRY(batch_upload(b, &a.desc, C.desc));
   TRY(batch_upload(b, &a.ganc, C.ganc));
   TRY(batch_upload(b, &a.gbound, C.gbound));
   TRY(batch_upload(b, &a.inactive_pos, C.inactive_pos));
   TRY(batch_upload(b, &a.sdfs, sd));
   TRY(batch_upload(b, &a.Aband, M.Aband));
   TRY(batch_upload(b, &a.Lband, M.Lband));
   TRY(batch_upload(b, &a.dinv, M.dinv));
   TRY(batch_upload(b, &a.bcoef_i, M.bi));
   TRY(batch_upload(b, &a.bcoef_f, M.bf));
   {
      /* the pose entries of a floating base are unbounded (mod.cpp:2640-2652) */
      std::vector<double> lo(n, -HUGE_VAL), hi(n, HUGE_VAL);
      for (int j = 0; j < robot->n_dof; j++)
      {
         lo[(floating ? 7 : 0) + j] = robot->limit_lower[j];
         hi[(floating ? 7 : 0) + j] = robot->limit_upper[j];
      }
      TRY(batch_upload(b, &a.lim_lo, lo));
      TRY(batch_upload(b, &a.lim_hi, hi));
   }
   const size_t R = (size_t) n_runs;
   if (params->n_constraints > 0)
   {
      if (!params->constraints) { ocb_batch_destroy(b); return fail(OCB_ERR_ARG, "constraints is null"); }
      std::vector<OcbConDev> cons;
      std::vector<int> row0, row_wp;
      TRY(compile_constraints(params, robot, C, m, cons, row0, row_wp));
      a.n_con = (int) cons.size();
      a.con_K = (int) row_wp.size();
      if (a.con_K > 0)
      {
         a.con_kmax = 0;
         for (int i = 0; i < m; i++) a.con_kmax = std::max(a.con_kmax, row0[i + 1] - row0[i]);
         a.con_kuniform = a.con_kmax;
         for (int i = 0; i < m; i++)
            if (row0[i + 1] - row0[i] != a.con_kmax) a.con_kuniform = 0;
         /* tridiagonal metric: no dense system (chomp_constraints.cuh); OCB_CON_DENSE=1 keeps the reference's form */
         const char *dense_env = getenv("OCB_CON_DENSE");
         /* a handful of rows (start / end constraints): the dense system is a few small steps, the sweep always m */
         const int dense_rows = 24;
         a.con_fast = (params->derivative == 1 && a.con_K > dense_rows && !(dense_env && dense_env[0] == '1')) ? 1 : 0;
         if (dense_env && dense_env[0] == '0' && params->derivative == 1) a.con_fast = 1; /* OCB_CON_DENSE=0: the sweep whatever the size */
         const size_t jh = (size_t) a.con_K * (n + 1), rec = ocb_con_tridiag_scratch(m, n, a.con_kmax);
         const size_t ws_doubles = (size_t) (3 * a.nsa + 12 * a.n_slots + 6 * a.ng) * a.Ppad;
         /* a sphere model too large for the persistent kernel takes the tiled path (decided below by the same
          * test): its update kernel has no forward sweep's workspace, so everything lives in global scratch,
          * including the branch frames the constraint evaluation rebuilds for itself */
         a.ws_stride = ws_doubles;
         const bool will_tile = ocb_chomp_smem_bytes(&a) > (size_t) e->smem_optin;
         a.con_jh_smem = (!will_tile && a.con_fast && jh <= (size_t) 3 * a.nsa * a.Ppad) ? 1 : 0;
         a.con_rec_off = a.con_jh_smem ? (int) jh : 0;
         a.con_rec_smem = (!will_tile && a.con_fast && a.con_rec_off + rec <= ws_doubles) ? 1 : 0;
         a.con_stride = (size_t) a.con_K * (n + 2) +
                        (a.con_fast ? (a.con_rec_smem ? 0 : rec) : (size_t) a.con_K * a.con_K);
         a.con_slots_off = a.con_stride;
         if (will_tile) a.con_stride += (size_t) 12 * a.n_slots * a.Ppad;
         if (R * a.con_stride * sizeof(double) > ((size_t) 48 << 30))
         {
            ocb_batch_destroy(b);
            return fail(OCB_ERR_ARG, "constraint systems of %d rows for %d runs need %.1f GB of scratch", a.con_K, n_runs,
                        R * a.con_stride * 8.0 / 1e9);
         }
         std::vector<double> Ainv;
         metric_inverse(M, Ainv);
         TRY(batch_upload(b, &a.cons, cons));
         TRY(batch_upload(b, &a.con_row0, row0));
         TRY(batch_upload(b, &a.con_row_wp, row_wp));
         TRY(batch_upload(b, &a.Ainv, Ainv));
         TRY(batch_alloc(b, &a.con_scratch, R * a.con_stride));
         TRY(batch_alloc(b, &a.con_singular, R));
         cudaError_t ce = cudaMemsetAsync(a.con_singular, 0, R * sizeof(int), e->stream);
         if (ce != cudaSuccess) { ocb_batch_destroy(b); return fail(OCB_ERR_CUDA, "batch_create: %s", cudaGetErrorString(ce)); }
      }
   }
   TRY(batch_alloc(b, &a.traj, R * P * n));
   TRY(batch_alloc(b, &a.costs, R * 3));
   TRY(batch_alloc(b, &a.status, R));
   TRY(batch_alloc(b, &a.iters_done, R));
   TRY(batch_alloc(b, &a.limit_rounds, R));
   TRY(batch_alloc(b, &b->d_start, R * n));
   TRY(batch_alloc(b, &b->d_goal, R * n));
   cudaError_t err = cudaMemsetAsync(a.costs, 0, R * 3 * sizeof(double), e->stream);
   if (err == cudaSuccess) err = cudaMemsetAsync(a.status, 0, R * sizeof(int), e->stream);
   if (err == cudaSuccess) err = cudaMemsetAsync(a.iters_done, 0, R * sizeof(int), e->stream);
   if (err == cudaSuccess) err = cudaMemsetAsync(a.limit_rounds, 0, R * sizeof(int), e->stream);
   if (err == cudaSuccess) err = cudaMemcpyAsync(b->d_start, q_start, R * n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess) err = cudaMemcpyAsync(b->d_goal, q_goal, R * n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
   if (err == cudaSuccess) err = ocb_launch_init_traj(a.traj, b->d_start, b->d_goal, n_runs, P, n, a.floating, e->stream);
   e->launches++;
   if (a.use_momentum && err == cudaSuccess)
   {
      TRY(batch_alloc(b, &a.AG, R * m * n));
      TRY(batch_alloc(b, &a.leapfrog_first, R));
      err = cudaMemsetAsync(a.AG, 0, R * m * n * sizeof(double), e->stream); /* chomp.c:114-115 */
      std::vector<int> ones(R, 1);                                            /* chomp.c:88 */
      if (err == cudaSuccess) err = cudaMemcpyAsync(a.leapfrog_first, ones.data(), R * sizeof(int), cudaMemcpyHostToDevice, e->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
   }
   if (a.use_hmc && err == cudaSuccess)
   {
      TRY(batch_alloc(b, &a.hmc_next, R));
      TRY(batch_alloc(b, &a.mt_state, R * 625));
      err = cudaMemsetAsync(a.hmc_next, 0, R * sizeof(int), e->stream); /* mod.cpp:2634 */
      b->mt_host.resize(R * 625);
      std::vector<uint32_t> &st = b->mt_host;
      for (size_t r = 0; r < R; r++) seed_mt(&st[r * 625], seeds ? seeds[r] : 0u);
      if (err == cudaSuccess) err = cudaMemcpyAsync(a.mt_state, st.data(), st.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
   }
   if (err != cudaSuccess)
   {
      ocb_batch_destroy(b);
      return fail(OCB_ERR_CUDA, "batch_create: %s", cudaGetErrorString(err));
   }

   /* workspace: per waypoint 3*nsa sphere coordinates + 12 per saved branch frame
    * + 6 per sphere-carrying joint frame (wrench accumulators).  When a run does not fit in
    * one SM's shared memory the iteration is tiled over waypoints (chomp_tiled.cu). */
   a.ws_stride = (size_t) (3 * a.nsa + 12 * a.n_slots + 6 * a.ng) * a.Ppad;
   b->smem = ocb_chomp_smem_bytes(&a);
   if (b->smem > (size_t) e->smem_optin && a.free_start)
   {
      ocb_batch_destroy(b);
      return fail(OCB_ERR_ARG, "start_tsr needs the run in one SM's shared memory (%d spheres, %d waypoints, %d dofs)",
                  a.nsa, P, n);
   }
   if (b->smem > (size_t) e->smem_optin)
   {
      a.tiled = 1;
      a.tile_w = 0;
      int tw_max = 32;
      if (const char *tw_env = getenv("OCB_TILE_W")) tw_max = std::max(8, std::min(32, atoi(tw_env))); /* development knob */
      for (int tw = tw_max; tw >= 8; tw /= 2)
         if (ocb_tile_smem_bytes(&a, tw) <= (size_t) e->smem_optin) { a.tile_w = tw; break; }
      b->run_smem = ocb_run_update_smem_bytes(&a);
      if (!a.tile_w || b->run_smem > (size_t) e->smem_optin)
      {
         ocb_batch_destroy(b);
         return fail(OCB_ERR_ARG, "problem too large for shared memory (%d spheres, %d waypoints, %d dofs)",
                     a.nsa, P, n);
      }
      b->tile_smem = ocb_tile_smem_bytes(&a, a.tile_w);
      a.n_tiles = (m + a.tile_w - 1) / a.tile_w;
      TRY(batch_alloc(b, &a.G_obs, R * m * n));
      TRY(batch_alloc(b, &a.tile_cost, R * a.n_tiles));
   }
   b->threads = std::min(256, ((a.Ppad + 31) / 32) * 32);
   if (e->jit && !a.tiled && a.con_K == 0 && !a.free_start) /* constraints: the library's kernel */
   {
      /* blocks per SM the shared memory allows, not more than ~160 registers per thread can feed */
      int min_blocks = (int) ((size_t) e->smem_per_sm / (b->smem + 1024));
      min_blocks = std::max(1, std::min(min_blocks, 65536 / (b->threads * 160)));
      if (const char *mb = getenv("OCB_JIT_MINBLOCKS")) min_blocks = std::max(1, atoi(mb)); /* development knob */
      char why[512] = "";
      /* the robot as straight-line code; OCB_JIT_ROBOT=0 in the environment keeps the table-driven kernel */
      const char *env = getenv("OCB_JIT_ROBOT");
      const std::string robot_hdr = (env && env[0] == '0') ? std::string() : jit_robot_header(C, params->epsilon_self);
      const size_t smem_generic = b->smem;
      if (!robot_hdr.empty())
      {
         a.robot_smem = 1; /* that kernel lays shared memory out differently (smem_layout) */
         b->smem = ocb_chomp_smem_bytes(&a);
         if (!a.trig_cache) TRY(batch_alloc(b, &a.trig_cache, R * 2 * a.nj * a.Ppad));
      }
      if (ocb_jit_chomp_kernel(&a, e->device, b->threads, min_blocks, b->smem, robot_hdr.c_str(), &b->jit_kernel, why,
                               sizeof(why)) != 0)
      {
         b->jit_kernel = nullptr; /* the library's own kernel runs instead */
         a.robot_smem = 0;
         b->smem = smem_generic;
         fail(OCB_ERR_CUDA, "run-time specialisation unavailable: %s", why);
      }
   }
   err = cudaStreamSynchronize(e->stream);
   if (err != cudaSuccess)
   {
      ocb_batch_destroy(b);
      return fail(OCB_ERR_CUDA, "batch_create: %s", cudaGetErrorString(err));
   }
#undef TRY
   *out = b;
   return OCB_OK;
}

/* the generated robot tables of the run-time specialised kernel (host only; for inspection and tests):
 * returns the length of the text (0: this robot takes the generic kernel), copies at most cap-1 bytes */
extern "C" long ocb_debug_jit_robot_header(const ocb_robot *robot, const ocb_params *params, char *buf, size_t cap)
{
   if (!robot || !params) return fail(OCB_ERR_ARG, "null argument");
   CompiledRobot C;
   int rc = compile_robot(robot, params->epsilon_self, params->floating_base != 0, C);
   if (rc) return rc;
   const std::string h = jit_robot_header(C, params->epsilon_self);
   if (buf && cap)
   {
      const size_t n = std::min(h.size(), cap - 1);
      memcpy(buf, h.data(), n);
      buf[n] = 0;
   }
   return (long) h.size();
}

/* the smoothness metric of a run as the engine builds it (host only; for inspection and tests): band of A
 * [m][2D+1], dense inverse [m][m] from the banded factor, B's two coefficient vectors [m] each, and whether
 * the closed forms of the default metric apply (bit 0: A = c tridiag(-1, 2, -1); bit 1: B = -c (q_s e_1 + q_g e_m)).
 * free_start: the start_tsr layout (m = n_points - 1, no initial boundary row).  Any output may be NULL. */
extern "C" int ocb_debug_metric(int n_points, int derivative, int free_start, double *Aband, double *Ainv, double *bi,
                                double *bf, int *closed_form, double *c_out)
{
   if (n_points < 3 || derivative < 1 || derivative > OCB_MAX_BW) return fail(OCB_ERR_ARG, "bad metric shape");
   const int m = n_points - 2 + (free_start ? 1 : 0);
   Metric M;
   const int rc = build_metric(m, derivative, 1.0 / (n_points - 1), M, free_start != 0);
   if (rc) return fail(OCB_ERR_ARG, "smoothness metric is not positive definite (code %d)", rc);
   if (Aband) memcpy(Aband, M.Aband.data(), M.Aband.size() * sizeof(double));
   if (Ainv)
   {
      std::vector<double> inv;
      metric_inverse(M, inv);
      memcpy(Ainv, inv.data(), inv.size() * sizeof(double));
   }
   if (bi) memcpy(bi, M.bi.data(), m * sizeof(double));
   if (bf) memcpy(bf, M.bf.data(), m * sizeof(double));
   if (closed_form || c_out)
   {
      const int bw = derivative, W = 2 * bw + 1, mid = m / 2;
      bool same = m > 2 * bw + 1;
      for (int i = 0; i < m && same; i++)
         for (int k = -bw; k <= bw; k++)
         {
            if (i + k < 0 || i + k >= m) continue;
            if (M.Aband[(size_t) i * W + k + bw] != M.Aband[(size_t) mid * W + k + bw]) { same = false; break; }
         }
      const double c = (same && bw == 1) ? -M.Aband[(size_t) mid * W] : 0.0;
      int flags = (same && bw == 1 && c > 0.0 && M.Aband[(size_t) mid * W + 2] == -c && M.Aband[(size_t) mid * W + 1] == 2.0 * c) ? 1 : 0;
      if (flags)
      {
         bool ends = M.bi[0] == -c && M.bf[m - 1] == -c;
         for (int i = 0; i < m && ends; i++)
            if ((i > 0 && M.bi[i] != 0.0) || (i < m - 1 && M.bf[i] != 0.0)) ends = false;
         if (ends) flags |= 2;
      }
      if (closed_form) *closed_form = flags;
      if (c_out) *c_out = c;
   }
   return OCB_OK;
}

extern "C" int ocb_batch_uses_jit(const ocb_batch *b) { return (b && b->jit_kernel) ? 1 : 0; }
extern "C" int ocb_batch_tile_width(const ocb_batch *b) { return (b && b->args.tiled) ? b->args.tile_w : 0; }

extern "C" int ocb_batch_dims(const ocb_batch *b, int *n_runs, int *n_points, int *n_dof)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   if (n_runs) *n_runs = b->args.R;
   if (n_points) *n_points = b->args.P;
   if (n_dof) *n_dof = b->args.n;
   return OCB_OK;
}

extern "C" int ocb_batch_reset(ocb_batch *b, const double *q_start, const double *q_goal,
                               const unsigned int *seeds)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   if ((q_start == nullptr) != (q_goal == nullptr)) return fail(OCB_ERR_ARG, "pass both end points or neither");
   CU(cudaSetDevice(b->e->device));
   OcbChompArgs &a = b->args;
   const size_t R = (size_t) a.R;
   cudaStream_t st = b->e->stream;
   if (q_start)
   {
      CU(cudaMemcpyAsync(b->d_start, q_start, R * a.n * sizeof(double), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(b->d_goal, q_goal, R * a.n * sizeof(double), cudaMemcpyHostToDevice, st));
   }
   CU(ocb_launch_init_traj(a.traj, b->d_start, b->d_goal, a.R, a.P, a.n, a.floating, st));
   b->e->launches++;
   CU(cudaMemsetAsync(a.costs, 0, R * 3 * sizeof(double), st));
   CU(cudaMemsetAsync(a.status, 0, R * sizeof(int), st));
   CU(cudaMemsetAsync(a.iters_done, 0, R * sizeof(int), st));
   CU(cudaMemsetAsync(a.limit_rounds, 0, R * sizeof(int), st));
   if (a.use_momentum)
   {
      CU(cudaMemsetAsync(a.AG, 0, R * a.m * a.n * sizeof(double), st));
      std::vector<int> ones(R, 1);
      CU(cudaMemcpyAsync(a.leapfrog_first, ones.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st)); /* `ones` is a host temporary */
   }
   if (a.use_hmc)
   {
      if (seeds)
         for (size_t r = 0; r < R; r++) seed_mt(&b->mt_host[r * 625], seeds[r]);
      CU(cudaMemsetAsync(a.hmc_next, 0, R * sizeof(int), st));
      CU(cudaMemcpyAsync(a.mt_state, b->mt_host.data(), b->mt_host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
   }
   return OCB_OK;
}

extern "C" int ocb_batch_set_traj(ocb_batch *b, const double *traj)
{
   if (!b || !traj) return fail(OCB_ERR_ARG, "null argument");
   CU(cudaSetDevice(b->e->device));
   const size_t bytes = (size_t) b->args.R * b->args.P * b->args.n * sizeof(double);
   CU(cudaMemcpyAsync(b->args.traj, traj, bytes, cudaMemcpyHostToDevice, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_set_momentum(ocb_batch *b, const double *AG, const int *leapfrog_first)
{
   if (!b || !AG || !leapfrog_first) return fail(OCB_ERR_ARG, "null argument");
   if (!b->args.use_momentum) return fail(OCB_ERR_ARG, "the batch was created without use_momentum");
   CU(cudaSetDevice(b->e->device));
   const OcbChompArgs &a = b->args;
   CU(cudaMemcpyAsync(a.AG, AG, (size_t) a.R * a.m * a.n * sizeof(double), cudaMemcpyHostToDevice, b->e->stream));
   CU(cudaMemcpyAsync(a.leapfrog_first, leapfrog_first, (size_t) a.R * sizeof(int), cudaMemcpyHostToDevice, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_get_momentum(ocb_batch *b, double *AG, int *leapfrog_first)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   if (!b->args.use_momentum) return fail(OCB_ERR_ARG, "the batch was created without use_momentum");
   CU(cudaSetDevice(b->e->device));
   const OcbChompArgs &a = b->args;
   if (AG) CU(cudaMemcpyAsync(AG, a.AG, (size_t) a.R * a.m * a.n * sizeof(double), cudaMemcpyDeviceToHost, b->e->stream));
   if (leapfrog_first)
      CU(cudaMemcpyAsync(leapfrog_first, a.leapfrog_first, (size_t) a.R * sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_set_lambda(ocb_batch *b, double lambda)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   if (!(lambda > 0.0)) return fail(OCB_ERR_ARG, "lambda must be positive");
   b->args.lambda = lambda;
   return OCB_OK;
}

extern "C" int ocb_batch_enable_trace(ocb_batch *b, int enable)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   b->args.trace_on = enable ? 1 : 0;
   return OCB_OK;
}

extern "C" int ocb_batch_capture_gradient(ocb_batch *b, int mode)
{
   if (!b || mode < 0 || mode > 2) return fail(OCB_ERR_ARG, "bad argument");
   CU(cudaSetDevice(b->e->device));
   if (mode && !b->args.grad_out)
   {
      int rc = batch_alloc(b, &b->args.grad_out, (size_t) b->args.R * b->args.m * b->args.n);
      if (rc) return rc;
   }
   b->args.grad_mode = mode;
   return OCB_OK;
}

extern "C" int ocb_batch_iterate_async(ocb_batch *b, int n_iter) { return ocb_batch_iterate_from_async(b, 0, n_iter); }

extern "C" int ocb_batch_iterate_from(ocb_batch *b, int first_iter, int n_iter, double *cost_total, double *cost_obs,
                                      double *cost_smooth, int *status)
{
   int rc = ocb_batch_iterate_from_async(b, first_iter, n_iter);
   if (rc) return rc;
   return ocb_batch_get_costs(b, cost_total, cost_obs, cost_smooth, status);
}

extern "C" int ocb_batch_iterate_from_async(ocb_batch *b, int first_iter, int n_iter)
{
   if (!b) return fail(OCB_ERR_ARG, "you must pass a created run!");
   if (n_iter < 0) return fail(OCB_ERR_ARG, "n_iter must be >=0!");
   if (first_iter < 0) return fail(OCB_ERR_ARG, "first_iter must be >=0!");
   b->args.iter_base = first_iter;
   CU(cudaSetDevice(b->e->device));
   OcbChompArgs &a = b->args;
   if (a.trace_on && n_iter > b->trace_cap)
   {
      int rc = batch_alloc(b, &a.trace, (size_t) a.R * n_iter * 3);
      if (rc) return rc;
      b->trace_cap = n_iter;
   }
   a.n_iter = n_iter;
   b->last_n_iter = n_iter;
   if (a.tiled)
   {
      /* the persistent kernel restarts status and costs on every call; same here */
      CU(cudaMemsetAsync(a.costs, 0, (size_t) a.R * 3 * sizeof(double), b->e->stream));
      CU(cudaMemsetAsync(a.status, 0, (size_t) a.R * sizeof(int), b->e->stream));
      CU(cudaMemsetAsync(a.iters_done, 0, (size_t) a.R * sizeof(int), b->e->stream));
      CU(cudaMemsetAsync(a.limit_rounds, 0, (size_t) a.R * sizeof(int), b->e->stream));
      CU(ocb_launch_chomp_tiled(&a, b->tile_smem, b->run_smem, 256, b->e->stream, &b->e->launches));
      return OCB_OK;
   }
   if (b->jit_kernel)
      CU(ocb_jit_launch(b->jit_kernel, &a, b->threads, b->smem, b->e->stream));
   else
      CU(ocb_launch_chomp(&a, b->smem, b->threads, b->e->stream));
   b->e->launches++;
   return OCB_OK;
}

extern "C" int ocb_batch_get_costs(ocb_batch *b, double *cost_total, double *cost_obs, double *cost_smooth, int *status)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   CU(cudaSetDevice(b->e->device));
   const int R = b->args.R;
   std::vector<double> c((size_t) R * 3);
   CU(cudaMemcpyAsync(c.data(), b->args.costs, c.size() * sizeof(double), cudaMemcpyDeviceToHost, b->e->stream));
   if (status) CU(cudaMemcpyAsync(status, b->args.status, R * sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   for (int r = 0; r < R; r++)
   {
      if (cost_total) cost_total[r] = c[(size_t) r * 3];
      if (cost_obs) cost_obs[r] = c[(size_t) r * 3 + 1];
      if (cost_smooth) cost_smooth[r] = c[(size_t) r * 3 + 2];
   }
   return OCB_OK;
}

extern "C" int ocb_batch_get_iterations(ocb_batch *b, int *iterations)
{
   if (!b || !iterations) return fail(OCB_ERR_ARG, "null argument");
   CU(cudaSetDevice(b->e->device));
   CU(cudaMemcpyAsync(iterations, b->args.iters_done, (size_t) b->args.R * sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_get_limit_rounds(ocb_batch *b, int *rounds)
{
   if (!b || !rounds) return fail(OCB_ERR_ARG, "null argument");
   CU(cudaSetDevice(b->e->device));
   CU(cudaMemcpyAsync(rounds, b->args.limit_rounds, (size_t) b->args.R * sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_get_constraint_skips(ocb_batch *b, int *skips)
{
   if (!b || !skips) return fail(OCB_ERR_ARG, "null argument");
   CU(cudaSetDevice(b->e->device));
   if (!b->args.con_singular)
   {
      for (int r = 0; r < b->args.R; r++) skips[r] = 0;
      return OCB_OK;
   }
   CU(cudaMemcpyAsync(skips, b->args.con_singular, (size_t) b->args.R * sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_iterate(ocb_batch *b, int n_iter, double *cost_total, double *cost_obs,
                                 double *cost_smooth, int *status)
{
   int rc = ocb_batch_iterate_async(b, n_iter);
   if (rc) return rc;
   return ocb_batch_get_costs(b, cost_total, cost_obs, cost_smooth, status);
}

extern "C" int ocb_batch_get_trace(ocb_batch *b, double *trace, int n_iter)
{
   if (!b || !trace) return fail(OCB_ERR_ARG, "null argument");
   if (!b->args.trace_on || n_iter != b->last_n_iter || !b->args.trace)
      return fail(OCB_ERR_ARG, "no trace recorded for %d iterations", n_iter);
   CU(cudaSetDevice(b->e->device));
   CU(cudaMemcpyAsync(trace, b->args.trace, (size_t) b->args.R * n_iter * 3 * sizeof(double),
                      cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_get_traj(ocb_batch *b, double *traj)
{
   if (!b || !traj) return fail(OCB_ERR_ARG, "null argument");
   CU(cudaSetDevice(b->e->device));
   const size_t bytes = (size_t) b->args.R * b->args.P * b->args.n * sizeof(double);
   CU(cudaMemcpyAsync(traj, b->args.traj, bytes, cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_get_gradient(ocb_batch *b, double *G)
{
   if (!b || !G) return fail(OCB_ERR_ARG, "null argument");
   if (!b->args.grad_mode || !b->args.grad_out) return fail(OCB_ERR_ARG, "gradient capture is off");
   CU(cudaSetDevice(b->e->device));
   const size_t bytes = (size_t) b->args.R * b->args.m * b->args.n * sizeof(double);
   CU(cudaMemcpyAsync(G, b->args.grad_out, bytes, cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   return OCB_OK;
}

extern "C" int ocb_batch_best(ocb_batch *b, int *best_run, double *best_cost)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   CU(cudaSetDevice(b->e->device));
   int rc = engine_scratch(b->e, 64);
   if (rc) return rc;
   int *d_idx = (int *) b->e->scratch;
   double *d_cost = (double *) ((char *) b->e->scratch + 8);
   CU(ocb_launch_best(b->args.costs, b->args.status, b->args.R, d_idx, d_cost, b->e->stream));
   b->e->launches++;
   int idx = -1;
   double c = 0;
   CU(cudaMemcpyAsync(&idx, d_idx, sizeof(int), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaMemcpyAsync(&c, d_cost, sizeof(double), cudaMemcpyDeviceToHost, b->e->stream));
   CU(cudaStreamSynchronize(b->e->stream));
   if (best_run) *best_run = idx;
   if (best_cost) *best_cost = c;
   return OCB_OK;
}

extern "C" int ocb_batch_device_ptrs(ocb_batch *b, void **d_traj, void **d_costs)
{
   if (!b) return fail(OCB_ERR_ARG, "null batch");
   if (d_traj) *d_traj = b->args.traj;
   if (d_costs) *d_costs = b->args.costs;
   return OCB_OK;
}

extern "C" int ocb_batch_copy_run_traj_device(ocb_batch *b, int run, void *d_dst)
{
   if (!b || !d_dst || run < 0 || run >= b->args.R) return fail(OCB_ERR_ARG, "bad argument");
   CU(cudaSetDevice(b->e->device));
   const size_t bytes = (size_t) b->args.P * b->args.n * sizeof(double);
   CU(cudaMemcpyAsync(d_dst, b->args.traj + (size_t) run * b->args.P * b->args.n, bytes,
                      cudaMemcpyDeviceToDevice, b->e->stream));
   return OCB_OK;
}
