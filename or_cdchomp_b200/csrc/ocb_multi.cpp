/* ocb_multi.cpp -- one process, several GPUs: the `createbatch` path of SURVEY.md section 8e behind the C ABI.
 *
 * Runs are independent (they share only read-only data: robot, fields, metric), so a batch of R runs
 * is dealt round-robin over G engines (global run r lives on device r mod G, so heavy runs spread
 * evenly), every field is replicated per device and nothing is exchanged during the iterations.
 * Each entry point drives all engines concurrently, one host thread per engine (the threading rule of
 * include/orcdchomp_b200.h), and returns results in global run order.  The one exchange of the path is
 * the final arg-min over cost_total:
 *   - each device reduces its own runs (best_kernel), one (cost, global id) pair per device;
 *   - with NCCL (libnccl.so.2, looked up with dlopen; needs distinct devices) the pairs are
 *     all-gathered and the winner's trajectory broadcast over NVLink, so that afterwards EVERY device
 *     holds it (ncclAllGather + ncclBroadcast inside one group per device thread);
 *   - without NCCL (or with two engines on one device) the host compares the G pairs and the
 *     winner's trajectory is copied from its owner (cudaMemcpyPeer between devices).
 * Ties go to the lowest global run id, so the answer does not depend on the number of devices.
 * The reference has no counterpart: its module advances runs one after the other
 * (src/orcdchomp_mod.cpp:2752-2828, one `struct run` per iterate command).
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orcdchomp_b200.h"

namespace
{
/* ---- the handful of NCCL entry points this file uses, bound at run time ---- */
typedef struct ncclComm *ncclComm_t;
struct Nccl
{
   void *handle = nullptr;
   bool tried = false;
   int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
   int (*CommDestroy)(ncclComm_t) = nullptr;
   int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
};
Nccl g_nccl;
const int NCCL_FLOAT64 = 8; /* ncclDouble (nccl.h: ncclFloat64 = 8) */

bool load_nccl()
{
   if (g_nccl.tried) return g_nccl.handle != nullptr;
   g_nccl.tried = true;
   const char *env = getenv("OCB_MULTI_NCCL");
   if (env && env[0] == '0') return false;
   for (const char *n : {"libnccl.so.2", "libnccl.so"})
   {
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (g_nccl.handle) break;
   }
   if (!g_nccl.handle) return false;
#define SYM(field, name)                                                                                     \
   *(void **) (&g_nccl.field) = dlsym(g_nccl.handle, name);                                                  \
   if (!g_nccl.field) { dlclose(g_nccl.handle); g_nccl.handle = nullptr; return false; }
   SYM(CommInitAll, "ncclCommInitAll")
   SYM(CommDestroy, "ncclCommDestroy")
   SYM(AllGather, "ncclAllGather")
   SYM(Broadcast, "ncclBroadcast")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
   return true;
}

thread_local char g_merr[512] = "";
int mfail(int code, const char *msg)
{
   snprintf(g_merr, sizeof(g_merr), "%s", msg);
   return code;
}

/* f(i) on one host thread per engine; the first non-zero return code (lowest i) is reported */
int for_each_engine(int G, const std::function<int(int)> &f, std::vector<std::string> *errs = nullptr)
{
   std::vector<int> rc(G, 0);
   std::vector<std::string> msg(G);
   std::vector<std::thread> th;
   for (int i = 0; i < G; i++)
      th.emplace_back([&, i]()
      {
         rc[i] = f(i);
         if (rc[i]) msg[i] = ocb_last_error(); /* thread-local in the engine library */
      });
   for (auto &t : th) t.join();
   for (int i = 0; i < G; i++)
      if (rc[i])
      {
         snprintf(g_merr, sizeof(g_merr), "device slot %d: %s", i, msg[i].c_str());
         if (errs) *errs = msg;
         return rc[i];
      }
   return OCB_OK;
}
} /* namespace */

struct ocb_multi
{
   int G = 0;
   std::vector<int> devices;
   std::vector<ocb_engine *> engines;
   std::vector<cudaStream_t> streams; /* one per engine: the engines run on them, NCCL is enqueued on them */
   std::vector<ncclComm_t> comms;     /* empty: host / peer-copy gather */
   std::vector<double *> d_pairs;     /* [G][2] per device: all-gathered (cost, global id) */
   std::vector<double *> d_pair;      /* [2] per device: this device's candidate */
   std::vector<double *> d_traj;      /* [P*n] per device: the winner's trajectory after a gather */
   std::vector<size_t> traj_cap;
};

struct ocb_multi_batch
{
   ocb_multi *m = nullptr;
   int R = 0, P = 0, n = 0;
   std::vector<ocb_batch *> batches;      /* one per engine, NULL when that engine got no run */
   std::vector<std::vector<int>> global;  /* global run ids per engine, ascending */
};

extern "C" const char *ocb_multi_last_error(void) { return g_merr; }

extern "C" int ocb_multi_destroy(ocb_multi *m)
{
   if (!m) return OCB_OK;
   for (int i = 0; i < (int) m->engines.size(); i++)
   {
      if (i < (int) m->devices.size()) cudaSetDevice(m->devices[i]);
      if (i < (int) m->comms.size() && m->comms[i]) g_nccl.CommDestroy(m->comms[i]);
      if (i < (int) m->d_pairs.size() && m->d_pairs[i]) cudaFree(m->d_pairs[i]);
      if (i < (int) m->d_pair.size() && m->d_pair[i]) cudaFree(m->d_pair[i]);
      if (i < (int) m->d_traj.size() && m->d_traj[i]) cudaFree(m->d_traj[i]);
      if (m->engines[i]) ocb_engine_destroy(m->engines[i]); /* also destroys the batches it still owns */
      if (i < (int) m->streams.size() && m->streams[i]) cudaStreamDestroy(m->streams[i]);
   }
   delete m;
   return OCB_OK;
}

extern "C" int ocb_multi_create(int n_devices, const int *devices, ocb_multi **out)
{
   if (!out || n_devices < 1 || n_devices > 64) return mfail(OCB_ERR_ARG, "bad argument");
   *out = nullptr;
   ocb_multi *m = new ocb_multi();
   m->G = n_devices;
   for (int i = 0; i < n_devices; i++) m->devices.push_back(devices ? devices[i] : i);
   m->engines.assign(n_devices, nullptr);
   m->streams.assign(n_devices, nullptr);
   m->d_pairs.assign(n_devices, nullptr);
   m->d_pair.assign(n_devices, nullptr);
   m->d_traj.assign(n_devices, nullptr);
   m->traj_cap.assign(n_devices, 0);
   for (int i = 0; i < n_devices; i++)
   {
      int rc = ocb_engine_create(m->devices[i], &m->engines[i]);
      if (rc)
      {
         snprintf(g_merr, sizeof(g_merr), "device %d: %s", m->devices[i], ocb_last_error());
         ocb_multi_destroy(m);
         return rc;
      }
      cudaSetDevice(m->devices[i]);
      if (cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
          cudaMalloc((void **) &m->d_pairs[i], sizeof(double) * 2 * n_devices) != cudaSuccess ||
          cudaMalloc((void **) &m->d_pair[i], sizeof(double) * 2) != cudaSuccess)
      {
         ocb_multi_destroy(m);
         return mfail(OCB_ERR_ALLOC, "stream / buffer creation failed");
      }
      ocb_engine_set_stream(m->engines[i], (void *) m->streams[i]);
   }
   /* NCCL needs every rank on its own device */
   bool distinct = n_devices > 1;
   for (int i = 0; i < n_devices; i++)
      for (int j = 0; j < i; j++)
         if (m->devices[i] == m->devices[j]) distinct = false;
   if (distinct && load_nccl())
   {
      m->comms.assign(n_devices, nullptr);
      if (g_nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data()) != 0) m->comms.clear();
   }
   /* direct copies between the devices (NVLink) for the fallback gather */
   for (int i = 0; i < n_devices; i++)
      for (int j = 0; j < n_devices; j++)
         if (m->devices[i] != m->devices[j])
         {
            int can = 0;
            cudaSetDevice(m->devices[i]);
            if (cudaDeviceCanAccessPeer(&can, m->devices[i], m->devices[j]) == cudaSuccess && can)
               if (cudaDeviceEnablePeerAccess(m->devices[j], 0) != cudaSuccess) cudaGetLastError();
         }
   *out = m;
   return OCB_OK;
}

extern "C" int ocb_multi_device_count(const ocb_multi *m) { return m ? m->G : 0; }
extern "C" int ocb_multi_uses_nccl(const ocb_multi *m) { return (m && !m->comms.empty()) ? 1 : 0; }

extern "C" int ocb_multi_engine(ocb_multi *m, int slot, ocb_engine **e)
{
   if (!m || !e || slot < 0 || slot >= m->G) return mfail(OCB_ERR_ARG, "bad argument");
   *e = m->engines[slot];
   return OCB_OK;
}

extern "C" int ocb_multi_enable_jit(ocb_multi *m, int on)
{
   if (!m) return mfail(OCB_ERR_ARG, "null handle");
   for (ocb_engine *e : m->engines) ocb_engine_enable_jit(e, on);
   return OCB_OK;
}

/* replicate a field on every device; the id is the same everywhere (engines are driven in lock step) */
extern "C" int ocb_multi_sdf_upload(ocb_multi *m, const ocb_sdf *sdf, int *id)
{
   if (!m || !sdf || !id) return mfail(OCB_ERR_ARG, "null argument");
   std::vector<int> ids(m->G, -1);
   int rc = for_each_engine(m->G, [&](int i) { return ocb_sdf_upload(m->engines[i], sdf, &ids[i]); });
   if (rc) return rc;
   for (int i = 1; i < m->G; i++)
      if (ids[i] != ids[0]) return mfail(OCB_ERR_ARG, "field ids diverged between devices: drive the engines only through ocb_multi_*");
   *id = ids[0];
   return OCB_OK;
}

/* build a field on every device from the same primitives (cheaper than shipping 512 MB) */
extern "C" int ocb_multi_computedistancefield_resident(ocb_multi *m, const ocb_prim *prims, int n_prims,
                                                       const int sizes[3], const double lengths[3], double cube_extent,
                                                       const double pose_world_gsdf[7], int *id)
{
   if (!m || !id) return mfail(OCB_ERR_ARG, "null argument");
   std::vector<int> ids(m->G, -1);
   int rc = for_each_engine(m->G, [&](int i)
   {
      return ocb_computedistancefield_resident(m->engines[i], prims, n_prims, sizes, lengths, cube_extent,
                                               pose_world_gsdf, &ids[i]);
   });
   if (rc) return rc;
   for (int i = 1; i < m->G; i++)
      if (ids[i] != ids[0]) return mfail(OCB_ERR_ARG, "field ids diverged between devices");
   *id = ids[0];
   return OCB_OK;
}

extern "C" int ocb_multi_sdf_remove(ocb_multi *m, int id)
{
   if (!m) return mfail(OCB_ERR_ARG, "null handle");
   return for_each_engine(m->G, [&](int i) { return ocb_sdf_remove(m->engines[i], id); });
}

extern "C" int ocb_multi_batch_destroy(ocb_multi_batch *b)
{
   if (!b) return OCB_OK;
   for (ocb_batch *x : b->batches)
      if (x) ocb_batch_destroy(x);
   delete b;
   return OCB_OK;
}

extern "C" int ocb_multi_batch_create(ocb_multi *m, const ocb_robot *robot, const ocb_params *params, int n_sdfs,
                                      const int *sdf_ids, int n_runs, const double *q_start, const double *q_goal,
                                      const unsigned int *seeds, ocb_multi_batch **out)
{
   if (!m || !robot || !params || !out || !q_start || !q_goal || n_runs < 1) return mfail(OCB_ERR_ARG, "bad argument");
   *out = nullptr;
   ocb_multi_batch *b = new ocb_multi_batch();
   b->m = m;
   b->R = n_runs;
   b->P = params->n_points;
   b->n = robot->n_dof + (params->floating_base ? 7 : 0);
   b->batches.assign(m->G, nullptr);
   b->global.assign(m->G, std::vector<int>());
   for (int r = 0; r < n_runs; r++) b->global[r % m->G].push_back(r);
   const int n = b->n;
   int rc = for_each_engine(m->G, [&](int i)
   {
      const std::vector<int> &g = b->global[i];
      if (g.empty()) return (int) OCB_OK;
      std::vector<double> qs(g.size() * n), qg(g.size() * n);
      std::vector<unsigned int> sd(g.size(), 0u);
      for (size_t k = 0; k < g.size(); k++)
      {
         memcpy(&qs[k * n], q_start + (size_t) g[k] * n, n * sizeof(double));
         memcpy(&qg[k * n], q_goal + (size_t) g[k] * n, n * sizeof(double));
         if (seeds) sd[k] = seeds[g[k]];
      }
      return ocb_batch_create(m->engines[i], robot, params, n_sdfs, sdf_ids, (int) g.size(), qs.data(), qg.data(),
                              seeds ? sd.data() : nullptr, &b->batches[i]);
   });
   if (rc)
   {
      ocb_multi_batch_destroy(b);
      return rc;
   }
   *out = b;
   return OCB_OK;
}

extern "C" int ocb_multi_batch_dims(const ocb_multi_batch *b, int *n_runs, int *n_points, int *n_dof)
{
   if (!b) return mfail(OCB_ERR_ARG, "null batch");
   if (n_runs) *n_runs = b->R;
   if (n_points) *n_points = b->P;
   if (n_dof) *n_dof = b->n;
   return OCB_OK;
}

/* `iterate run ... n_iter N` on every device at once; outputs [R] in global run order, each may be NULL */
extern "C" int ocb_multi_batch_iterate(ocb_multi_batch *b, int n_iter, double *cost_total, double *cost_obs,
                                       double *cost_smooth, int *status)
{
   if (!b) return mfail(OCB_ERR_ARG, "null batch");
   ocb_multi *m = b->m;
   return for_each_engine(m->G, [&](int i)
   {
      if (!b->batches[i]) return (int) OCB_OK;
      const std::vector<int> &g = b->global[i];
      std::vector<double> ct(g.size()), co(g.size()), cs(g.size());
      std::vector<int> st(g.size());
      int rc = ocb_batch_iterate(b->batches[i], n_iter, ct.data(), co.data(), cs.data(), st.data());
      if (rc) return rc;
      for (size_t k = 0; k < g.size(); k++)
      {
         if (cost_total) cost_total[g[k]] = ct[k];
         if (cost_obs) cost_obs[g[k]] = co[k];
         if (cost_smooth) cost_smooth[g[k]] = cs[k];
         if (status) status[g[k]] = st[k];
      }
      return (int) OCB_OK;
   });
}

/* gettraj of every run: [R][n_points][n_dof] in global run order */
extern "C" int ocb_multi_batch_get_traj(ocb_multi_batch *b, double *traj)
{
   if (!b || !traj) return mfail(OCB_ERR_ARG, "null argument");
   ocb_multi *m = b->m;
   const size_t row = (size_t) b->P * b->n;
   return for_each_engine(m->G, [&](int i)
   {
      if (!b->batches[i]) return (int) OCB_OK;
      const std::vector<int> &g = b->global[i];
      std::vector<double> t(g.size() * row);
      int rc = ocb_batch_get_traj(b->batches[i], t.data());
      if (rc) return rc;
      for (size_t k = 0; k < g.size(); k++) memcpy(traj + (size_t) g[k] * row, &t[k * row], row * sizeof(double));
      return (int) OCB_OK;
   });
}

/* The one exchange of the path: global arg-min of cost_total (failed runs skipped, ties to the lowest
 * global id) and the winner's trajectory.  best_run = -1 when every run failed.  traj: host
 * [n_points][n_dof] or NULL.  With NCCL every device afterwards holds the trajectory in its own
 * buffer (ocb_multi_best_traj_device). */
extern "C" int ocb_multi_batch_best(ocb_multi_batch *b, int *best_run, double *best_cost, double *traj)
{
   if (!b || !best_run || !best_cost) return mfail(OCB_ERR_ARG, "null argument");
   ocb_multi *m = b->m;
   const int G = m->G;
   const size_t row = (size_t) b->P * b->n;
   std::vector<int> local_idx(G, -1);
   std::vector<double> local_cost(G, HUGE_VAL);
   int rc = for_each_engine(G, [&](int i)
   {
      if (!b->batches[i]) return (int) OCB_OK;
      return ocb_batch_best(b->batches[i], &local_idx[i], &local_cost[i]);
   });
   if (rc) return rc;
   int win = -1, win_id = -1;
   double win_cost = HUGE_VAL;
   for (int i = 0; i < G; i++)
   {
      if (local_idx[i] < 0) continue;
      const int gid = b->global[i][local_idx[i]];
      if (local_cost[i] < win_cost || (local_cost[i] == win_cost && gid < win_id)) { win = i; win_cost = local_cost[i]; win_id = gid; }
   }
   *best_run = win_id;
   *best_cost = win_cost;
   /* room for the winner's trajectory on every device */
   for (int i = 0; i < G; i++)
      if (m->traj_cap[i] < row)
      {
         cudaSetDevice(m->devices[i]);
         if (m->d_traj[i]) cudaFree(m->d_traj[i]);
         m->d_traj[i] = nullptr;
         if (cudaMalloc((void **) &m->d_traj[i], row * sizeof(double)) != cudaSuccess) return mfail(OCB_ERR_ALLOC, "trajectory buffer");
         m->traj_cap[i] = row;
      }
   if (!m->comms.empty())
   {
      /* device-side exchange: every device contributes its (cost, global id), learns all of them and
       * receives the winner's trajectory from its owner; the root is known to the host already (it has
       * the same pairs), so both collectives go into one group per device thread */
      rc = for_each_engine(G, [&](int i)
      {
         cudaSetDevice(m->devices[i]);
         const double pair[2] = {local_idx[i] >= 0 ? local_cost[i] : HUGE_VAL,
                                 local_idx[i] >= 0 ? (double) b->global[i][local_idx[i]] : -1.0};
         cudaStream_t st = m->streams[i];
         if (cudaMemcpyAsync(m->d_pair[i], pair, sizeof(pair), cudaMemcpyHostToDevice, st) != cudaSuccess) return (int) OCB_ERR_CUDA;
         if (i == win && ocb_batch_copy_run_traj_device(b->batches[i], local_idx[i], m->d_traj[i]) != OCB_OK) return (int) OCB_ERR_CUDA;
         g_nccl.GroupStart();
         int e1 = g_nccl.AllGather(m->d_pair[i], m->d_pairs[i], 2, NCCL_FLOAT64, m->comms[i], st);
         int e2 = (win >= 0) ? g_nccl.Broadcast(m->d_traj[i], m->d_traj[i], row, NCCL_FLOAT64, win, m->comms[i], st) : 0;
         g_nccl.GroupEnd();
         if (e1 || e2) return (int) OCB_ERR_CUDA;
         return cudaStreamSynchronize(st) == cudaSuccess ? (int) OCB_OK : (int) OCB_ERR_CUDA;
      });
      if (rc) return mfail(rc, "NCCL best-cost gather failed");
      if (traj && win >= 0)
      {
         /* read it back from a device that did NOT own it: proves the broadcast */
         const int src = (G > 1) ? (win + 1) % G : win;
         cudaSetDevice(m->devices[src]);
         if (cudaMemcpy(traj, m->d_traj[src], row * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return mfail(OCB_ERR_CUDA, "trajectory read-back");
      }
   }
   else if (win >= 0)
   {
      cudaSetDevice(m->devices[win]);
      if (ocb_batch_copy_run_traj_device(b->batches[win], local_idx[win], m->d_traj[win]) != OCB_OK) return mfail(OCB_ERR_CUDA, ocb_last_error());
      ocb_engine_sync(m->engines[win]);
      for (int i = 0; i < G; i++)
      {
         if (i == win) continue;
         cudaError_t ce = (m->devices[i] == m->devices[win])
                             ? cudaMemcpy(m->d_traj[i], m->d_traj[win], row * sizeof(double), cudaMemcpyDeviceToDevice)
                             : cudaMemcpyPeer(m->d_traj[i], m->devices[i], m->d_traj[win], m->devices[win], row * sizeof(double));
         if (ce != cudaSuccess) return mfail(OCB_ERR_CUDA, cudaGetErrorString(ce));
      }
      if (traj)
      {
         const int src = (G > 1) ? (win + 1) % G : win;
         cudaSetDevice(m->devices[src]);
         if (cudaMemcpy(traj, m->d_traj[src], row * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return mfail(OCB_ERR_CUDA, "trajectory read-back");
      }
   }
   return OCB_OK;
}

/* device pointer (on device slot `slot`) of the winner's trajectory after ocb_multi_batch_best */
extern "C" int ocb_multi_best_traj_device(ocb_multi *m, int slot, void **d_traj)
{
   if (!m || !d_traj || slot < 0 || slot >= m->G) return mfail(OCB_ERR_ARG, "bad argument");
   *d_traj = m->d_traj[slot];
   return OCB_OK;
}
