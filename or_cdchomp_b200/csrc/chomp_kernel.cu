/* chomp_kernel.cu -- the batched CHOMP iteration for sm_100a.
 *
 * One thread block owns one run for a whole `iterate` call: the trajectory, the
 * gradient and the momentum stay in shared memory across all n_iter iterations,
 * so HBM sees the trajectory once on the way in and once on the way out; the
 * signed distance fields are gathered through L1/L2 (they are shared by every
 * run on the GPU).  One thread owns one waypoint.
 *
 * What it replaces in the reference (paths relative to the reference root):
 *   cd_chomp_iterate                     src/libcd/chomp.c:430-683
 *   sphere_cost_pre (FK, Jacobians,      src/orcdchomp_mod.cpp:968-1132
 *     finite-difference vel / acc)
 *   sphere_cost (SDF cost + gradient,    src/orcdchomp_mod.cpp:1134-1327
 *     self collision, J^T accumulation)
 *   cd_grid_lookup_index / interp / grad src/libcd/grid.c:191-209, 386-454, 331-384
 *   cd_kin_pose_compos / compose_vec     src/libcd/kin.c:180-212, 244-271
 *   HMC momentum resampling              src/orcdchomp_mod.cpp:2755-2768
 *   the iterate loop + final cost pass   src/orcdchomp_mod.cpp:2752-2831
 *
 * Differences of FORM (not of result):
 *   - no Jacobian is materialised: each sphere's workspace force is folded into a
 *     per-joint-frame wrench (F, M) and J^T f is evaluated as axis . (M - o x F)
 *     for every ancestor joint;
 *   - A^-1 G is a banded LDL^T solve (factor computed once on the host), not a
 *     product with an explicit dense inverse; for the default metric c tridiag(-1, 2, -1)
 *     it IS the product with the explicit inverse, whose entries are known in closed
 *     form, evaluated as two weighted running sums (band_solve_121_scan);
 *   - A T + B is the banded stencil;  B and trC are evaluated from the end points; with
 *     the default metric and fixed end points A^-1 (A T + B) = T - (straight line between
 *     the end points) and no stencil is evaluated in the update (line_form);
 *   - hard constraints (chomp.c:553-600) are a block-tridiagonal sweep over the waypoints
 *     instead of a dense system when the metric is tridiagonal (chomp_constraints.cuh);
 *   - each SDF sample reads its 4 cells once and yields value and gradient.
 */
#include "chomp_device.cuh"

/* Sizes of the compiled robot and the mode flags: kernel arguments in the library's own
 * instantiations, literal constants when the kernel is compiled at run time for one batch
 * (ocb_jit.cpp passes -DOCB_JIT -DOCB_JIT_nsa=15 ...): loop bounds, table strides and the
 * workspace carve-up then fold into immediates. */
#ifdef OCB_JIT
#define DIM(a, f) (OCB_JIT_##f)
#else
#define DIM(a, f) ((a).f)
#endif

/* development aid (run-time compilation with OCB_JIT_FLAGS=-DOCB_PHASE_CLOCKS): per-warp cycle totals of
 * the phases of an iteration, written over the cost trace of the run (scripts/dev_phase_clocks.py) */
#ifdef OCB_PHASE_CLOCKS
#define PHASE_DECL long long ph_acc[16] = {0}; long long ph_t = clock64()
#define PHASE(k) do { const long long ph_now = clock64(); ph_acc[k] += ph_now - ph_t; ph_t = ph_now; } while (0)
#define PHASE_ARG , long long *ph_acc, long long &ph_t
#define PHASE_PASS , ph_acc, ph_t
#else
#define PHASE_DECL
#define PHASE(k)
#define PHASE_ARG
#define PHASE_PASS
#endif

#ifdef OCB_JIT_ROBOT
#include "chomp_jit_robot.cuh" /* this batch's robot as straight-line code (run-time compilation only) */
#endif
#ifndef OCB_JIT
#include "chomp_constraints.cuh" /* hard constraints: library kernel only */
#endif

namespace
{

/* ------------------------------------------------------------------------- */
/* Shared-memory carve-up, computed identically on host and device. */
struct SmemLayout
{
   int T, G, AG, red, ws, cut2, radius, metric; /* offsets in doubles */
   int sdf, sph, desc, mt, ired;        /* offsets in bytes   */
   int bytes;
};

__host__ __device__ inline SmemLayout smem_layout(const OcbChompArgs &a, const int Pp, const int n)
{
   SmemLayout l;
   int d = 0;
   l.T = d; d += n * Pp;
   l.G = d; d += n * Pp;
   l.AG = d; d += DIM(a, use_momentum) ? n * Pp : 0;
   l.red = d; d += 36;
   l.ws = d; d += (int) a.ws_stride;
   /* a kernel that has the robot as code (robot_smem) needs none of its tables here and keeps the
    * tridiagonal factor (L, 1/d: the solve's operands in every iteration) in their place */
   const bool robot = a.robot_smem != 0;
   l.cut2 = d; d += robot ? 0 : DIM(a, nsa) * (DIM(a, NAp) + DIM(a, nsi));
   l.radius = d; d += robot ? 0 : DIM(a, nsa) + DIM(a, nsi);
   l.metric = d; d += (robot && a.bw == 1) ? 2 * (Pp - 2) : 0;
   int b = d * 8;
   l.sdf = b; b += DIM(a, nsdf) * (int) sizeof(OcbSdfDev);
   l.sph = b; b += robot ? 0 : DIM(a, nsa) * (int) sizeof(OcbSphereDev);
   l.desc = b; b += robot ? 0 : DIM(a, n_desc) * 4;
   l.mt = b; b += DIM(a, use_hmc) ? (626 + 626 + 16) * 4 : 0; /* state, saved copy (serial fallback), scratch */
   l.ired = b; b += 40 * 4;
   l.bytes = b;
   return l;
}

/* read-only tables, staged in shared memory */
struct Tables
{
   const OcbSphereDev *sph;
   const int *desc;
   const double *cut2;
   const double *radius;
   const OcbSdfDev *sdfs;
};

/* forward kinematics of waypoint t: world positions of the active spheres.
 * Replaces SetActiveDOFValues + GetTransform()*pos (mod.cpp:1026-1038). */
template <bool FLOAT, int PP>
__device__ __forceinline__ void fk_waypoint(const OcbChompArgs &a, const Tables &tb,
                                            const double *__restrict__ Ts, double *__restrict__ ws, int t)
{
   const int Pp = PP ? PP : a.Ppad;
   double *slots = ws + 3 * DIM(a, nsa) * Pp;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   for (int j = 0; j < DIM(a, nj); j++)
   {
      const OcbJointDev &J = a.joints[j];
      fk_step<true, FLOAT>(J, Ts[J.dof * Pp + t], slots, Pp, t, R, tr, ax, org, Ts + t, Pp);
      for (int s = J.sph_begin; s < J.sph_end; s++)
      {
         const double px = tb.sph[s].pos[0], py = tb.sph[s].pos[1], pz = tb.sph[s].pos[2];
         double *o = ws + 3 * s * Pp + t;
         o[0] = R[0] * px + R[1] * py + R[2] * pz + tr[0];
         o[Pp] = R[3] * px + R[4] * py + R[5] * pz + tr[1];
         o[2 * Pp] = R[6] * px + R[7] * py + R[8] * pz + tr[2];
      }
   }
}

/* J^T f for waypoint t.  The workspace forces of sphere_cost have been gathered as
 * one wrench (F, M about the world origin) per sphere-carrying joint frame; a second
 * forward sweep regenerates each joint's axis and origin and contracts them with
 * the wrench of the joint's subtree:  dC/dq_j = c0 * axis . (M - origin x F)  for a
 * revolute joint, c0 * axis . F for a prismatic one.  This is the product with the
 * CalculateJacobian columns (mod.cpp:1048, 1244, 1314) without storing them. */
template <bool FLOAT, int PP>
__device__ __forceinline__ void flush_wrenches(const OcbChompArgs &a, const Tables &tb,
                                               const double *__restrict__ Ts, double *__restrict__ ws,
                                               double *__restrict__ Gs, int t)
{
   const int Pp = PP ? PP : a.Ppad;
   double *slots = ws + 3 * DIM(a, nsa) * Pp;
   const double *Wg = ws + (3 * DIM(a, nsa) + 12 * DIM(a, n_slots)) * Pp + t;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   for (int j = 0; j < DIM(a, nj); j++)
   {
      const OcbJointDev &J = a.joints[j];
      fk_step<false, FLOAT>(J, Ts[J.dof * Pp + t], slots, Pp, t, R, tr, ax, org, Ts + t, Pp);
      double F0 = 0.0, F1 = 0.0, F2 = 0.0, M0 = 0.0, M1 = 0.0, M2 = 0.0;
      for (int di = J.desc_begin; di < J.desc_end; di++)
      {
         const double *Wo = Wg + 6 * tb.desc[di] * Pp;
         F0 += Wo[0]; F1 += Wo[Pp]; F2 += Wo[2 * Pp];
         M0 += Wo[3 * Pp]; M1 += Wo[4 * Pp]; M2 += Wo[5 * Pp];
      }
      double val;
      if (J.type == OCB_JOINT_REVOLUTE)
      {
         const double mx = M0 - (org[1] * F2 - org[2] * F1);
         const double my = M1 - (org[2] * F0 - org[0] * F2);
         const double mz = M2 - (org[0] * F1 - org[1] * F0);
         val = ax[0] * mx + ax[1] * my + ax[2] * mz;
      }
      else
         val = ax[0] * F0 + ax[1] * F1 + ax[2] * F2;
      Gs[J.dof * Pp + t] = fma(J.c0, val, Gs[J.dof * Pp + t]);
   }
   if (FLOAT)
   {
      /* the base pose sees the total wrench of the waypoint */
      double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
      for (int g = 0; g < DIM(a, ng); g++)
      {
         const double *Wo = Wg + 6 * g * Pp;
         F[0] += Wo[0]; F[1] += Wo[Pp]; F[2] += Wo[2 * Pp];
         M[0] += Wo[3 * Pp]; M[1] += Wo[4 * Pp]; M[2] += Wo[5 * Pp];
      }
      pose_gradient(Ts + t, Pp, F, M, Gs + t, Pp);
   }
}

/* ------------------------------------------------------------------------- */
/* cost (and, when want_grad, the configuration-space gradient row) of moving
 * waypoint t (1..P-2).  Restates sphere_cost (mod.cpp:1134-1327) on top of the
 * finite differences of sphere_cost_pre (mod.cpp:1099-1127).
 *
 * Self collision: the reference visits every ordered pair (s, s2) and adds
 * (J_s - J_s2)^T x(s, s2).  Here each unordered pair is visited once; both
 * directed terms x(s,o) and x(o,s) are formed, their difference is the net
 * workspace force on s and its negative the force on o.  Forces are gathered as
 * wrenches per joint frame in ws and mapped to joint space by flush_wrenches. */
template <bool FLOAT, int PP>
__device__ __forceinline__ double waypoint_cost(const OcbChompArgs &a, const Tables &tb,
                                                const double *__restrict__ Ts,
                                                double *__restrict__ ws, double *__restrict__ Gs,
                                                int t, bool want_grad PHASE_ARG)
{
   const int Pp = PP ? PP : a.Ppad;
   const int nsa = DIM(a, nsa);
   const int row = DIM(a, NAp) + DIM(a, nsi);
   double *Wg = ws + (3 * nsa + 12 * DIM(a, n_slots)) * Pp + t;
   const double inv2dt = 1.0 / (2.0 * a.dt);
   const double invdt2 = 1.0 / (a.dt * a.dt);
   const double es = a.eps_self, inv_es = 1.0 / es, half_inv_es = 0.5 / es;
   double cost = 0.0;
   /* start_tsr: the first moving point is the start itself; its velocity is the one-sided difference and
    * its acceleration that of the next point (mod.cpp:1108-1114, 1126-1128) */
   const bool first_free = a.free_start && t == 1;
   const double invdt = 1.0 / a.dt;

   if (want_grad)
      for (int k = 0; k < 6 * DIM(a, ng); k++) Wg[k * Pp] = 0.0;

   PHASE(2);

   for (int j = 0; j < DIM(a, nj); j++)
   {
      const OcbJointDev &J = a.joints[j];
      if (J.sph_begin == J.sph_end) continue;
      double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
      for (int s = J.sph_begin; s < J.sph_end; s++)
      {
         const double *ps = ws + 3 * s * Pp + t;
         const double p[3] = {ps[0], ps[Pp], ps[2 * Pp]};
         const double radius = tb.radius[s];

         /* --- obstacle term, first step: which field, how far (mod.cpp:1169-1198) --- */
         double d_obs, bg[3];
         const int best = obstacle_probe(a, tb.sdfs, p, radius, DIM(a, nsdf), d_obs, bg);
         const bool obs = (best >= 0) && (d_obs < a.eps);
         double vel[3];
#pragma unroll
         for (int k = 0; k < 3; k++)
            vel[k] = first_free ? (ps[k * Pp + 1] - ps[k * Pp]) * invdt : (ps[k * Pp + 1] - ps[k * Pp - 1]) * inv2dt;
         const double vn2 = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2];
         double vn, iv2; /* |v| and 1 / |v|^2, unguarded (inf at rest), as mod.cpp:1239 */
         speed_terms(vn2, vn, iv2);
         const bool moving = vn > 0.000001;
         double cost_s = 0.0;
         double f[3] = {0.0, 0.0, 0.0};

         /* --- obstacle term, second step (mod.cpp:1200-1249) --- */
         if (obs)
         {
            double acc[3] = {0.0, 0.0, 0.0};
            if (want_grad)
            {
#pragma unroll
               for (int k = 0; k < 3; k++)
                  acc[k] = first_free ? (ps[k * Pp + 1] * -2.0 + p[k] + ps[k * Pp + 2]) * invdt2
                                      : (p[k] * -2.0 + ps[k * Pp - 1] + ps[k * Pp + 1]) * invdt2;
            }
            obstacle_apply(a, tb.sdfs[best], d_obs, bg, vel, acc, vn, iv2, moving, want_grad, cost_s, f);
         }

         /* --- self collision, each unordered pair once (1251-1317) --- */
         const double ws_self = vn * a.obs_factor_self;
         /* a partner within range: q its centre, po its column in ws (NULL: inactive, frozen) */
         auto in_range = [&](const double q[3], const double *po, int o)
         {
            const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            const double inv = fast_rsqrt(d2); /* spheres on different links within range of each other: no denormal */
            const double dist = d2 * inv;
            const double dd = dist - (radius + tb.radius[o]);
            /* cost shape shared by both directions (1281-1289) */
            const double cshape = (dd < 0.0) ? (0.5 * es - dd) : half_inv_es * (dd - es) * (dd - es);
            cost_s += ws_self * cshape;
            double w2 = 0.0, v2[3] = {0.0, 0.0, 0.0}, r2 = 0.0; /* r2: 1 / |v2|^2 */
            bool moving2 = false;
            if (po)
            {
#pragma unroll
               for (int r = 0; r < 3; r++)
                  v2[r] = first_free ? (po[r * Pp + 1] - po[r * Pp]) * invdt : (po[r * Pp + 1] - po[r * Pp - 1]) * inv2dt;
               const double v2n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
               double v2n, iv22;
               speed_terms(v2n2, v2n, iv22);
               r2 = iv22;
               moving2 = v2n > 0.000001;
               w2 = v2n * a.obs_factor_self;
               cost_s += w2 * cshape; /* the other sphere's own cost_sphere term */
            }
            if (!want_grad) return;
            const double sc = (dd < 0.0) ? -1.0 : ((dd < es) ? (dd * inv_es - 1.0) : 1.0);
            const double gh[3] = {dx * inv, dy * inv, dz * inv};
            double x[3];
            const double wa = sc * ws_self;
#pragma unroll
            for (int r = 0; r < 3; r++) x[r] = gh[r] * wa;
            if (moving)
            {
               const double pj = (x[0] * vel[0] + x[1] * vel[1] + x[2] * vel[2]) * iv2;
#pragma unroll
               for (int r = 0; r < 3; r++) x[r] = fma(-pj, vel[r], x[r]);
            }
            if (po)
            {
               /* the pair seen from o: unit vector -gh, weighted by o's speed */
               double y[3];
               const double wb = -sc * w2;
#pragma unroll
               for (int r = 0; r < 3; r++) y[r] = gh[r] * wb;
               if (moving2)
               {
                  const double pj = (y[0] * v2[0] + y[1] * v2[1] + y[2] * v2[2]) * r2;
#pragma unroll
                  for (int r = 0; r < 3; r++) y[r] = fma(-pj, v2[r], y[r]);
               }
               /* (J_s - J_o)^T x + (J_o - J_s)^T y = J_s^T (x - y) - J_o^T (x - y) */
#pragma unroll
               for (int r = 0; r < 3; r++) x[r] -= y[r];
               double *Wo = Wg + 6 * tb.sph[o].group * Pp;
               Wo[0] -= x[0];
               Wo[Pp] -= x[1];
               Wo[2 * Pp] -= x[2];
               Wo[3 * Pp] -= q[1] * x[2] - q[2] * x[1];
               Wo[4 * Pp] -= q[2] * x[0] - q[0] * x[2];
               Wo[5 * Pp] -= q[0] * x[1] - q[1] * x[0];
            }
#pragma unroll
            for (int r = 0; r < 3; r++) f[r] += x[r];
         };
         const double *crow = tb.cut2 + s * row;
         /* active partners: range tests four at a time (independent chains); the padded
          * tail of the cut2 row is -1 and never passes */
         for (int o0 = s + 1; o0 < nsa; o0 += 4)
         {
            unsigned mask = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
               const double *po = ws + 3 * min(o0 + k, nsa - 1) * Pp + t;
               const double dx = p[0] - po[0], dy = p[1] - po[Pp], dz = p[2] - po[2 * Pp];
               const double d2 = dx * dx + dy * dy + dz * dz;
               if (d2 <= crow[o0 + k]) mask |= 1u << k;
            }
            if (mask == 0) continue;
#pragma unroll 1
            for (int k = 0; k < 4; k++)
            {
               if (!((mask >> k) & 1u)) continue;
               const double *po = ws + 3 * (o0 + k) * Pp + t;
               const double q[3] = {po[0], po[Pp], po[2 * Pp]};
               in_range(q, po, o0 + k);
            }
         }
         /* inactive partners are frozen in the world (mod.cpp:2332-2345) */
         for (int i = 0; i < DIM(a, nsi); i++)
         {
            const double q[3] = {__ldg(a.inactive_pos + 3 * i), __ldg(a.inactive_pos + 3 * i + 1),
                                 __ldg(a.inactive_pos + 3 * i + 2)};
            const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            if (dx * dx + dy * dy + dz * dz <= crow[DIM(a, NAp) + i]) in_range(q, nullptr, nsa + i);
         }
         cost += cost_s;
         if (want_grad)
         {
            F[0] += f[0]; F[1] += f[1]; F[2] += f[2];
            M[0] += p[1] * f[2] - p[2] * f[1];
            M[1] += p[2] * f[0] - p[0] * f[2];
            M[2] += p[0] * f[1] - p[1] * f[0];
         }
      }
      if (want_grad)
      {
         double *Wo = Wg + 6 * tb.sph[J.sph_begin].group * Pp;
         Wo[0] += F[0]; Wo[Pp] += F[1]; Wo[2 * Pp] += F[2];
         Wo[3 * Pp] += M[0]; Wo[4 * Pp] += M[1]; Wo[5 * Pp] += M[2];
      }
   }
   PHASE(3);
   if (want_grad) flush_wrenches<FLOAT, PP>(a, tb, Ts, ws, Gs, t);
   PHASE(4);
   return cost;
}

template <bool FLOAT, int PP, int NN, bool CONS = false>
__device__ __forceinline__ void chomp_iterate_body(const OcbChompArgs &a)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int tid = threadIdx.x;
   const int NT = blockDim.x;
   const int run = blockIdx.x;
   const int Pp = PP ? PP : a.Ppad;
   /* start_tsr (free_start): the start point is optimised too.  The kernel then works on P + 1 columns --
    * column 0 is a stand-in that nothing reads with a non-zero weight, columns 1..P are the run's P
    * waypoints, of which 1..P-1 move (m = P - 1, mod.cpp:2316) */
   const int fs = CONS ? a.free_start : 0;
   const int P = PP ? PP : a.P + fs, m = PP ? PP - 2 : a.m, n = NN ? NN : a.n;

   /* ---- shared memory carve-up ---- */
   const SmemLayout lay = smem_layout(a, Pp, n);
   double *sd = reinterpret_cast<double *>(smem_raw);
   double *Ts = sd + lay.T;     /* [n][Pp] */
   double *Gs = sd + lay.G;     /* [n][Pp] */
   double *AGs = sd + lay.AG;   /* [n][Pp] (momentum only) */
   double *red = sd + lay.red;
   OcbSdfDev *sdfs = reinterpret_cast<OcbSdfDev *>(smem_raw + lay.sdf);
   uint32_t *mts = reinterpret_cast<uint32_t *>(smem_raw + lay.mt);
   int *ired = reinterpret_cast<int *>(smem_raw + lay.ired);
   double *ws = sd + lay.ws;
   Tables tb;
   tb.sdfs = sdfs;
   const double *Ls = a.Lband, *dinv = a.dinv; /* the solve's factor: global, or staged below */
#ifdef OCB_JIT_ROBOT
   tb.sph = nullptr; tb.desc = nullptr; tb.cut2 = nullptr; tb.radius = nullptr;
   if (a.bw == 1)
   {
      double *ml = sd + lay.metric, *md = ml + m;
      for (int e = tid; e < m; e += NT) { ml[e] = __ldg(a.Lband + e); md[e] = __ldg(a.dinv + e); }
      Ls = ml;
      dinv = md;
   }
#else
   {
      double *c2 = sd + lay.cut2, *rad = sd + lay.radius;
      OcbSphereDev *sph = reinterpret_cast<OcbSphereDev *>(smem_raw + lay.sph);
      int *dsc = reinterpret_cast<int *>(smem_raw + lay.desc);
      for (int e = tid; e < DIM(a, nsa) * (DIM(a, NAp) + DIM(a, nsi)); e += NT) c2[e] = __ldg(a.cut2 + e);
      for (int e = tid; e < DIM(a, nsa) + DIM(a, nsi); e += NT) rad[e] = __ldg(a.radius + e);
      {
         const int words = DIM(a, nsa) * (int) (sizeof(OcbSphereDev) / 4);
         const uint32_t *src = reinterpret_cast<const uint32_t *>(a.spheres);
         uint32_t *dst = reinterpret_cast<uint32_t *>(sph);
         for (int e = tid; e < words; e += NT) dst[e] = __ldg(src + e);
      }
      for (int e = tid; e < DIM(a, n_desc); e += NT) dsc[e] = __ldg(a.desc + e);
      tb.sph = sph; tb.desc = dsc; tb.cut2 = c2; tb.radius = rad;
   }
#endif

   /* ---- stage per-run state and shared constants ---- */
   double *traj = a.traj + (size_t) run * (P - fs) * n;
   for (int e = tid; e < (P - fs) * n; e += NT) Ts[(e % n) * Pp + (e / n) + fs] = traj[e];
   if (fs)
      for (int j = tid; j < n; j += NT) Ts[j * Pp] = traj[j];
   if (DIM(a, use_momentum))
   {
      const double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) AGs[(e % n) * Pp + (e / n) + 1] = ag[e];
   }
   {
      const int words = (int) (sizeof(OcbSdfDev) / 4) * DIM(a, nsdf);
      const uint32_t *src = reinterpret_cast<const uint32_t *>(a.sdfs);
      uint32_t *dst = reinterpret_cast<uint32_t *>(sdfs);
      for (int e = tid; e < words; e += NT) dst[e] = src[e];
   }
   if (DIM(a, use_hmc))
      for (int e = tid; e < 625; e += NT) mts[e] = a.mt_state[(size_t) run * 625 + e];
   int leapfrog_first = DIM(a, use_momentum) ? a.leapfrog_first[run] : 0;
   int hmc_next = DIM(a, use_hmc) ? a.hmc_next[run] : -1;
   int status = 0;
   const double inv_m = 1.0 / m;
   const double inv_lambda = 1.0 / a.lambda;

   /* end points define B and trC (chomp.c:278-296, 318-331) */
   double trC = 0.0;
   {
      double ss = 0.0, sg = 0.0, gg = 0.0;
      for (int j = 0; j < n; j++)
      {
         const double qs = traj[j], qg = traj[(size_t) (P - fs - 1) * n + j];
         ss += qs * qs; sg += qs * qg; gg += qg * qg;
      }
      trC = 0.5 * (a.trc_ss * ss + 2.0 * a.trc_sg * sg + a.trc_gg * gg);
   }
   __syncthreads();

   /* Default metric with both end points fixed (band_121 & 2): A = c tridiag(-1, 2, -1) and B = -c (q_start e_1 +
    * q_goal e_m), so A^-1 (A T + B) = T - L with L the straight line from q_start to q_goal -- exactly, no
    * stencil and no rounding of a product with A that the solve then undoes.  The update becomes
    * T -= (A^-1 G_obs / m + T - L) / lambda.  The full gradient G (grad_mode 1) is only formed when asked for. */
   const bool line_form = !CONS && (a.band_121 & 2) && a.grad_mode != 1 && band_scan_lanes(a, n) >= 2;
   const double inv_mp1 = 1.0 / (m + 1);
   double cost_obs = 0.0, cost_smooth = 0.0;
   double csum_done = 0.0, ssum_done = 0.0; /* this thread's cost partials of the last completed iteration */
   int red_parity = 0;
   int iters_done = 0; /* iterations completed (the reference's r->iter when iterate returns or throws) */
   int limit_rounds = 0; /* most joint-limit projection steps any iteration of this call needed */
   PHASE_DECL;
   for (int iter = 0; iter <= a.n_iter; iter++)
   {
      const bool final_pass = (iter == a.n_iter);
      PHASE(15);

      /* ---- HMC momentum resample (mod.cpp:2755-2768) ---- */
      if (DIM(a, use_hmc) && !final_pass && a.iter_base + iter == hmc_next)
      {
         const double alpha = 100.0 * exp(0.02 * (a.iter_base + iter)); /* r->iter of the whole command */
         const double sigma = 1.0 / sqrt(alpha);
         uint32_t *saved = mts + 626;
         int *scratch = reinterpret_cast<int *>(mts + 1252);
         for (int e = tid; e < 625; e += NT) saved[e] = mts[e];
         __syncthreads();
         double u = 0.0;
         if (DIM(a, use_hmc) == 2 || !hmc_resample_parallel(mts, scratch, AGs, Pp, m, n, sigma, &u))
         {
            /* a zero word was drawn (or use_hmc == 2, the test hook that forces this path): redo this
             * resample serially, with the exact gsl_rng_uniform_pos semantics */
            __syncthreads();
            for (int e = tid; e < 625; e += NT) mts[e] = saved[e];
            __syncthreads();
            if (tid == 0)
            {
               for (int i = 0; i < m; i++)
                  for (int j = 0; j < n; j++) AGs[j * Pp + i + 1] = mt_gaussian(mts, sigma);
               red[35] = mt_uniform(mts);
            }
            __syncthreads();
            u = red[35];
         }
         hmc_next = hmc_next + 1 + (int) (-log(u) / a.hmc_lambda);
         leapfrog_first = 1;
         __syncthreads();
      }

      /* ---- forward kinematics of all P waypoints ---- */
#ifdef OCB_JIT_ROBOT
      for (int t = tid; t < P; t += NT) jr_fk_waypoint<FLOAT>(Ts, ws, Pp, t, a.trig_cache + (size_t) run * 2 * DIM(a, nj) * Pp);
#else
      for (int t = tid; t < P; t += NT) fk_waypoint<FLOAT, PP>(a, tb, Ts, ws, t);
#endif
      PHASE(0);
      __syncthreads();
      PHASE(1);

      /* ---- obstacle + self-collision cost / gradient, then G = G/m + A T + B ---- */
      double csum = 0.0, ssum = 0.0;
      for (int t = tid + 1; t <= m; t += NT)
      {
         if (!final_pass)
            for (int j = 0; j < n; j++) Gs[j * Pp + t] = 0.0;
#ifdef OCB_JIT_ROBOT
         /* few fields: their descriptors are read straight from the kernel parameters */
         const OcbSdfDev *fields = (OCB_JIT_nsdf <= OCB_INLINE_SDFS) ? a.sdf_inline : tb.sdfs;
         const double *trig = a.trig_cache + (size_t) run * 2 * DIM(a, nj) * Pp;
         csum += final_pass ? jr_waypoint_cost<FLOAT, false>(a, fields, Ts, ws, Gs, Pp, t, trig PHASE_PASS)
                            : jr_waypoint_cost<FLOAT, true>(a, fields, Ts, ws, Gs, Pp, t, trig PHASE_PASS);
#else
         csum += waypoint_cost<FLOAT, PP>(a, tb, Ts, ws, Gs, t, !final_pass PHASE_PASS);
#endif
         if (!final_pass && line_form)
         {
            /* the smoothness part of the update is T - L in closed form (see line_form): G stays the raw
             * obstacle gradient, its 1/m goes into the solve's scale */
            if (a.grad_mode == 2)
               for (int j = 0; j < n; j++) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = Gs[j * Pp + t] * inv_m;
         }
         else if (!final_pass)
         {
            const double bi = __ldg(a.bcoef_i + t - 1), bf = __ldg(a.bcoef_f + t - 1);
            for (int j = 0; j < n; j++)
            {
               const double *Tj = Ts + j * Pp;
               double g = Gs[j * Pp + t] * inv_m;
               if (a.grad_mode == 2) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
               g += band_AT(a, Tj, t, m) + (bi * Tj[0] + bf * Tj[P - 1]);
               Gs[j * Pp + t] = g;
               if (a.grad_mode == 1) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
            }
         }
         else
            ssum += smooth_row(a, Ts, t, Pp, P, n);
      }
      if (final_pass)
      {
         block_sum2(csum, ssum, red, red_parity);
         cost_obs = csum * inv_m;
         cost_smooth = ssum + trC;
         break;
      }
      PHASE(5);
      __syncthreads(); /* every row of G is complete */
      PHASE(6);

      /* ---- AG = A^-1 G, then the momentum / plain update T -= AG/lambda (chomp.c:525-548, 604-605) with
       * the joint-limit check of every entry written.  Tridiagonal metric: a block-wide scan solves all
       * dofs at once and the lane that owns an entry updates it on the spot -- one barrier for solve,
       * update and the any-violation vote.  Wider metrics: one thread per dof, then thread per waypoint. ---- */
      int violated = 0;
#ifndef OCB_JIT
      if (CONS)
         /* ---- the same with hard constraints (chomp.c:553-600): chomp_constraints.cuh ---- */
         violated = con_update<FLOAT, false>(a, run, Ts, Gs, AGs, ws, ws + 3 * DIM(a, nsa) * Pp, red, ired, Pp, m, n,
                                             inv_lambda, leapfrog_first);
      else
#endif
      {
         const double coef = (leapfrog_first ? 0.5 : 1.0) * inv_lambda;
         auto update = [&](const int j, const int t, double step)
         {
            if (line_form) /* + A^-1 (A T + B) = T - L, L the straight line between the fixed end points */
               step += Ts[j * Pp + t] - fma(Ts[j * Pp + P - 1] - Ts[j * Pp], (double) t * inv_mp1, Ts[j * Pp]);
            if (DIM(a, use_momentum))
            {
               step = fma(coef, step, AGs[j * Pp + t]);
               AGs[j * Pp + t] = step;
            }
            const double q = fma(-inv_lambda, step, Ts[j * Pp + t]);
            Ts[j * Pp + t] = q;
            violated |= (q < __ldg(a.lim_lo + j)) | (q > __ldg(a.lim_hi + j));
         };
         const int lpd = band_scan_lanes(a, n);
         if (line_form)
            band_solve_121_scan(a, Gs, Pp, m, n, lpd, [&](const int j, const int i, const double ag) { update(j, i + 1, ag); },
                                a.band_121_scale * inv_m);
         else if (lpd >= 2)
            band_solve_scan(a, Gs, Pp, m, n, lpd, [&](const int j, const int i, const double ag) { update(j, i + 1, ag); }, Ls, dinv);
         else
         {
            if (tid < n) band_solve(a, Gs + tid * Pp + 1, m);
            PHASE(7);
            __syncthreads();
            PHASE(8);
            for (int t = tid + 1; t <= m; t += NT)
               for (int j = 0; j < n; j++) update(j, t, Gs[j * Pp + t]);
         }
         if (DIM(a, use_momentum)) leapfrog_first = 0;
      }
      PHASE(9);
      const int any_violation = __syncthreads_or(violated);
      PHASE(10);

      /* ---- joint-limit projection (chomp.c:608-655) ---- */
      int rounds = 0;
      const bool limits_ok = !any_violation || project_joint_limits(a, Ts, Gs, red, ired, Pp, m, n, rounds);
      limit_rounds = max(limit_rounds, rounds);
      if (!limits_ok)
      {
         status = OCB_ERR_JLIMIT; /* chomp.c:651-655 returns -1 before the smoothness cost */
         /* the run keeps the costs of its last completed iteration */
         if (iters_done > 0 && !a.trace_on)
         {
            block_sum2(csum_done, ssum_done, red, red_parity);
            cost_obs = csum_done * inv_m;
            cost_smooth = ssum_done + trC;
         }
         break;
      }

      /* ---- smoothness cost of the updated trajectory (chomp.c:660-671) ---- */
      PHASE(11);
      ssum = 0.0;
      for (int t = tid + 1; t <= m; t += NT) ssum += smooth_row(a, Ts, t, Pp, P, n);
      PHASE(12);
      /* the costs of an iteration are read only by the trace and, when a later iteration leaves the
       * joint limits, as the run's last costs: the block-wide sums (and their barrier) are deferred
       * to those two places; each thread keeps the partial sums of the last completed iteration */
      csum_done = csum;
      ssum_done = ssum;
      if (a.trace_on)
      {
         block_sum2(csum, ssum, red, red_parity);
         cost_obs = csum * inv_m;
         cost_smooth = ssum + trC;
      }
      PHASE(13);
#ifndef OCB_PHASE_CLOCKS
      if (a.trace_on && tid == 0)
#else
      if (false)
#endif
      {
         double *tr = a.trace + ((size_t) run * a.n_iter + iter) * 3;
         tr[0] = cost_obs + cost_smooth;
         tr[1] = cost_obs;
         tr[2] = cost_smooth;
      }
      iters_done = iter + 1;
      if (FLOAT)
      {
         /* base quaternions back to unit length (mod.cpp:2805-2808); the end rows are unit already */
         __syncthreads();
         for (int t = tid + 1; t <= m; t += NT) pose_normalize(Ts + t, Pp);
         __syncthreads();
      }
   }

#ifdef OCB_PHASE_CLOCKS
   if (a.trace_on && (tid & 31) == 0)
      for (int k = 0; k < 16; k++) a.trace[(size_t) run * a.n_iter * 3 + (tid >> 5) * 16 + k] = (double) ph_acc[k];
#endif
   /* ---- write the run back ---- */
   __syncthreads();
   for (int e = tid; e < (P - fs) * n; e += NT) traj[e] = Ts[(e % n) * Pp + (e / n) + fs];
   if (DIM(a, use_momentum))
   {
      double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) ag[e] = AGs[(e % n) * Pp + (e / n) + 1];
   }
   if (DIM(a, use_hmc))
      for (int e = tid; e < 625; e += NT) a.mt_state[(size_t) run * 625 + e] = mts[e];
   if (tid == 0)
   {
      if (DIM(a, use_momentum)) a.leapfrog_first[run] = leapfrog_first;
      if (DIM(a, use_hmc)) a.hmc_next[run] = hmc_next;
      a.costs[(size_t) run * 3 + 0] = cost_obs + cost_smooth;
      a.costs[(size_t) run * 3 + 1] = cost_obs;
      a.costs[(size_t) run * 3 + 2] = cost_smooth;
      a.status[run] = status;
      a.iters_done[run] = iters_done;
      a.limit_rounds[run] = limit_rounds;
   }
}

#ifndef OCB_JIT
template <int NT_MAX, bool FLOAT, int PP, int NN, bool CONS = false>
__global__ void __launch_bounds__(NT_MAX, (NT_MAX == 128 && !FLOAT) ? 3 : 1)
chomp_iterate_kernel(const __grid_constant__ OcbChompArgs a)
{
   chomp_iterate_body<FLOAT, PP, NN, CONS>(a);
}
#endif

/* straight-line initial trajectory, evaluated exactly as mod.cpp:2456-2458:
 * traj[i] = q_s + ((q_g - q_s) * i) / (P-1), every operation rounded on its own */
__global__ void init_traj_kernel(double *traj, const double *q_start, const double *q_goal, int R, int P, int n)
{
   const size_t total = (size_t) R * P * n;
   for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < total;
        e += (size_t) gridDim.x * blockDim.x)
   {
      const int j = (int) (e % n);
      const int i = (int) ((e / n) % P);
      const size_t r = e / ((size_t) n * P);
      const double qs = q_start[r * n + j], qg = q_goal[r * n + j];
      const double num = __dmul_rn(__dsub_rn(qg, qs), (double) i);
      traj[e] = __dadd_rn(qs, __ddiv_rn(num, (double) (P - 1)));
   }
}

/* floating base: every row's base quaternion to unit length after the interpolation (mod.cpp:2461-2464) */
__global__ void normalize_rows_kernel(double *traj, size_t rows, int n)
{
   for (size_t r = blockIdx.x * (size_t) blockDim.x + threadIdx.x; r < rows; r += (size_t) gridDim.x * blockDim.x)
      pose_normalize(traj + r * n, 1);
}

/* arg-min of cost_total over the runs of this GPU (first wins ties; failed runs skipped) */
__global__ void best_kernel(const double *costs, const int *status, int R, int *best_run, double *best_cost)
{
   __shared__ double sv[32];
   __shared__ int si[32];
   double v = HUGE_VAL;
   int idx = 0x7fffffff;
   for (int r = threadIdx.x; r < R; r += blockDim.x)
   {
      const double c = costs[(size_t) r * 3];
      if (status[r] != 0 || !(c == c)) continue;
      if (c < v || (c == v && r < idx)) { v = c; idx = r; }
   }
   for (int o = 16; o > 0; o >>= 1)
   {
      const double ov = __shfl_xor_sync(FULL_MASK, v, o);
      const int oi = __shfl_xor_sync(FULL_MASK, idx, o);
      if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
   }
   if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
   __syncthreads();
   if (threadIdx.x == 0)
   {
      for (int w = 1; w < (int) ((blockDim.x + 31) >> 5); w++)
         if (sv[w] < v || (sv[w] == v && si[w] < idx)) { v = sv[w]; idx = si[w]; }
      *best_run = (idx == 0x7fffffff) ? -1 : idx;
      *best_cost = v;
   }
}

#ifndef OCB_JIT
/* parity hook: the kernels' own sdf_sample at arbitrary points of the grid frame; err = 1 where
 * cd_grid_lookup_index rejects the point, value = HUGE_VAL where interp returns it (the gradient
 * is computed regardless, as cd_grid_double_grad does) */
__global__ void sdf_sample_kernel(const OcbSdfDev S, const double *__restrict__ pts, int k, double *__restrict__ vals,
                                  double *__restrict__ grads, int *__restrict__ errs)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
   {
      const double g[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
      double v = 0.0, gg[3] = {0.0, 0.0, 0.0};
      const bool in = sdf_sample(S, g, v, gg);
      errs[i] = in ? 0 : 1;
      vals[i] = in ? v : 0.0;
      grads[3 * i] = in ? gg[0] : 0.0;
      grads[3 * i + 1] = in ? gg[1] : 0.0;
      grads[3 * i + 2] = in ? gg[2] : 0.0;
   }
}
#endif

} /* namespace */

#ifdef OCB_JIT
/* the one entry point of a run-time compiled instance */
extern "C" __global__ void __launch_bounds__(OCB_JIT_NT, OCB_JIT_MINBLOCKS)
chomp_iterate_jit(const __grid_constant__ OcbChompArgs a)
{
   chomp_iterate_body<OCB_JIT_FLOAT != 0, OCB_JIT_PP, OCB_JIT_NN>(a);
}
#else


extern "C" size_t ocb_chomp_smem_bytes(const OcbChompArgs *a)
{
   return (size_t) smem_layout(*a, a->Ppad, a->n).bytes;
}

template <int NT_MAX, bool FLOAT, int PP, int NN, bool CONS = false>
static cudaError_t launch_variant(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st)
{
   static OcbSmemOptIn optin; /* one per instantiation, keyed by device inside */
   cudaError_t e = optin.ensure(chomp_iterate_kernel<NT_MAX, FLOAT, PP, NN, CONS>, smem_bytes);
   if (e != cudaSuccess) return e;
   chomp_iterate_kernel<NT_MAX, FLOAT, PP, NN, CONS><<<args->R, threads, smem_bytes, st>>>(*args);
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_chomp(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st)
{
   if (threads > 256 || threads % 32) return cudaErrorInvalidValue;
   if (args->con_K > 0 || args->free_start) /* hard constraints: the generic instantiations only */
   {
      if (args->floating)
         return threads <= 128 ? launch_variant<128, true, 0, 0, true>(args, smem_bytes, threads, st)
                               : launch_variant<256, true, 0, 0, true>(args, smem_bytes, threads, st);
      return threads <= 128 ? launch_variant<128, false, 0, 0, true>(args, smem_bytes, threads, st)
                            : launch_variant<256, false, 0, 0, true>(args, smem_bytes, threads, st);
   }
   if (args->floating)
      return threads <= 128 ? launch_variant<128, true, 0, 0>(args, smem_bytes, threads, st)
                            : launch_variant<256, true, 0, 0>(args, smem_bytes, threads, st);
   /* the common trajectory shapes -- n_points 100 (BASELINE), 101 (the reference's default,
    * mod.cpp:1840) and 256, seven dofs -- with P and n as compile-time constants: every
    * [item][waypoint] shared-memory address becomes base + immediate and the dof loops unroll
    * (+15 % on config 2; one third of the generic kernel's instructions is address arithmetic).
    * Same source and arithmetic; variants agree up to the compiler's choice of fused multiply-adds. */
   if (args->Ppad == args->P && args->n == 7)
   {
      if (args->P == 100 && threads <= 128) return launch_variant<128, false, 100, 7>(args, smem_bytes, threads, st);
      if (args->P == 101 && threads <= 128) return launch_variant<128, false, 101, 7>(args, smem_bytes, threads, st);
      if (args->P == 256 && threads == 256) return launch_variant<256, false, 256, 7>(args, smem_bytes, threads, st);
   }
   return threads <= 128 ? launch_variant<128, false, 0, 0>(args, smem_bytes, threads, st)
                         : launch_variant<256, false, 0, 0>(args, smem_bytes, threads, st);
}

extern "C" cudaError_t ocb_launch_init_traj(double *traj, const double *q_start, const double *q_goal,
                                            int R, int P, int n, int floating, cudaStream_t st)
{
   const size_t total = (size_t) R * P * n;
   int blocks = (int) ((total + 255) / 256);
   if (blocks > 148 * 8) blocks = 148 * 8;
   if (blocks < 1) blocks = 1;
   init_traj_kernel<<<blocks, 256, 0, st>>>(traj, q_start, q_goal, R, P, n);
   if (floating)
   {
      const size_t rows = (size_t) R * P;
      int nb = (int) ((rows + 255) / 256);
      if (nb > 148 * 8) nb = 148 * 8;
      normalize_rows_kernel<<<nb, 256, 0, st>>>(traj, rows, n);
   }
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_sdf_sample(const OcbSdfDev *sdf, const double *d_points, int k, double *d_values,
                                             double *d_grads, int *d_errs, cudaStream_t st)
{
   int blocks = (k + 127) / 128;
   if (blocks > 148 * 8) blocks = 148 * 8;
   sdf_sample_kernel<<<blocks, 128, 0, st>>>(*sdf, d_points, k, d_values, d_grads, d_errs);
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_best(const double *costs, const int *status, int R, int *best_run,
                                       double *best_cost, cudaStream_t st)
{
   best_kernel<<<1, 1024, 0, st>>>(costs, status, R, best_run, best_cost);
   return cudaGetLastError();
}

#endif /* !OCB_JIT */
