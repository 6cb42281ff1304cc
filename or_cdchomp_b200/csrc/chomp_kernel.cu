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
 *     product with an explicit dense inverse;
 *   - A T + B is the banded stencil;  B and trC are evaluated from the end points;
 *   - each SDF sample reads its 4 cells once and yields value and gradient.
 */
#include <math.h>
#include "ocb_internal.h"
#include "../../include/orcdchomp_b200.h"

namespace
{

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
   return v;
}

/* deterministic block-wide sums of two values with ONE barrier: per-warp partials go to one
 * of two alternating 16-double buffers in red[0..31]; every thread then adds the partials in
 * warp order.  Two consecutive calls use different buffers and any later reuse of a buffer
 * is separated from its readers by the barrier of the call in between. */
__device__ __forceinline__ void block_sum2(double &v1, double &v2, double *red, int &parity)
{
   const int tid = threadIdx.x;
   const int nwarps = (blockDim.x + 31) >> 5;
   double *buf = red + 16 * parity;
   parity ^= 1;
   v1 = warp_sum(v1);
   v2 = warp_sum(v2);
   if ((tid & 31) == 0) { buf[2 * (tid >> 5)] = v1; buf[2 * (tid >> 5) + 1] = v2; }
   __syncthreads();
   double s1 = 0.0, s2 = 0.0;
   for (int w = 0; w < nwarps; w++) { s1 += buf[2 * w]; s2 += buf[2 * w + 1]; }
   v1 = s1;
   v2 = s2;
}

/* ------------------------------------------------------------------------- */
/* Shared-memory carve-up, computed identically on host and device. */
struct SmemLayout
{
   int T, G, AG, red, ws, cut2, radius; /* offsets in doubles */
   int sdf, sph, desc, mt, ired;        /* offsets in bytes   */
   int bytes;
};

__host__ __device__ inline SmemLayout smem_layout(const OcbChompArgs &a, int ws_in_smem)
{
   SmemLayout l;
   int d = 0;
   l.T = d; d += a.n * a.Ppad;
   l.G = d; d += a.n * a.Ppad;
   l.AG = d; d += a.use_momentum ? a.n * a.Ppad : 0;
   l.red = d; d += 36;
   l.ws = d; d += ws_in_smem ? (int) a.ws_stride : 0;
   l.cut2 = d; d += ws_in_smem ? a.nsa * (a.NAp + a.nsi) : 0;
   l.radius = d; d += ws_in_smem ? (a.nsa + a.nsi) : 0;
   int b = d * 8;
   l.sdf = b; b += a.nsdf * (int) sizeof(OcbSdfDev);
   l.sph = b; b += ws_in_smem ? a.nsa * (int) sizeof(OcbSphereDev) : 0;
   l.desc = b; b += ws_in_smem ? a.n_desc * 4 : 0;
   l.mt = b; b += a.use_hmc ? (626 + 626 + 16) * 4 : 0; /* state, saved copy (serial fallback), scratch */
   l.ired = b; b += 40 * 4;
   l.bytes = b;
   return l;
}

/* read-only tables: in shared memory for the common small-robot case, else in HBM */
struct Tables
{
   const OcbSphereDev *sph;
   const int *desc;
   const double *cut2;
   const double *radius;
   const OcbSdfDev *sdfs;
};

/* ------------------------------------------------------------------------- */
/* One step of the forward sweep over the compiled joint tree for waypoint t:
 * on return (R, tr) is joint j's frame after its motion and (ax, org) its axis
 * (local z) and a point on it, in the world frame.  Frames needed again at a
 * branch are saved to / loaded from the `slots` rows of ws. */
template <bool SAVE>
__device__ __forceinline__ void fk_step(const OcbJointDev &J, const double *__restrict__ Ts, double *__restrict__ slots,
                                        int Pp, int t, double R[9], double tr[3], double ax[3], double org[3])
{
   double Rn[9], tn[3];
   if (J.load == OCB_LOAD_BASE)
   {
#pragma unroll
      for (int k = 0; k < 9; k++) Rn[k] = J.XR[k];
      tn[0] = J.Xt[0]; tn[1] = J.Xt[1]; tn[2] = J.Xt[2];
   }
   else
   {
      if (J.load >= 0)
      {
         const double *sl = slots + 12 * J.load * Pp + t;
#pragma unroll
         for (int k = 0; k < 9; k++) R[k] = sl[k * Pp];
#pragma unroll
         for (int k = 0; k < 3; k++) tr[k] = sl[(9 + k) * Pp];
      }
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
#pragma unroll
         for (int c = 0; c < 3; c++)
            Rn[3 * r + c] = R[3 * r] * J.XR[c] + R[3 * r + 1] * J.XR[3 + c] + R[3 * r + 2] * J.XR[6 + c];
         tn[r] = R[3 * r] * J.Xt[0] + R[3 * r + 1] * J.Xt[1] + R[3 * r + 2] * J.Xt[2] + tr[r];
      }
   }
   const double q = Ts[J.dof * Pp + t];
   const double v = fma(J.c0, q, J.c1);
   ax[0] = Rn[2]; ax[1] = Rn[5]; ax[2] = Rn[8];
   org[0] = tn[0]; org[1] = tn[1]; org[2] = tn[2];
   if (J.type == OCB_JOINT_REVOLUTE)
   {
      double s, c;
      sincos(v, &s, &c);
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
         R[3 * r] = c * Rn[3 * r] + s * Rn[3 * r + 1];
         R[3 * r + 1] = c * Rn[3 * r + 1] - s * Rn[3 * r];
         R[3 * r + 2] = Rn[3 * r + 2];
         tr[r] = tn[r];
      }
   }
   else
   {
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
         R[3 * r] = Rn[3 * r];
         R[3 * r + 1] = Rn[3 * r + 1];
         R[3 * r + 2] = Rn[3 * r + 2];
         tr[r] = fma(v, Rn[3 * r + 2], tn[r]);
      }
   }
   if (SAVE && J.save >= 0)
   {
      double *sl = slots + 12 * J.save * Pp + t;
#pragma unroll
      for (int k = 0; k < 9; k++) sl[k * Pp] = R[k];
#pragma unroll
      for (int k = 0; k < 3; k++) sl[(9 + k) * Pp] = tr[k];
   }
}

/* forward kinematics of waypoint t: world positions of the active spheres.
 * Replaces SetActiveDOFValues + GetTransform()*pos (mod.cpp:1026-1038). */
__device__ __forceinline__ void fk_waypoint(const OcbChompArgs &a, const Tables &tb,
                                            const double *__restrict__ Ts, double *__restrict__ ws, int t)
{
   const int Pp = a.Ppad;
   double *slots = ws + 3 * a.nsa * Pp;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   for (int j = 0; j < a.nj; j++)
   {
      const OcbJointDev &J = a.joints[j];
      fk_step<true>(J, Ts, slots, Pp, t, R, tr, ax, org);
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
__device__ __forceinline__ void flush_wrenches(const OcbChompArgs &a, const Tables &tb,
                                               const double *__restrict__ Ts, double *__restrict__ ws,
                                               double *__restrict__ Gs, int t)
{
   const int Pp = a.Ppad;
   double *slots = ws + 3 * a.nsa * Pp;
   const double *Wg = ws + (3 * a.nsa + 12 * a.n_slots) * Pp + t;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   for (int j = 0; j < a.nj; j++)
   {
      const OcbJointDev &J = a.joints[j];
      fk_step<false>(J, Ts, slots, Pp, t, R, tr, ax, org);
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
}

/* ------------------------------------------------------------------------- */
/* one SDF sample: cell lookup (grid.c:191-209), first-order value
 * (grid.c:386-454) and one-sided gradient (grid.c:331-384) from the same four
 * cells.  Returns false when the point is outside the grid. */
__device__ __forceinline__ bool sdf_sample(const OcbSdfDev &S, const double g[3], double &val,
                                           double gg[3])
{
   int sub[3];
#pragma unroll
   for (int ax = 0; ax < 3; ax++)
   {
      if (g[ax] < 0.0 || g[ax] > S.length[ax]) return false;
      int s = (int) floor(g[ax] * S.scale[ax]);
      if (s >= S.size[ax]) s = S.size[ax] - 1;
      sub[ax] = s;
   }
   const long long stride0 = (long long) S.size[1] * S.size[2], stride1 = S.size[2];
   const long long idx = ((long long) sub[0] * S.size[1] + sub[1]) * S.size[2] + sub[2];
   double c0 = (0.5 + sub[0]) * S.cell[0], c1 = (0.5 + sub[1]) * S.cell[1], c2 = (0.5 + sub[2]) * S.cell[2];
   const bool nx0 = (sub[0] == 0) || (sub[0] != S.size[0] - 1 && !(g[0] < c0));
   const bool nx1 = (sub[1] == 0) || (sub[1] != S.size[1] - 1 && !(g[1] < c1));
   const bool nx2 = (sub[2] == 0) || (sub[2] != S.size[2] - 1 && !(g[2] < c2));
   const double c = __ldg(S.data + idx);
   const double n0 = __ldg(S.data + (nx0 ? idx + stride0 : idx - stride0));
   const double n1 = __ldg(S.data + (nx1 ? idx + stride1 : idx - stride1));
   const double n2 = __ldg(S.data + (nx2 ? idx + 1 : idx - 1));
   const double inf = HUGE_VAL;
   const bool bad = (c == inf) || (n0 == inf) || (n1 == inf) || (n2 == inf);
   const double s2 = (nx2 ? (n2 - c) : (c - n2)) * S.scale[2];
   const double s1 = (nx1 ? (n1 - c) : (c - n1)) * S.scale[1];
   const double s0 = (nx0 ? (n0 - c) : (c - n0)) * S.scale[0];
   double value = c;
   value = fma(s2, g[2] - c2, value);
   value = fma(s1, g[1] - c1, value);
   value = fma(s0, g[0] - c0, value);
   gg[0] = s0; gg[1] = s1; gg[2] = s2;
   val = bad ? inf : value;
   return true;
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
__device__ __forceinline__ double waypoint_cost(const OcbChompArgs &a, const Tables &tb,
                                                const double *__restrict__ Ts,
                                                double *__restrict__ ws, double *__restrict__ Gs,
                                                int t, bool want_grad)
{
   const int Pp = a.Ppad;
   const int nsa = a.nsa;
   const int row = a.NAp + a.nsi;
   double *Wg = ws + (3 * nsa + 12 * a.n_slots) * Pp + t;
   const double inv2dt = 1.0 / (2.0 * a.dt);
   const double invdt2 = 1.0 / (a.dt * a.dt);
   const double es = a.eps_self, inv_es = 1.0 / es, half_inv_es = 0.5 / es;
   const double eps = a.eps, inv_eps = 1.0 / eps, half_inv_eps = 0.5 / eps;
   double cost = 0.0;

   if (want_grad)
      for (int k = 0; k < 6 * a.ng; k++) Wg[k * Pp] = 0.0;

   for (int j = 0; j < a.nj; j++)
   {
      const OcbJointDev &J = a.joints[j];
      if (J.sph_begin == J.sph_end) continue;
      double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
      for (int s = J.sph_begin; s < J.sph_end; s++)
      {
         const double *ps = ws + 3 * s * Pp + t;
         double p[3], vel[3], acc[3];
#pragma unroll
         for (int k = 0; k < 3; k++)
         {
            const double pc = ps[k * Pp];
            const double pm = ps[k * Pp - 1];
            const double pp = ps[k * Pp + 1];
            p[k] = pc;
            vel[k] = (pp - pm) * inv2dt;
            acc[k] = (pc * -2.0 + pm + pp) * invdt2;
         }
         const double vn2 = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2];
         const double rv = rsqrt(vn2);
         const double vn = (vn2 > 0.0) ? vn2 * rv : 0.0;
         const double iv2 = rv * rv; /* 1 / |v|^2, unguarded (inf at rest), as mod.cpp:1239 */
         const bool moving = vn > 0.000001;
         const double radius = tb.radius[s];
         double cost_s = 0.0;
         double f[3] = {0.0, 0.0, 0.0};

         /* --- obstacle term: smallest interpolated field value wins (1169-1189) --- */
         int best = -1;
         double best_d = HUGE_VAL;
         double bg[3] = {0.0, 0.0, 0.0};
         for (int k = 0; k < a.nsdf; k++)
         {
            const OcbSdfDev &S = tb.sdfs[k];
            double g[3], d, gg[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
               g[r] = S.Rgw[3 * r] * p[0] + S.Rgw[3 * r + 1] * p[1] + S.Rgw[3 * r + 2] * p[2] + S.tgw[r];
            if (!sdf_sample(S, g, d, gg)) continue;
            if (d < best_d)
            {
               best_d = d;
               best = k;
               bg[0] = gg[0]; bg[1] = gg[1]; bg[2] = gg[2];
            }
         }
         if (best >= 0)
         {
            const double d = best_d - radius;
            if (d < 0.0)
               cost_s += vn * a.obs_factor * (0.5 * eps - d);
            else if (d < eps)
               cost_s += vn * a.obs_factor * half_inv_eps * (d - eps) * (d - eps);
            if (want_grad)
            {
               const OcbSdfDev &S = tb.sdfs[best];
               double x[3], cv[3];
               const double sc = (d < 0.0) ? -1.0 : ((d < eps) ? (d * inv_eps - 1.0) : 0.0);
               const double w = vn * a.obs_factor;
#pragma unroll
               for (int r = 0; r < 3; r++)
               {
                  const double gw = S.Rwg[3 * r] * bg[0] + S.Rwg[3 * r + 1] * bg[1] + S.Rwg[3 * r + 2] * bg[2];
                  x[r] = (d < eps) ? gw * sc * w : 0.0;
                  cv[r] = acc[r];
               }
               if (moving)
               {
                  const double pj = (x[0] * vel[0] + x[1] * vel[1] + x[2] * vel[2]) * iv2;
                  const double pc = (cv[0] * vel[0] + cv[1] * vel[1] + cv[2] * vel[2]) * iv2;
#pragma unroll
                  for (int r = 0; r < 3; r++)
                  {
                     x[r] = fma(-pj, vel[r], x[r]);
                     cv[r] = fma(-pc, vel[r], cv[r]);
                  }
               }
#pragma unroll
               for (int r = 0; r < 3; r++)
               {
                  x[r] = fma(-cost_s, cv[r] * iv2, x[r]);
                  f[r] = vn * x[r]; /* dgemv alpha = x_vel_norm (1244) */
               }
            }
         }

         /* --- self collision, each unordered pair once (1251-1317) --- */
         const double *crow = tb.cut2 + s * row;
         const double ws_self = vn * a.obs_factor_self;
         /* a partner within range: q its centre, po its column in ws (NULL: inactive, frozen) */
         auto in_range = [&](const double q[3], const double *po, int o)
         {
            const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            const double inv = rsqrt(d2);
            const double dist = d2 * inv;
            const double dd = dist - (radius + tb.radius[o]);
            /* cost shape shared by both directions (1281-1289) */
            const double cshape = (dd < 0.0) ? (0.5 * es - dd) : half_inv_es * (dd - es) * (dd - es);
            cost_s += ws_self * cshape;
            double w2 = 0.0, v2[3] = {0.0, 0.0, 0.0}, r2 = 0.0;
            bool moving2 = false;
            if (po)
            {
#pragma unroll
               for (int r = 0; r < 3; r++) v2[r] = (po[r * Pp + 1] - po[r * Pp - 1]) * inv2dt;
               const double v2n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
               r2 = rsqrt(v2n2);
               const double v2n = (v2n2 > 0.0) ? v2n2 * r2 : 0.0;
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
                  const double pj = (y[0] * v2[0] + y[1] * v2[1] + y[2] * v2[2]) * (r2 * r2);
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
         for (int i = 0; i < a.nsi; i++)
         {
            const double q[3] = {__ldg(a.inactive_pos + 3 * i), __ldg(a.inactive_pos + 3 * i + 1),
                                 __ldg(a.inactive_pos + 3 * i + 2)};
            const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            if (dx * dx + dy * dy + dz * dz <= crow[a.NAp + i]) in_range(q, nullptr, nsa + i);
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
   if (want_grad) flush_wrenches(a, tb, Ts, ws, Gs, t);
   return cost;
}

/* (A T)[i][j] for moving waypoint t = i+1 from the band of A (chomp.c:515-517, 665) */
__device__ __forceinline__ double band_AT(const OcbChompArgs &a, const double *__restrict__ Tj, int t)
{
   const int bw = a.bw, i = t - 1;
   const double *Ab = a.Aband + (size_t) i * (2 * bw + 1);
   double acc = 0.0;
   for (int k = -bw; k <= bw; k++)
   {
      const int i2 = i + k;
      if (i2 < 0 || i2 >= a.m) continue;
      acc = fma(__ldg(Ab + k + bw), Tj[t + k], acc);
   }
   return acc;
}

/* banded LDL^T solve in place on x[0..m) (one dof column); replaces the product
 * with the explicit inverse (chomp.c:529-530, 540-546, 640-641) */
__device__ __forceinline__ void band_solve(const OcbChompArgs &a, double *__restrict__ x)
{
   const int m = a.m, bw = a.bw;
   const double *__restrict__ Ls = a.Lband;
   const double *__restrict__ dinv = a.dinv;
   if (bw == 1)
   {
      double prev = x[0];
      for (int i = 1; i < m; i++)
      {
         prev = fma(-__ldg(Ls + i), prev, x[i]);
         x[i] = prev;
      }
      prev = x[m - 1] * __ldg(dinv + m - 1);
      x[m - 1] = prev;
      for (int i = m - 2; i >= 0; i--)
      {
         prev = fma(-__ldg(Ls + i + 1), prev, x[i] * __ldg(dinv + i));
         x[i] = prev;
      }
      return;
   }
   for (int i = 0; i < m; i++)
   {
      double acc = x[i];
      for (int k = 1; k <= bw && k <= i; k++) acc = fma(-__ldg(Ls + i * bw + (k - 1)), x[i - k], acc);
      x[i] = acc;
   }
   for (int i = m - 1; i >= 0; i--)
   {
      double acc = x[i] * __ldg(dinv + i);
      for (int k = 1; k <= bw && i + k < m; k++) acc = fma(-__ldg(Ls + (i + k) * bw + (k - 1)), x[i + k], acc);
      x[i] = acc;
   }
}

/* ------------------------------------------------------------------ MT19937 */
/* gsl_rng_mt19937 / gsl_ran_gaussian semantics (mod.cpp:2303-2304, 2763, 2767);
 * state = 624 words + index, one per run. */
__device__ __forceinline__ uint32_t mt_next(uint32_t *mt)
{
   uint32_t idx = mt[624];
   if (idx >= 624)
   {
      for (int k = 0; k < 624; k++)
      {
         const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
         mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
   }
   uint32_t y = mt[idx];
   mt[624] = idx + 1;
   y ^= (y >> 11);
   y ^= (y << 7) & 0x9d2c5680u;
   y ^= (y << 15) & 0xefc60000u;
   y ^= (y >> 18);
   return y;
}

__device__ __forceinline__ double mt_uniform(uint32_t *mt) { return mt_next(mt) / 4294967296.0; }

__device__ __forceinline__ double mt_uniform_pos(uint32_t *mt)
{
   double x;
   do { x = mt_uniform(mt); } while (x == 0.0);
   return x;
}

__device__ double mt_gaussian(uint32_t *mt, double sigma)
{
   double x, y, r2;
   do
   {
      x = -1.0 + 2.0 * mt_uniform_pos(mt);
      y = -1.0 + 2.0 * mt_uniform_pos(mt);
      r2 = x * x + y * y;
   } while (r2 > 1.0 || r2 == 0.0);
   return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

/* ---- block-parallel HMC momentum resample --------------------------------------------
 * Produces exactly the stream of the serial code above (gsl_ran_gaussian draws for AG in
 * row-major (i, j) order, then one gsl_rng_uniform), but cooperatively:
 *   - the MT19937 state is "twisted" 624 words at a time in three dependency phases;
 *   - the tempered words of a chunk are consumed as (x, y) pairs by all threads at once,
 *     accepted pairs are ranked by a block-wide prefix count so variate k lands in AG[k];
 *   - consumption stops at the pair that yields the last variate; the next raw word is
 *     the uniform for the resample gap.
 * gsl_rng_uniform_pos re-draws a zero word (probability 2^-32 each), which would shift the
 * pairing: if any zero word shows up the caller redoes the resample with the serial code
 * from a saved copy of the state, so the result is exact in every case.
 * scratch: >= 16 ints of shared memory.  Returns false when a zero word was met. */
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
   y ^= (y >> 11);
   y ^= (y << 7) & 0x9d2c5680u;
   y ^= (y << 15) & 0xefc60000u;
   y ^= (y >> 18);
   return y;
}

__device__ __forceinline__ uint32_t mt_twist_word(uint32_t a, uint32_t b, uint32_t far)
{
   const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
   return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ void mt_twist_parallel(uint32_t *mt)
{
   const int tid = threadIdx.x, NT = blockDim.x;
   /* phase bounds: [0,227) reads old only; [227,454) and [454,623) read words made by the
    * previous phase; word 623 reads new[0] and new[396] */
   const int lo[4] = {0, 227, 454, 623}, hi[4] = {227, 454, 623, 624};
   for (int ph = 0; ph < 4; ph++)
   {
      uint32_t val[8];
      int cnt = 0;
      for (int k = lo[ph] + tid; k < hi[ph]; k += NT)
         val[cnt++] = mt_twist_word(mt[k], mt[(k + 1) % 624], mt[(k + 397) % 624]);
      __syncthreads();
      cnt = 0;
      for (int k = lo[ph] + tid; k < hi[ph]; k += NT) mt[k] = val[cnt++];
      __syncthreads();
   }
}

__device__ bool hmc_resample_parallel(uint32_t *mt, int *scratch, double *AGs, int Pp, int m, int n, double sigma,
                                      double *uniform_out)
{
   const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
   const int need = m * n;
   int produced = 0;      /* variates written so far (uniform over the block) */
   int idx = (int) mt[624];
   bool have_x = false;   /* a pair's x was the last word of the previous chunk */
   double carry_x = 0.0;
   int *wtot = scratch;   /* [8] accepted pairs per warp */
   int *zero_seen = scratch + 8;
   int *stop_at = scratch + 9; /* word position right after the pair that produced the last variate */
   if (tid == 0) { *zero_seen = 0; *stop_at = -1; }
   __syncthreads();
   for (;;)
   {
      if (idx >= 624)
      {
         mt_twist_parallel(mt);
         idx = 0;
      }
      /* words idx..623 are available; with a carried x the first word is that pair's y */
      const int first = idx + (have_x ? 1 : 0);
      const int npairs = (624 - first) >> 1;
      const bool tail_x = ((624 - first) & 1) != 0; /* an x left without its y */
      bool done = false;
      /* the carried pair, handled by thread 0 as pair -1 of this chunk */
      for (int base = have_x ? -1 : 0; base < npairs && !done; base += NT)
      {
         const int j = base + tid;
         bool ok = false;
         double z = 0.0;
         int end_pos = 0;
         if (j < npairs)
         {
            double x, y;
            uint32_t wx, wy;
            if (j < 0) { wx = 1; wy = mt_temper(mt[idx]); x = carry_x; y = -1.0 + 2.0 * (wy / 4294967296.0); end_pos = idx + 1; }
            else
            {
               wx = mt_temper(mt[first + 2 * j]);
               wy = mt_temper(mt[first + 2 * j + 1]);
               x = -1.0 + 2.0 * (wx / 4294967296.0);
               y = -1.0 + 2.0 * (wy / 4294967296.0);
               end_pos = first + 2 * j + 2;
            }
            if (wx == 0 || wy == 0) atomicOr(zero_seen, 1);
            const double r2 = x * x + y * y;
            ok = !(r2 > 1.0 || r2 == 0.0);
            if (ok) z = sigma * y * sqrt(-2.0 * log(r2) / r2);
         }
         const unsigned bal = __ballot_sync(FULL_MASK, ok);
         if (lane == 0) wtot[warp] = __popc(bal);
         __syncthreads();
         int before = 0, total = 0;
         for (int w = 0; w < nwarps; w++)
         {
            const int c = wtot[w];
            if (w < warp) before += c;
            total += c;
         }
         const int rank = produced + before + __popc(bal & ((1u << lane) - 1u));
         if (ok && rank < need)
         {
            AGs[(rank % n) * Pp + (rank / n) + 1] = z; /* AG[i][j], i = rank / n */
            if (rank == need - 1) *stop_at = end_pos;
         }
         produced += total;
         __syncthreads();
         if (produced >= need) done = true;
      }
      if (*zero_seen) return false;
      if (done)
      {
         idx = *stop_at; /* uniform over the block after the barrier above */
         break;
      }
      /* chunk exhausted without finishing */
      if (tail_x)
      {
         const uint32_t wx = mt_temper(mt[623]);
         if (wx == 0) return false;
         carry_x = -1.0 + 2.0 * (wx / 4294967296.0);
         have_x = true;
      }
      else
         have_x = false;
      idx = 624;
   }
   /* one more raw word (zero allowed) for the resample gap */
   if (idx >= 624)
   {
      mt_twist_parallel(mt);
      idx = 0;
   }
   *uniform_out = mt_temper(mt[idx]) / 4294967296.0;
   __syncthreads();
   if (tid == 0) mt[624] = (uint32_t) (idx + 1);
   __syncthreads();
   return true;
}

/* ------------------------------------------------------------------------- */
struct ArgMax
{
   double v;
   int idx;
};

__device__ __forceinline__ ArgMax argmax_pick(ArgMax a, ArgMax b)
{
   /* largest value; first (lowest linear index) on ties, as the strict > of chomp.c:619-633 */
   if (b.v > a.v || (b.v == a.v && b.idx < a.idx)) return b;
   return a;
}

template <bool WS_SMEM, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, NT_MAX == 128 ? 3 : 1)
chomp_iterate_kernel(const __grid_constant__ OcbChompArgs a)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int tid = threadIdx.x;
   const int NT = blockDim.x;
   const int run = blockIdx.x;
   const int P = a.P, m = a.m, n = a.n, Pp = a.Ppad;

   /* ---- shared memory carve-up ---- */
   const SmemLayout lay = smem_layout(a, WS_SMEM ? 1 : 0);
   double *sd = reinterpret_cast<double *>(smem_raw);
   double *Ts = sd + lay.T;     /* [n][Pp] */
   double *Gs = sd + lay.G;     /* [n][Pp] */
   double *AGs = sd + lay.AG;   /* [n][Pp] (momentum only) */
   double *red = sd + lay.red;
   OcbSdfDev *sdfs = reinterpret_cast<OcbSdfDev *>(smem_raw + lay.sdf);
   uint32_t *mts = reinterpret_cast<uint32_t *>(smem_raw + lay.mt);
   int *ired = reinterpret_cast<int *>(smem_raw + lay.ired);
   double *ws;
   Tables tb;
   tb.sdfs = sdfs;
   if (WS_SMEM)
   {
      ws = sd + lay.ws;
      double *c2 = sd + lay.cut2, *rad = sd + lay.radius;
      OcbSphereDev *sph = reinterpret_cast<OcbSphereDev *>(smem_raw + lay.sph);
      int *dsc = reinterpret_cast<int *>(smem_raw + lay.desc);
      for (int e = tid; e < a.nsa * (a.NAp + a.nsi); e += NT) c2[e] = __ldg(a.cut2 + e);
      for (int e = tid; e < a.nsa + a.nsi; e += NT) rad[e] = __ldg(a.radius + e);
      {
         const int words = a.nsa * (int) (sizeof(OcbSphereDev) / 4);
         const uint32_t *src = reinterpret_cast<const uint32_t *>(a.spheres);
         uint32_t *dst = reinterpret_cast<uint32_t *>(sph);
         for (int e = tid; e < words; e += NT) dst[e] = __ldg(src + e);
      }
      for (int e = tid; e < a.n_desc; e += NT) dsc[e] = __ldg(a.desc + e);
      tb.sph = sph; tb.desc = dsc; tb.cut2 = c2; tb.radius = rad;
   }
   else
   {
      ws = a.ws_global + (size_t) run * a.ws_stride;
      tb.sph = a.spheres; tb.desc = a.desc; tb.cut2 = a.cut2; tb.radius = a.radius;
   }

   /* ---- stage per-run state and shared constants ---- */
   double *traj = a.traj + (size_t) run * P * n;
   for (int e = tid; e < P * n; e += NT) Ts[(e % n) * Pp + (e / n)] = traj[e];
   if (a.use_momentum)
   {
      const double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) AGs[(e % n) * Pp + (e / n) + 1] = ag[e];
   }
   {
      const int words = (int) (sizeof(OcbSdfDev) / 4) * a.nsdf;
      const uint32_t *src = reinterpret_cast<const uint32_t *>(a.sdfs);
      uint32_t *dst = reinterpret_cast<uint32_t *>(sdfs);
      for (int e = tid; e < words; e += NT) dst[e] = src[e];
   }
   if (a.use_hmc)
      for (int e = tid; e < 625; e += NT) mts[e] = a.mt_state[(size_t) run * 625 + e];
   int leapfrog_first = a.use_momentum ? a.leapfrog_first[run] : 0;
   int hmc_next = a.use_hmc ? a.hmc_next[run] : -1;
   int status = 0;
   const double inv_m = 1.0 / m;
   const double inv_lambda = 1.0 / a.lambda;

   /* end points define B and trC (chomp.c:278-296, 318-331) */
   double trC = 0.0;
   {
      double ss = 0.0, sg = 0.0, gg = 0.0;
      for (int j = 0; j < n; j++)
      {
         const double qs = traj[j], qg = traj[(size_t) (P - 1) * n + j];
         ss += qs * qs; sg += qs * qg; gg += qg * qg;
      }
      trC = 0.5 * (a.trc_ss * ss + 2.0 * a.trc_sg * sg + a.trc_gg * gg);
   }
   __syncthreads();

   double cost_obs = 0.0, cost_smooth = 0.0;
   int red_parity = 0;
   for (int iter = 0; iter <= a.n_iter; iter++)
   {
      const bool final_pass = (iter == a.n_iter);

      /* ---- HMC momentum resample (mod.cpp:2755-2768) ---- */
      if (a.use_hmc && !final_pass && iter == hmc_next)
      {
         const double alpha = 100.0 * exp(0.02 * iter);
         const double sigma = 1.0 / sqrt(alpha);
         uint32_t *saved = mts + 626;
         int *scratch = reinterpret_cast<int *>(mts + 1252);
         for (int e = tid; e < 625; e += NT) saved[e] = mts[e];
         __syncthreads();
         double u = 0.0;
         if (a.use_hmc == 2 || !hmc_resample_parallel(mts, scratch, AGs, Pp, m, n, sigma, &u))
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
      for (int t = tid; t < P; t += NT) fk_waypoint(a, tb, Ts, ws, t);
      __syncthreads();

      /* ---- obstacle + self-collision cost / gradient, then G = G/m + A T + B ---- */
      double csum = 0.0, ssum = 0.0;
      for (int t = tid + 1; t <= m; t += NT)
      {
         if (!final_pass)
            for (int j = 0; j < n; j++) Gs[j * Pp + t] = 0.0;
         csum += waypoint_cost(a, tb, Ts, ws, Gs, t, !final_pass);
         if (!final_pass)
         {
            const double bi = __ldg(a.bcoef_i + t - 1), bf = __ldg(a.bcoef_f + t - 1);
            for (int j = 0; j < n; j++)
            {
               const double *Tj = Ts + j * Pp;
               double g = Gs[j * Pp + t] * inv_m;
               if (a.grad_mode == 2) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
               g += band_AT(a, Tj, t) + (bi * Tj[0] + bf * Tj[P - 1]);
               Gs[j * Pp + t] = g;
               if (a.grad_mode == 1) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
            }
         }
         else
         {
            const double bi = __ldg(a.bcoef_i + t - 1), bf = __ldg(a.bcoef_f + t - 1);
            for (int j = 0; j < n; j++)
            {
               const double *Tj = Ts + j * Pp;
               const double b = bi * Tj[0] + bf * Tj[P - 1];
               ssum += (0.5 * band_AT(a, Tj, t) + b) * Tj[t];
            }
         }
      }
      if (final_pass)
      {
         block_sum2(csum, ssum, red, red_parity);
         cost_obs = csum * inv_m;
         cost_smooth = ssum + trC;
         break;
      }
      __syncthreads(); /* every row of G is complete */

      /* ---- AG = A^-1 G (banded solve, one thread per dof) ---- */
      if (tid < n) band_solve(a, Gs + tid * Pp + 1);
      __syncthreads();

      /* ---- momentum / plain update, T -= AG/lambda (chomp.c:525-548, 604-605); each thread
       * also checks the rows it has just written against the joint limits ---- */
      int violated = 0;
      {
         const double coef = (leapfrog_first ? 0.5 : 1.0) * inv_lambda;
         for (int t = tid + 1; t <= m; t += NT)
            for (int j = 0; j < n; j++)
            {
               double step = Gs[j * Pp + t];
               if (a.use_momentum)
               {
                  step = fma(coef, step, AGs[j * Pp + t]);
                  AGs[j * Pp + t] = step;
               }
               const double q = fma(-inv_lambda, step, Ts[j * Pp + t]);
               Ts[j * Pp + t] = q;
               violated |= (q < __ldg(a.lim_lo + j)) | (q > __ldg(a.lim_hi + j));
            }
         if (a.use_momentum) leapfrog_first = 0;
      }
      const int any_violation = __syncthreads_or(violated);

      /* ---- joint-limit projection (chomp.c:608-655) ---- */
      int round = 0;
      for (; any_violation && round < 1000; round++)
      {
         ArgMax best;
         best.v = 0.0;
         best.idx = 0x7fffffff;
         for (int t = tid + 1; t <= m; t += NT)
            for (int j = 0; j < n; j++)
            {
               const double q = Ts[j * Pp + t];
               const double lo = __ldg(a.lim_lo + j), hi = __ldg(a.lim_hi + j);
               double v = 0.0;
               if (q < lo) v = lo - q;
               if (q > hi) v = hi - q;
               Gs[j * Pp + t] = v;
               ArgMax c;
               c.v = fabs(v);
               c.idx = (t - 1) * n + j;
               if (c.v > 0.0) best = argmax_pick(best, c);
            }
#pragma unroll
         for (int o = 16; o > 0; o >>= 1)
         {
            ArgMax other;
            other.v = __shfl_xor_sync(FULL_MASK, best.v, o);
            other.idx = __shfl_xor_sync(FULL_MASK, best.idx, o);
            best = argmax_pick(best, other);
         }
         if ((tid & 31) == 0) { red[tid >> 5] = best.v; ired[tid >> 5] = best.idx; }
         __syncthreads();
         if (tid == 0)
         {
            ArgMax b;
            b.v = red[0];
            b.idx = ired[0];
            for (int w = 1; w < ((NT + 31) >> 5); w++)
            {
               ArgMax c;
               c.v = red[w];
               c.idx = ired[w];
               b = argmax_pick(b, c);
            }
            red[33] = b.v;
            ired[33] = b.idx;
            if (b.v > 0.0)
               red[34] = Gs[(b.idx % n) * Pp + (b.idx / n) + 1]; /* signed violation at the arg-max */
         }
         __syncthreads();
         const double worst = red[33];
         const int worst_idx = ired[33];
         if (worst == 0.0) break;
         if (tid < n) band_solve(a, Gs + tid * Pp + 1);
         __syncthreads();
         const double scale = 1.01 * red[34] / Gs[(worst_idx % n) * Pp + (worst_idx / n) + 1];
         for (int t = tid + 1; t <= m; t += NT)
            for (int j = 0; j < n; j++) Ts[j * Pp + t] = fma(scale, Gs[j * Pp + t], Ts[j * Pp + t]);
         __syncthreads();
      }
      if (round >= 1000)
      {
         status = OCB_ERR_JLIMIT; /* chomp.c:651-655 returns -1 before the smoothness cost */
         break;
      }

      /* ---- smoothness cost of the updated trajectory (chomp.c:660-671) ---- */
      ssum = 0.0;
      for (int t = tid + 1; t <= m; t += NT)
      {
         const double bi = __ldg(a.bcoef_i + t - 1), bf = __ldg(a.bcoef_f + t - 1);
         for (int j = 0; j < n; j++)
         {
            const double *Tj = Ts + j * Pp;
            const double b = bi * Tj[0] + bf * Tj[P - 1];
            ssum += (0.5 * band_AT(a, Tj, t) + b) * Tj[t];
         }
      }
      block_sum2(csum, ssum, red, red_parity);
      cost_obs = csum * inv_m;
      cost_smooth = ssum + trC;
      if (a.trace_on && tid == 0)
      {
         double *tr = a.trace + ((size_t) run * a.n_iter + iter) * 3;
         tr[0] = cost_obs + cost_smooth;
         tr[1] = cost_obs;
         tr[2] = cost_smooth;
      }
   }

   /* ---- write the run back ---- */
   __syncthreads();
   for (int e = tid; e < P * n; e += NT) traj[e] = Ts[(e % n) * Pp + (e / n)];
   if (a.use_momentum)
   {
      double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) ag[e] = AGs[(e % n) * Pp + (e / n) + 1];
   }
   if (a.use_hmc)
      for (int e = tid; e < 625; e += NT) a.mt_state[(size_t) run * 625 + e] = mts[e];
   if (tid == 0)
   {
      if (a.use_momentum) a.leapfrog_first[run] = leapfrog_first;
      if (a.use_hmc) a.hmc_next[run] = hmc_next;
      a.costs[(size_t) run * 3 + 0] = cost_obs + cost_smooth;
      a.costs[(size_t) run * 3 + 1] = cost_obs;
      a.costs[(size_t) run * 3 + 2] = cost_smooth;
      a.status[run] = status;
   }
}

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

} /* namespace */

extern "C" size_t ocb_chomp_smem_bytes(const OcbChompArgs *a, int ws_in_smem)
{
   return (size_t) smem_layout(*a, ws_in_smem).bytes;
}

template <bool WS_SMEM, int NT_MAX>
static cudaError_t launch_variant(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st)
{
   static size_t configured = 0;
   if (smem_bytes > configured)
   {
      cudaError_t e = cudaFuncSetAttribute(chomp_iterate_kernel<WS_SMEM, NT_MAX>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
      if (e != cudaSuccess) return e;
      configured = smem_bytes;
   }
   chomp_iterate_kernel<WS_SMEM, NT_MAX><<<args->R, threads, smem_bytes, st>>>(*args);
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_chomp(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st)
{
   if (threads > 256 || threads % 32) return cudaErrorInvalidValue;
   if (args->ws_in_smem)
      return threads <= 128 ? launch_variant<true, 128>(args, smem_bytes, threads, st)
                            : launch_variant<true, 256>(args, smem_bytes, threads, st);
   return threads <= 128 ? launch_variant<false, 128>(args, smem_bytes, threads, st)
                         : launch_variant<false, 256>(args, smem_bytes, threads, st);
}

extern "C" cudaError_t ocb_launch_init_traj(double *traj, const double *q_start, const double *q_goal,
                                            int R, int P, int n, cudaStream_t st)
{
   const size_t total = (size_t) R * P * n;
   int blocks = (int) ((total + 255) / 256);
   if (blocks > 148 * 8) blocks = 148 * 8;
   if (blocks < 1) blocks = 1;
   init_traj_kernel<<<blocks, 256, 0, st>>>(traj, q_start, q_goal, R, P, n);
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_best(const double *costs, const int *status, int R, int *best_run,
                                       double *best_cost, cudaStream_t st)
{
   best_kernel<<<1, 1024, 0, st>>>(costs, status, R, best_run, best_cost);
   return cudaGetLastError();
}
