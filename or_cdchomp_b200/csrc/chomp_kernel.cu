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

/* deterministic block-wide sum; red has 33 doubles; every thread gets the result */
__device__ __forceinline__ double block_sum(double v, double *red)
{
   const int tid = threadIdx.x;
   const int nwarps = (blockDim.x + 31) >> 5;
   v = warp_sum(v);
   if ((tid & 31) == 0) red[tid >> 5] = v;
   __syncthreads();
   if (tid < 32)
   {
      double x = (tid < nwarps) ? red[tid] : 0.0;
      x = warp_sum(x);
      if (tid == 0) red[32] = x;
   }
   __syncthreads();
   double out = red[32];
   __syncthreads();
   return out;
}

/* ------------------------------------------------------------------------- */
/* forward kinematics of waypoint t: sphere centres, joint axes and origins.
 * Replaces SetActiveDOFValues + GetTransform()*pos (mod.cpp:1026-1038). */
__device__ __forceinline__ void fk_waypoint(const OcbChompArgs &a, const double *__restrict__ Ts,
                                            double *__restrict__ ws, int t)
{
   const int Pp = a.Ppad;
   double *jax = ws + (size_t) 3 * a.nsa * Pp;
   double *slots = jax + (size_t) 6 * a.nj * Pp;
   double R[9], tr[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;

   for (int j = 0; j < a.nj; j++)
   {
      const OcbJointDev &J = a.joints[j];
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
            const double *sl = slots + (size_t) 12 * J.load * Pp + t;
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = sl[(size_t) k * Pp];
#pragma unroll
            for (int k = 0; k < 3; k++) tr[k] = sl[(size_t) (9 + k) * Pp];
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
      /* joint axis (local z) and a point on it, in the world frame */
      double *jx = jax + (size_t) 6 * j * Pp + t;
      jx[0] = Rn[2];
      jx[(size_t) Pp] = Rn[5];
      jx[(size_t) 2 * Pp] = Rn[8];
      jx[(size_t) 3 * Pp] = tn[0];
      jx[(size_t) 4 * Pp] = tn[1];
      jx[(size_t) 5 * Pp] = tn[2];
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
      if (J.save >= 0)
      {
         double *sl = slots + (size_t) 12 * J.save * Pp + t;
#pragma unroll
         for (int k = 0; k < 9; k++) sl[(size_t) k * Pp] = R[k];
#pragma unroll
         for (int k = 0; k < 3; k++) sl[(size_t) (9 + k) * Pp] = tr[k];
      }
      for (int s = J.sph_begin; s < J.sph_end; s++)
      {
         const double px = __ldg(&a.spheres[s].pos[0]);
         const double py = __ldg(&a.spheres[s].pos[1]);
         const double pz = __ldg(&a.spheres[s].pos[2]);
         double *o = ws + (size_t) 3 * s * Pp + t;
         o[0] = R[0] * px + R[1] * py + R[2] * pz + tr[0];
         o[(size_t) Pp] = R[3] * px + R[4] * py + R[5] * pz + tr[1];
         o[(size_t) 2 * Pp] = R[6] * px + R[7] * py + R[8] * pz + tr[2];
      }
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
   const size_t stride[3] = {(size_t) S.size[1] * S.size[2], (size_t) S.size[2], 1};
   const size_t idx = ((size_t) sub[0] * S.size[1] + sub[1]) * S.size[2] + sub[2];
   double centre[3];
   bool next[3];
   size_t nb[3];
#pragma unroll
   for (int ax = 0; ax < 3; ax++)
   {
      centre[ax] = (0.5 + sub[ax]) * S.cell[ax];
      next[ax] = (sub[ax] == 0) || (sub[ax] != S.size[ax] - 1 && !(g[ax] < centre[ax]));
      nb[ax] = next[ax] ? idx + stride[ax] : idx - stride[ax];
   }
   const double c = __ldg(S.data + idx);
   const double n0 = __ldg(S.data + nb[0]);
   const double n1 = __ldg(S.data + nb[1]);
   const double n2 = __ldg(S.data + nb[2]);
   const double nbv[3] = {n0, n1, n2};
   const double inf = HUGE_VAL;
   bool bad = (c == inf) || (n0 == inf) || (n1 == inf) || (n2 == inf);
   double value = c;
#pragma unroll
   for (int ax = 2; ax >= 0; ax--)
   {
      const double diff = next[ax] ? (nbv[ax] - c) : (c - nbv[ax]);
      const double slope = diff * S.scale[ax];
      gg[ax] = slope;
      value = fma(slope, g[ax] - centre[ax], value);
   }
   val = bad ? inf : value;
   return true;
}

/* ------------------------------------------------------------------------- */
/* cost (and, when want_grad, the configuration-space gradient row) of moving
 * waypoint t (1..P-2).  Restates sphere_cost (mod.cpp:1134-1327) on top of the
 * finite differences of sphere_cost_pre (mod.cpp:1099-1127). */
__device__ __forceinline__ double waypoint_cost(const OcbChompArgs &a, const OcbSdfDev *__restrict__ sdfs,
                                                const double *__restrict__ ws, double *__restrict__ Gs,
                                                int t, bool want_grad)
{
   const int Pp = a.Ppad;
   const double *jax = ws + (size_t) 3 * a.nsa * Pp;
   const double inv2dt = 1.0 / (2.0 * a.dt);
   const double invdt2 = 1.0 / (a.dt * a.dt);
   double cost = 0.0;

   for (int j = 0; j < a.nj; j++)
   {
      const OcbJointDev &J = a.joints[j];
      if (J.sph_begin == J.sph_end) continue;
      double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
      for (int s = J.sph_begin; s < J.sph_end; s++)
      {
         const double *ps = ws + (size_t) 3 * s * Pp + t;
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
         const double vn = sqrt(vn2);
         const double radius = __ldg(&a.spheres[s].radius);
         double cost_s = 0.0;
         double f[3] = {0.0, 0.0, 0.0};

         /* --- obstacle term: smallest interpolated field value wins (1169-1189) --- */
         int best = -1;
         double best_d = HUGE_VAL;
         double bg[3] = {0.0, 0.0, 0.0};
         for (int k = 0; k < a.nsdf; k++)
         {
            const OcbSdfDev &S = sdfs[k];
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
            const double eps = a.eps;
            if (d < 0.0)
               cost_s += vn * a.obs_factor * (0.5 * eps - d);
            else if (d < eps)
               cost_s += vn * a.obs_factor * (0.5 / eps) * (d - eps) * (d - eps);
            if (want_grad)
            {
               const OcbSdfDev &S = sdfs[best];
               double x[3], cv[3];
               const double sc = (d < 0.0) ? -1.0 : ((d < eps) ? (d / eps - 1.0) : 0.0);
               const double w = vn * a.obs_factor;
#pragma unroll
               for (int r = 0; r < 3; r++)
               {
                  const double gw = S.Rwg[3 * r] * bg[0] + S.Rwg[3 * r + 1] * bg[1] + S.Rwg[3 * r + 2] * bg[2];
                  x[r] = (d < eps) ? gw * sc * w : 0.0;
                  cv[r] = acc[r];
               }
               if (vn > 0.000001)
               {
                  const double pj = (x[0] * vel[0] + x[1] * vel[1] + x[2] * vel[2]) / vn2;
                  const double pc = (cv[0] * vel[0] + cv[1] * vel[1] + cv[2] * vel[2]) / vn2;
#pragma unroll
                  for (int r = 0; r < 3; r++)
                  {
                     x[r] = fma(-pj, vel[r], x[r]);
                     cv[r] = fma(-pc, vel[r], cv[r]);
                  }
               }
               const double iv2 = 1.0 / vn2; /* unguarded, as mod.cpp:1239 */
#pragma unroll
               for (int r = 0; r < 3; r++)
               {
                  x[r] = fma(-cost_s, cv[r] * iv2, x[r]);
                  f[r] = vn * x[r]; /* dgemv alpha = x_vel_norm (1244) */
               }
            }
         }

         /* --- self collision against spheres of other links (1251-1317) --- */
         const int pb = __ldg(&a.spheres[s].pair_begin);
         const int pe = __ldg(&a.spheres[s].pair_end);
         for (int pi = pb; pi < pe; pi++)
         {
            const int o = __ldg(&a.pairs[pi].other);
            const double cut2 = __ldg(&a.pairs[pi].cut2);
            double q[3];
            if (o < a.nsa)
            {
               const double *po = ws + (size_t) 3 * o * Pp + t;
               q[0] = po[0]; q[1] = po[(size_t) Pp]; q[2] = po[(size_t) 2 * Pp];
            }
            else
            {
               const double *po = a.inactive_pos + 3 * (o - a.nsa);
               q[0] = __ldg(po); q[1] = __ldg(po + 1); q[2] = __ldg(po + 2);
            }
            const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 > cut2) continue;
            const double rsum = __ldg(&a.pairs[pi].rsum);
            const double dist = sqrt(d2);
            const double dd = dist - rsum;
            const double es = a.eps_self;
            if (dd < 0.0)
               cost_s += vn * a.obs_factor_self * (0.5 * es - dd);
            else
               cost_s += vn * a.obs_factor_self * (0.5 / es) * (dd - es) * (dd - es);
            if (want_grad)
            {
               const double sc = (dd < 0.0) ? -1.0 : ((dd < es) ? (dd / es - 1.0) : 1.0);
               const double gh[3] = {dx / dist, dy / dist, dz / dist};
               double x[3];
               const double w = vn * a.obs_factor_self;
#pragma unroll
               for (int r = 0; r < 3; r++) x[r] = gh[r] * sc * w;
               if (vn > 0.000001)
               {
                  const double pj = (x[0] * vel[0] + x[1] * vel[1] + x[2] * vel[2]) / vn2;
#pragma unroll
                  for (int r = 0; r < 3; r++) x[r] = fma(-pj, vel[r], x[r]);
               }
#pragma unroll
               for (int r = 0; r < 3; r++) f[r] += x[r];
               if (o < a.nsa)
               {
                  /* the same pair seen from the other sphere: (J2 - J)^T x2 puts -x2 on us */
                  const double *po = ws + (size_t) 3 * o * Pp + t;
                  double v2[3];
#pragma unroll
                  for (int r = 0; r < 3; r++)
                     v2[r] = (po[r * Pp + 1] - po[r * Pp - 1]) * inv2dt;
                  const double v2n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
                  const double v2n = sqrt(v2n2);
                  const double w2 = v2n * a.obs_factor_self;
                  double y[3];
#pragma unroll
                  for (int r = 0; r < 3; r++) y[r] = -gh[r] * sc * w2;
                  if (v2n > 0.000001)
                  {
                     const double pj = (y[0] * v2[0] + y[1] * v2[1] + y[2] * v2[2]) / v2n2;
#pragma unroll
                     for (int r = 0; r < 3; r++) y[r] = fma(-pj, v2[r], y[r]);
                  }
#pragma unroll
                  for (int r = 0; r < 3; r++) f[r] -= y[r];
               }
            }
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
         /* J^T f for the whole joint frame: every ancestor joint sees the wrench */
         for (int ai = J.anc_begin; ai < J.anc_end; ai++)
         {
            const int aj = __ldg(&a.ancs[ai].joint);
            const int dof = __ldg(&a.ancs[ai].dof);
            const int type = __ldg(&a.ancs[ai].type);
            const double c0 = __ldg(&a.ancs[ai].c0);
            const double *jx = jax + (size_t) 6 * aj * Pp + t;
            const double ax = jx[0], ay = jx[(size_t) Pp], az = jx[(size_t) 2 * Pp];
            double val;
            if (type == OCB_JOINT_REVOLUTE)
            {
               const double ox = jx[(size_t) 3 * Pp], oy = jx[(size_t) 4 * Pp], oz = jx[(size_t) 5 * Pp];
               const double mx = M[0] - (oy * F[2] - oz * F[1]);
               const double my = M[1] - (oz * F[0] - ox * F[2]);
               const double mz = M[2] - (ox * F[1] - oy * F[0]);
               val = ax * mx + ay * my + az * mz;
            }
            else
               val = ax * F[0] + ay * F[1] + az * F[2];
            Gs[dof * Pp + t] = fma(c0, val, Gs[dof * Pp + t]);
         }
      }
   }
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
__device__ __forceinline__ void band_solve(const OcbChompArgs &a, const double *__restrict__ Ls,
                                           const double *__restrict__ dinv, double *__restrict__ x)
{
   const int m = a.m, bw = a.bw;
   if (bw == 1)
   {
      double prev = x[0];
      for (int i = 1; i < m; i++)
      {
         prev = fma(-Ls[i], prev, x[i]);
         x[i] = prev;
      }
      prev = x[m - 1] * dinv[m - 1];
      x[m - 1] = prev;
      for (int i = m - 2; i >= 0; i--)
      {
         prev = fma(-Ls[i + 1], prev, x[i] * dinv[i]);
         x[i] = prev;
      }
      return;
   }
   for (int i = 0; i < m; i++)
   {
      double acc = x[i];
      for (int k = 1; k <= bw && k <= i; k++) acc = fma(-Ls[i * bw + (k - 1)], x[i - k], acc);
      x[i] = acc;
   }
   for (int i = m - 1; i >= 0; i--)
   {
      double acc = x[i] * dinv[i];
      for (int k = 1; k <= bw && i + k < m; k++) acc = fma(-Ls[(i + k) * bw + (k - 1)], x[i + k], acc);
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

__global__ void __launch_bounds__(256)
chomp_iterate_kernel(const __grid_constant__ OcbChompArgs a)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int tid = threadIdx.x;
   const int NT = blockDim.x;
   const int run = blockIdx.x;
   const int P = a.P, m = a.m, n = a.n, Pp = a.Ppad, bw = a.bw;

   /* ---- shared memory carve-up (doubles first) ---- */
   double *Ts = reinterpret_cast<double *>(smem_raw);        /* [n][Pp] */
   double *Gs = Ts + (size_t) n * Pp;                        /* [n][Pp] */
   double *AGs = Gs + (size_t) n * Pp;                       /* [n][Pp] (momentum only) */
   double *Ls = AGs + (a.use_momentum ? (size_t) n * Pp : 0);/* [m][bw] */
   double *dinv = Ls + (size_t) m * bw;                      /* [m] */
   double *red = dinv + m;                                   /* [36] */
   double *wsS = red + 36;
   OcbSdfDev *sdfs = reinterpret_cast<OcbSdfDev *>(wsS + (a.ws_in_smem ? a.ws_stride : 0));
   uint32_t *mts = reinterpret_cast<uint32_t *>(sdfs + a.nsdf); /* [625] (hmc only) */
   int *ired = reinterpret_cast<int *>(mts + (a.use_hmc ? 626 : 0)); /* [40] */
   double *ws = a.ws_in_smem ? wsS : (a.ws_global + (size_t) run * a.ws_stride);

   /* ---- stage per-run state and shared constants ---- */
   double *traj = a.traj + (size_t) run * P * n;
   for (int e = tid; e < P * n; e += NT) Ts[(e % n) * Pp + (e / n)] = traj[e];
   if (a.use_momentum)
   {
      const double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) AGs[(e % n) * Pp + (e / n) + 1] = ag[e];
   }
   for (int e = tid; e < m * bw; e += NT) Ls[e] = __ldg(a.Lband + e);
   for (int e = tid; e < m; e += NT) dinv[e] = __ldg(a.dinv + e);
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
   for (int iter = 0; iter <= a.n_iter; iter++)
   {
      const bool final_pass = (iter == a.n_iter);

      /* ---- HMC momentum resample (mod.cpp:2755-2768) ---- */
      if (a.use_hmc && !final_pass && iter == hmc_next)
      {
         if (tid == 0)
         {
            const double alpha = 100.0 * exp(0.02 * iter);
            const double sigma = 1.0 / sqrt(alpha);
            for (int i = 0; i < m; i++)
               for (int j = 0; j < n; j++) AGs[j * Pp + i + 1] = mt_gaussian(mts, sigma);
            const double u = mt_uniform(mts);
            ired[32] = hmc_next + 1 + (int) (-log(u) / a.hmc_lambda);
         }
         __syncthreads();
         hmc_next = ired[32];
         leapfrog_first = 1;
         __syncthreads();
      }

      /* ---- forward kinematics of all P waypoints ---- */
      for (int t = tid; t < P; t += NT) fk_waypoint(a, Ts, ws, t);
      __syncthreads();

      /* ---- obstacle + self-collision cost / gradient, then G = G/m + A T + B ---- */
      double csum = 0.0, ssum = 0.0;
      for (int t = tid + 1; t <= m; t += NT)
      {
         if (!final_pass)
            for (int j = 0; j < n; j++) Gs[j * Pp + t] = 0.0;
         csum += waypoint_cost(a, sdfs, ws, Gs, t, !final_pass);
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
      cost_obs = block_sum(csum, red) * inv_m;
      if (final_pass)
      {
         cost_smooth = block_sum(ssum, red) + trC;
         break;
      }

      /* ---- AG = A^-1 G (banded solve, one thread per dof) ---- */
      if (tid < n) band_solve(a, Ls, dinv, Gs + tid * Pp + 1);
      __syncthreads();

      /* ---- momentum / plain update, T -= AG/lambda (chomp.c:525-548, 604-605) ---- */
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
               Ts[j * Pp + t] = fma(-inv_lambda, step, Ts[j * Pp + t]);
            }
         if (a.use_momentum) leapfrog_first = 0;
      }
      __syncthreads();

      /* ---- joint-limit projection (chomp.c:608-655) ---- */
      int round = 0;
      for (; round < 1000; round++)
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
         if (tid < n) band_solve(a, Ls, dinv, Gs + tid * Pp + 1);
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
      cost_smooth = block_sum(ssum, red) + trC;
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
   size_t d = (size_t) 2 * a->n * a->Ppad;
   if (a->use_momentum) d += (size_t) a->n * a->Ppad;
   d += (size_t) a->m * a->bw + a->m + 36;
   if (ws_in_smem) d += a->ws_stride;
   size_t bytes = d * sizeof(double) + (size_t) a->nsdf * sizeof(OcbSdfDev);
   if (a->use_hmc) bytes += 626 * sizeof(uint32_t);
   bytes += 40 * sizeof(int);
   return bytes;
}

extern "C" cudaError_t ocb_launch_chomp(const OcbChompArgs *args, size_t smem_bytes, int threads, cudaStream_t st)
{
   static size_t configured = 0;
   if (smem_bytes > configured)
   {
      cudaError_t e = cudaFuncSetAttribute(chomp_iterate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int) smem_bytes);
      if (e != cudaSuccess) return e;
      configured = smem_bytes;
   }
   chomp_iterate_kernel<<<args->R, threads, smem_bytes, st>>>(*args);
   return cudaGetLastError();
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
