/* chomp_tiled.cu -- the CHOMP iteration for robots whose sphere model does not fit the
 * persistent kernel's shared-memory workspace (hundreds of spheres, long trajectories).
 *
 * The persistent kernel (chomp_kernel.cu) gives one thread a whole waypoint; with hundreds of
 * spheres the per-run workspace (3 doubles per sphere per waypoint) outgrows shared memory
 * and the quadratic self-collision sweep leaves most of the GPU idle when there are few
 * runs.  Here one iteration is two launches:
 *
 *   chomp_tile_cost_kernel    grid = runs x tiles of TW waypoints, 256 threads.  The block
 *       runs the forward kinematics of its TW + 2 waypoints once into shared memory, then
 *       splits the SPHERES over its 256 / TW workers (a worker = TW lanes, one per waypoint):
 *       every worker evaluates obstacle and self-collision terms of its own spheres against
 *       all partners and folds the forces into joint space through the stored joint axes.
 *       Output: the obstacle part of the gradient for the tile's rows and one cost partial.
 *   chomp_run_update_kernel   grid = runs.  G = G_obs / m + A T + B, banded solve, momentum /
 *       HMC, joint-limit projection, smoothness cost -- the same steps, through the same
 *       device functions, as the persistent kernel.
 *
 * What it replaces in the reference is what chomp_kernel.cu lists (cd_chomp_iterate,
 * sphere_cost_pre, sphere_cost); only the decomposition differs.
 *
 * Self collision: the reference visits every ordered pair (mod.cpp:1251-1317); here each
 * unordered pair {s, o > s} is visited once by the worker that owns s.  Both directed terms
 * are formed, their difference is the net force on s, its negative the reaction on o.  The
 * reaction is summed per partner joint frame in registers while the partner sweep stays in
 * that frame's sphere range and is folded into joint space when the sweep leaves it, so no
 * worker touches accumulators owned by another one.
 */
#include "chomp_device.cuh"
#include "chomp_constraints.cuh"

namespace
{

constexpr int TILE_THREADS = 256;

struct TileLayout
{
   int pos, jfr, slots, Gp, cp, rad, gb; /* offsets in doubles */
   int sdf, link;                    /* offsets in bytes   */
   int bytes;
};

__host__ __device__ inline TileLayout tile_layout(const OcbChompArgs &a, int TW)
{
   TileLayout l;
   const int CS = TW + 2, NW = TILE_THREADS / TW;
   int d = 0;
   l.pos = d; d += 3 * a.nsa * CS;      /* sphere centres, [3 nsa][CS]; column c <-> waypoint t_first - 1 + c */
   l.jfr = d; d += 6 * a.nj * TW;       /* joint axis + origin, [6 nj][TW] */
   l.slots = d; d += 12 * a.n_slots * CS;
   l.Gp = d; d += NW * a.n * TW;        /* per-worker gradient rows, [NW][n][TW] */
   l.cp = d; d += TILE_THREADS;         /* per-thread cost */
   l.rad = d; d += a.nsa;
   l.gb = d; d += 3 * a.nj * CS;        /* world centres of the joint frames' bounding spheres, [3 nj][CS] */
   int b = d * 8;
   l.sdf = b; b += a.nsdf * (int) sizeof(OcbSdfDev);
   l.link = b; b += a.nsa * 4;
   l.bytes = b;
   return l;
}

/* finite differences of sphere s at column c (mod.cpp:1099-1127) */
struct SphereState
{
   double p[3], vel[3], acc[3];
   double vn, iv2;
   bool moving;
};

__device__ __forceinline__ void sphere_state(const double *__restrict__ pcol, int CS, int s, double inv2dt,
                                             double invdt2, SphereState &S)
{
   const double *ps = pcol + 3 * s * CS;
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      const double pc = ps[k * CS], pm = ps[k * CS - 1], pp = ps[k * CS + 1];
      S.p[k] = pc;
      S.vel[k] = (pp - pm) * inv2dt;
      S.acc[k] = (pc * -2.0 + pm + pp) * invdt2;
   }
   const double vn2 = S.vel[0] * S.vel[0] + S.vel[1] * S.vel[1] + S.vel[2] * S.vel[2];
   speed_terms(vn2, S.vn, S.iv2); /* unguarded, as mod.cpp:1239 */
   S.moving = S.vn > 0.000001;
}

/* sphere s (state S) is within range of a partner centred at q (radius sum rsum).  Adds the
 * cost of the pair -- seen from s and, for an active partner (po = its centre column), from
 * the partner as well -- to cost_s and returns in x the net workspace force on s,
 * x(s, o) - x(o, s); the reaction on an active partner is -x. */
struct PairConst
{
   double es, inv_es, half_inv_es, inv2dt, obs_factor_self;
};

__device__ __forceinline__ void self_pair_term(const PairConst &K, const SphereState &S, const double q[3],
                                               const double *__restrict__ po, int CS, double rsum,
                                               bool want_grad, double &cost_s, double x[3])
{
   const double es = K.es, inv_es = K.inv_es, half_inv_es = K.half_inv_es;
   const double dx = S.p[0] - q[0], dy = S.p[1] - q[1], dz = S.p[2] - q[2];
   const double d2 = dx * dx + dy * dy + dz * dz;
   const double inv = fast_rsqrt(d2);
   const double dd = d2 * inv - rsum;
   const double cshape = (dd < 0.0) ? (0.5 * es - dd) : half_inv_es * (dd - es) * (dd - es);
   const double ws_self = S.vn * K.obs_factor_self;
   cost_s += ws_self * cshape;
   double v2[3] = {0.0, 0.0, 0.0}, r2 = 0.0, w2 = 0.0;
   bool moving2 = false;
   if (po)
   {
      const double inv2dt = K.inv2dt;
#pragma unroll
      for (int r = 0; r < 3; r++) v2[r] = (po[r * CS + 1] - po[r * CS - 1]) * inv2dt;
      const double v2n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
      double v2n;
      speed_terms(v2n2, v2n, r2); /* r2: 1 / |v2|^2 */
      moving2 = v2n > 0.000001;
      w2 = v2n * K.obs_factor_self;
      cost_s += w2 * cshape; /* the partner's own cost_sphere term */
   }
   x[0] = x[1] = x[2] = 0.0;
   if (!want_grad) return;
   const double sc = (dd < 0.0) ? -1.0 : ((dd < es) ? (dd * inv_es - 1.0) : 1.0);
   const double gh[3] = {dx * inv, dy * inv, dz * inv};
   const double wa = sc * ws_self;
#pragma unroll
   for (int r = 0; r < 3; r++) x[r] = gh[r] * wa;
   if (S.moving)
   {
      const double pj = (x[0] * S.vel[0] + x[1] * S.vel[1] + x[2] * S.vel[2]) * S.iv2;
#pragma unroll
      for (int r = 0; r < 3; r++) x[r] = fma(-pj, S.vel[r], x[r]);
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
#pragma unroll
      for (int r = 0; r < 3; r++) x[r] -= y[r];
   }
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
chomp_tile_cost_kernel(const __grid_constant__ OcbChompArgs a, const int want_grad)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int tid = threadIdx.x;
   const int TW = a.tile_w, CS = TW + 2, NW = TILE_THREADS / TW;
   const int run = blockIdx.x / a.n_tiles, tile = blockIdx.x - run * a.n_tiles;
   if (a.status[run] != 0) return; /* a failed run stays as it was (chomp.c:651-655) */
   const int P = a.P, m = a.m, n = a.n, nsa = a.nsa;

   const TileLayout lay = tile_layout(a, TW);
   double *sd = reinterpret_cast<double *>(smem_raw);
   double *pos = sd + lay.pos, *jfr = sd + lay.jfr, *slots = sd + lay.slots;
   double *Gp = sd + lay.Gp, *cp = sd + lay.cp, *rad = sd + lay.rad, *gb = sd + lay.gb;
   OcbSdfDev *sdfs = reinterpret_cast<OcbSdfDev *>(smem_raw + lay.sdf);
   int *link = reinterpret_cast<int *>(smem_raw + lay.link);
   const int t_first = 1 + tile * TW;

   for (int e = tid; e < nsa; e += TILE_THREADS)
   {
      rad[e] = __ldg(a.radius + e);
      link[e] = a.spheres[e].link;
   }
   {
      const int words = (int) (sizeof(OcbSdfDev) / 4) * a.nsdf;
      const uint32_t *src = reinterpret_cast<const uint32_t *>(a.sdfs);
      uint32_t *dst = reinterpret_cast<uint32_t *>(sdfs);
      for (int e = tid; e < words; e += TILE_THREADS) dst[e] = src[e];
   }
   if (want_grad)
      for (int e = tid; e < NW * n * TW; e += TILE_THREADS) Gp[e] = 0.0;

   /* ---- phase 1: forward kinematics of the tile's columns (one thread per column) ---- */
   if (tid < CS && t_first - 1 + tid <= P - 1)
   {
      const int c = tid;
      const double *Tr = a.traj + ((size_t) run * P + (t_first - 1 + c)) * n;
      double R[9], tr[3], ax[3], org[3];
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = 0.0;
      tr[0] = tr[1] = tr[2] = 0.0;
      for (int j = 0; j < a.nj; j++)
      {
         const OcbJointDev &J = a.joints[j];
         if (a.floating) fk_step<true, true>(J, Tr[J.dof], slots, CS, c, R, tr, ax, org, Tr, 1);
         else fk_step<true>(J, Tr[J.dof], slots, CS, c, R, tr, ax, org);
         if (c >= 1 && c <= TW)
         {
            double *fr = jfr + 6 * j * TW + (c - 1);
#pragma unroll
            for (int k = 0; k < 3; k++) { fr[k * TW] = ax[k]; fr[(3 + k) * TW] = org[k]; }
         }
         {
            const double bx = __ldg(a.gbound + 4 * j), by = __ldg(a.gbound + 4 * j + 1), bz = __ldg(a.gbound + 4 * j + 2);
            double *o = gb + 3 * j * CS + c;
            o[0] = R[0] * bx + R[1] * by + R[2] * bz + tr[0];
            o[CS] = R[3] * bx + R[4] * by + R[5] * bz + tr[1];
            o[2 * CS] = R[6] * bx + R[7] * by + R[8] * bz + tr[2];
         }
         for (int s = J.sph_begin; s < J.sph_end; s++)
         {
            const double px = __ldg(&a.spheres[s].pos[0]), py = __ldg(&a.spheres[s].pos[1]),
                         pz = __ldg(&a.spheres[s].pos[2]);
            double *o = pos + 3 * s * CS + c;
            o[0] = R[0] * px + R[1] * py + R[2] * pz + tr[0];
            o[CS] = R[3] * px + R[4] * py + R[5] * pz + tr[1];
            o[2 * CS] = R[6] * px + R[7] * py + R[8] * pz + tr[2];
         }
      }
   }
   __syncthreads();

   /* ---- phase 2: a worker's lanes are the tile's waypoints ---- */
   const int worker = tid / TW, wp = tid - worker * TW;
   const bool lane_valid = t_first + wp <= m;
   const int wpc = lane_valid ? wp : (m - t_first); /* idle tail lanes shadow the last waypoint */
   const int c = wpc + 1;
   const double *pcol = pos + c;
   const int n_quads = (nsa + 3) >> 2;
   double *Gw = Gp + worker * n * TW + wp;
   const double inv2dt = 1.0 / (2.0 * a.dt), invdt2 = 1.0 / (a.dt * a.dt);
   const double es = a.eps_self;
   const PairConst K = {es, 1.0 / es, 0.5 / es, inv2dt, a.obs_factor_self};
   const int row = a.NAp + a.nsi;
   double cost = 0.0;

   /* J^T of a wrench (F, M about the world origin) acting on joint frame group g, through the
    * stored axes of the joints above it:  c0 axis . (M - org x F), or c0 axis . F (prismatic) */
   auto flush_group = [&](int g, const double F[3], const double M[3])
   {
      const int *ga = a.ganc;
      for (int e = __ldg(ga + g); e < __ldg(ga + g + 1); e++)
      {
         const int jj = __ldg(ga + e);
         const OcbJointDev &J = a.joints[jj];
         const double *fr = jfr + 6 * jj * TW + wpc;
         const double ax0 = fr[0], ax1 = fr[TW], ax2 = fr[2 * TW];
         double val;
         if (J.type == OCB_JOINT_REVOLUTE)
         {
            const double o0 = fr[3 * TW], o1 = fr[4 * TW], o2 = fr[5 * TW];
            const double mx = M[0] - (o1 * F[2] - o2 * F[1]);
            const double my = M[1] - (o2 * F[0] - o0 * F[2]);
            const double mz = M[2] - (o0 * F[1] - o1 * F[0]);
            val = ax0 * mx + ax1 * my + ax2 * mz;
         }
         else
            val = ax0 * F[0] + ax1 * F[1] + ax2 * F[2];
         Gw[J.dof * TW] = fma(J.c0, val, Gw[J.dof * TW]);
      }
      /* floating base: the pose entries see every wrench (pose_gradient is linear in it) */
      if (a.floating) pose_gradient(a.traj + ((size_t) run * P + (t_first + wpc)) * n, 1, F, M, Gw, TW);
   };

   int gcur = -1;
   double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
   /* own spheres are dealt to the workers four at a time, back and forth over the workers, so
    * that every worker sees every part of the robot (in-range pairs cluster on neighbouring
    * links) and long and short partner sweeps alike */
   /* full rounds hand every worker a quad; what is left after them (fewer than 4 NW spheres, the ones with the
    * shortest partner sweeps but the same obstacle work) is split evenly, 1..4 spheres per worker, so that no
    * worker carries a whole extra quad while the others wait at the barrier */
   const int full_rounds = n_quads / NW;
   const int rem_base = full_rounds * NW * 4, rem = nsa - rem_base;
   const int share = (rem + NW - 1) / NW;
   for (int round = 0; round < full_rounds + (rem > 0 ? 1 : 0); round++)
   {
      const int widx = (round & 1) ? NW - 1 - worker : worker;
      const int s0 = (round < full_rounds) ? ((round * NW + widx) << 2) : rem_base + widx * share;
      const int se = min((round < full_rounds) ? s0 + 4 : s0 + share, nsa);
      if (s0 >= se) continue;
      /* (up to) four own spheres share every partner load of the range sweep */
      int sk[4], lk[4];
      double p[4][3], rk[4], f[4][3], cs[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
         sk[k] = min(s0 + k, se - 1);
         lk[k] = (s0 + k < se) ? link[sk[k]] : -1; /* a padded slot shadows the last sphere */
         rk[k] = rad[sk[k]] + es;
         const double *ps = pcol + 3 * sk[k] * CS;
         p[k][0] = ps[0]; p[k][1] = ps[CS]; p[k][2] = ps[2 * CS];
         f[k][0] = f[k][1] = f[k][2] = 0.0;
         cs[k] = 0.0;
      }
      /* partners above the own spheres, one joint frame's range at a time */
      for (int j = 0; j < a.nj; j++)
      {
         const int ge = a.joints[j].sph_end;
         const int o_lo = max(a.joints[j].sph_begin, s0 + 1);
         if (o_lo >= ge) continue;
         {
            /* none of the four own spheres can reach anything this frame carries: skip its sweep
             * (decided by the lanes that are here together; a lane that could reach keeps them all) */
            const double Rj = __ldg(a.gbound + 4 * j + 3);
            const double *gc = gb + 3 * j * CS + c;
            const double g0 = gc[0], g1 = gc[CS], g2 = gc[2 * CS];
            bool near = false;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
               const double dx = p[k][0] - g0, dy = p[k][1] - g1, dz = p[k][2] - g2;
               const double lim = rk[k] + Rj;
               near |= (dx * dx + dy * dy + dz * dz <= lim * lim);
            }
            if (!__any_sync(__activemask(), near)) continue;
         }
         double Rf[3] = {0.0, 0.0, 0.0}, Rm[3] = {0.0, 0.0, 0.0}; /* reaction on this frame */
         bool any = false;
         for (int ob = o_lo; ob < ge; ob += 32)
         {
            const int oe = min(32, ge - ob);
            unsigned mk[4] = {0u, 0u, 0u, 0u};
            for (int i = 0; i < oe; i++)
            {
               const int o = ob + i;
               const double *po = pcol + 3 * o * CS;
               const double q0 = po[0], q1 = po[CS], q2 = po[2 * CS];
               const double ro = rad[o];
               const int lo = link[o];
#pragma unroll
               for (int k = 0; k < 4; k++)
               {
                  const double dx = p[k][0] - q0, dy = p[k][1] - q1, dz = p[k][2] - q2;
                  const double d2 = dx * dx + dy * dy + dz * dz;
                  const double cut = rk[k] + ro; /* r_s + r_o + epsilon_self, mod.cpp:1268 */
                  if (d2 <= cut * cut && lo != lk[k] && o > sk[k]) mk[k] |= 1u << i;
               }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
               unsigned mm = (lk[k] >= 0) ? mk[k] : 0u;
               if (mm == 0u) continue;
               any = true;
               SphereState S;
               sphere_state(pcol, CS, sk[k], inv2dt, invdt2, S);
               while (mm)
               {
                  const int o = ob + __ffs(mm) - 1;
                  mm &= mm - 1;
                  const double *po = pcol + 3 * o * CS;
                  const double q[3] = {po[0], po[CS], po[2 * CS]};
                  double x[3];
                  self_pair_term(K, S, q, po, CS, rk[k] - es + rad[o], want_grad != 0, cs[k], x);
                  f[k][0] += x[0]; f[k][1] += x[1]; f[k][2] += x[2];
                  Rf[0] -= x[0]; Rf[1] -= x[1]; Rf[2] -= x[2];
                  Rm[0] -= q[1] * x[2] - q[2] * x[1];
                  Rm[1] -= q[2] * x[0] - q[0] * x[2];
                  Rm[2] -= q[0] * x[1] - q[1] * x[0];
               }
            }
         }
         if (any && want_grad) flush_group(a.spheres[o_lo].group, Rf, Rm);
      }
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
         if (lk[k] < 0) continue;
         const int s = sk[k];
         SphereState S;
         sphere_state(pcol, CS, s, inv2dt, invdt2, S);
         const double radius = rk[k] - es;
         /* inactive partners are frozen in the world (mod.cpp:2332-2345) */
         const double *crow = a.cut2 + (size_t) s * row + a.NAp;
         for (int i = 0; i < a.nsi; i++)
         {
            const double q[3] = {__ldg(a.inactive_pos + 3 * i), __ldg(a.inactive_pos + 3 * i + 1),
                                 __ldg(a.inactive_pos + 3 * i + 2)};
            const double dx = S.p[0] - q[0], dy = S.p[1] - q[1], dz = S.p[2] - q[2];
            if (dx * dx + dy * dy + dz * dz <= __ldg(crow + i))
            {
               double x[3];
               self_pair_term(K, S, q, nullptr, CS, radius + __ldg(a.radius + nsa + i), want_grad != 0, cs[k], x);
               f[k][0] += x[0]; f[k][1] += x[1]; f[k][2] += x[2];
            }
         }
         double co = 0.0, fo[3] = {0.0, 0.0, 0.0};
         obstacle_term(a, sdfs, S.p, S.vel, S.acc, S.vn, S.iv2, S.moving, radius, want_grad != 0, co, fo);
         cost += co + cs[k];
         if (want_grad)
         {
            const int g = a.spheres[s].group;
            if (g != gcur)
            {
               if (gcur >= 0) flush_group(gcur, F, M);
               gcur = g;
               F[0] = F[1] = F[2] = 0.0;
               M[0] = M[1] = M[2] = 0.0;
            }
            const double fx = fo[0] + f[k][0], fy = fo[1] + f[k][1], fz = fo[2] + f[k][2];
            F[0] += fx; F[1] += fy; F[2] += fz;
            M[0] += S.p[1] * fz - S.p[2] * fy;
            M[1] += S.p[2] * fx - S.p[0] * fz;
            M[2] += S.p[0] * fy - S.p[1] * fx;
         }
      }
   }
   if (want_grad && gcur >= 0) flush_group(gcur, F, M);
   cp[tid] = lane_valid ? cost : 0.0;
   __syncthreads();

   /* ---- fixed-order reduction over the workers ---- */
   if (want_grad)
      for (int e = tid; e < n * TW; e += TILE_THREADS)
      {
         const int j = e / TW, w = e - j * TW, t = t_first + w;
         if (t > m) continue;
         double acc = 0.0;
         for (int wk = 0; wk < NW; wk++) acc += Gp[(wk * n + j) * TW + w];
         a.G_obs[((size_t) run * m + (t - 1)) * n + j] = acc;
      }
   if (tid < 32)
   {
      double acc = 0.0;
      for (int i = tid; i < TILE_THREADS; i += 32) acc += cp[i];
      acc = warp_sum(acc);
      if (tid == 0) a.tile_cost[(size_t) run * a.n_tiles + tile] = acc;
   }
}

/* smem of the per-run update kernel: T, G, (AG) as [n][Ppad] + reduction scratch + MT state */
struct RunLayout
{
   int T, G, AG, red; /* doubles */
   int mt, ired;      /* bytes */
   int bytes;
};

__host__ __device__ inline RunLayout run_layout(const OcbChompArgs &a)
{
   RunLayout l;
   int d = 0;
   l.T = d; d += a.n * a.Ppad;
   l.G = d; d += a.n * a.Ppad;
   l.AG = d; d += a.use_momentum ? a.n * a.Ppad : 0;
   l.red = d; d += 36;
   int b = d * 8;
   l.mt = b; b += a.use_hmc ? (626 + 626 + 16) * 4 : 0;
   l.ired = b; b += 40 * 4;
   l.bytes = b;
   return l;
}

/* One CHOMP update of one run from the obstacle gradient / cost partials of the tile kernel
 * (iter < n_iter), or the final cost evaluation (final_pass). */
__global__ void __launch_bounds__(256, 1)
chomp_run_update_kernel(const __grid_constant__ OcbChompArgs a, const int iter, const int final_pass)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   const int tid = threadIdx.x, NT = blockDim.x;
   const int run = blockIdx.x;
   if (a.status[run] != 0) return;
   const int P = a.P, m = a.m, n = a.n, Pp = a.Ppad;
   const RunLayout lay = run_layout(a);
   double *sd = reinterpret_cast<double *>(smem_raw);
   double *Ts = sd + lay.T, *Gs = sd + lay.G, *AGs = sd + lay.AG, *red = sd + lay.red;
   uint32_t *mts = reinterpret_cast<uint32_t *>(smem_raw + lay.mt);
   int *ired = reinterpret_cast<int *>(smem_raw + lay.ired);

   double *traj = a.traj + (size_t) run * P * n;
   for (int e = tid; e < P * n; e += NT) Ts[(e % n) * Pp + (e / n)] = traj[e];
   const double inv_m = 1.0 / m, inv_lambda = 1.0 / a.lambda;
   double trC;
   {
      double ss = 0.0, sg = 0.0, gg = 0.0;
      for (int j = 0; j < n; j++)
      {
         const double qs = traj[j], qg = traj[(size_t) (P - 1) * n + j];
         ss += qs * qs; sg += qs * qg; gg += qg * qg;
      }
      trC = 0.5 * (a.trc_ss * ss + 2.0 * a.trc_sg * sg + a.trc_gg * gg);
   }
   /* obstacle cost of the trajectory the tile kernel has just looked at */
   double csum = 0.0;
   for (int k = 0; k < a.n_tiles; k++) csum += a.tile_cost[(size_t) run * a.n_tiles + k];
   int red_parity = 0;
   __syncthreads();

   if (final_pass)
   {
      double ssum = 0.0, zero = 0.0;
      for (int t = tid + 1; t <= m; t += NT) ssum += smooth_row(a, Ts, t, Pp, P, n);
      block_sum2(ssum, zero, red, red_parity);
      if (tid == 0)
      {
         a.costs[(size_t) run * 3 + 0] = csum * inv_m + (ssum + trC);
         a.costs[(size_t) run * 3 + 1] = csum * inv_m;
         a.costs[(size_t) run * 3 + 2] = ssum + trC;
      }
      return;
   }

   int leapfrog_first = 0;
   if (a.use_momentum)
   {
      const double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) AGs[(e % n) * Pp + (e / n) + 1] = ag[e];
      leapfrog_first = a.leapfrog_first[run];
   }
   __syncthreads();

   /* ---- HMC momentum resample (mod.cpp:2755-2768) ---- */
   const int ref_iter = a.iter_base + iter; /* the reference's r->iter (mod.cpp:2752) */
   if (a.use_hmc && ref_iter == a.hmc_next[run])
   {
      for (int e = tid; e < 625; e += NT) mts[e] = a.mt_state[(size_t) run * 625 + e];
      const double alpha = 100.0 * exp(0.02 * ref_iter);
      const double sigma = 1.0 / sqrt(alpha);
      uint32_t *saved = mts + 626;
      int *scratch = reinterpret_cast<int *>(mts + 1252);
      __syncthreads();
      for (int e = tid; e < 625; e += NT) saved[e] = mts[e];
      __syncthreads();
      double u = 0.0;
      if (a.use_hmc == 2 || !hmc_resample_parallel(mts, scratch, AGs, Pp, m, n, sigma, &u))
      {
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
      leapfrog_first = 1;
      __syncthreads();
      for (int e = tid; e < 625; e += NT) a.mt_state[(size_t) run * 625 + e] = mts[e];
      if (tid == 0) a.hmc_next[run] = ref_iter + 1 + (int) (-log(u) / a.hmc_lambda);
   }

   /* ---- G = G_obs / m + A T + B (chomp.c:496-517) ---- */
   {
      const double *go = a.G_obs + (size_t) run * m * n;
      for (int t = tid + 1; t <= m; t += NT)
      {
         const double bi = __ldg(a.bcoef_i + t - 1), bf = __ldg(a.bcoef_f + t - 1);
         for (int j = 0; j < n; j++)
         {
            const double *Tj = Ts + j * Pp;
            double g = go[(size_t) (t - 1) * n + j] * inv_m;
            if (a.grad_mode == 2) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
            g += band_AT(a, Tj, t, m) + (bi * Tj[0] + bf * Tj[P - 1]);
            Gs[j * Pp + t] = g;
            if (a.grad_mode == 1) a.grad_out[((size_t) run * m + (t - 1)) * n + j] = g;
         }
      }
   }
   __syncthreads();
   int violated = 0;
   if (a.con_K > 0)
   {
      /* ---- the same with hard constraints (chomp.c:553-600; chomp_constraints.cuh): the branch frames of the
       * constrained waypoints are rebuilt into global scratch, the constraint system lives there too ---- */
      double *slots = a.con_scratch + (size_t) run * a.con_stride + a.con_slots_off;
      violated = a.floating ? con_update<true, true>(a, run, Ts, Gs, AGs, nullptr, slots, red, ired, Pp, m, n, inv_lambda, leapfrog_first)
                            : con_update<false, true>(a, run, Ts, Gs, AGs, nullptr, slots, red, ired, Pp, m, n, inv_lambda, leapfrog_first);
   }
   else
   {
   block_band_solve(a, Gs, Pp, m, n);
   __syncthreads();

   /* ---- momentum / plain update (chomp.c:525-548, 604-605) ---- */
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
   }
   }
   const int any_violation = __syncthreads_or(violated);
   int rounds = 0;
   const bool ok = !any_violation || project_joint_limits(a, Ts, Gs, red, ired, Pp, m, n, rounds);
   if (tid == 0 && rounds > a.limit_rounds[run]) a.limit_rounds[run] = rounds;

   /* the reference has already moved the trajectory when it gives up (chomp.c:651-655) */
   __syncthreads();
   for (int e = tid; e < P * n; e += NT) traj[e] = Ts[(e % n) * Pp + (e / n)];
   if (a.use_momentum)
   {
      double *ag = a.AG + (size_t) run * m * n;
      for (int e = tid; e < m * n; e += NT) ag[e] = AGs[(e % n) * Pp + (e / n) + 1];
      if (tid == 0) a.leapfrog_first[run] = 0;
   }
   if (!ok)
   {
      if (tid == 0) a.status[run] = OCB_ERR_JLIMIT;
      return;
   }

   /* ---- smoothness cost of the updated trajectory (chomp.c:660-671) ---- */
   double ssum = 0.0, zero = 0.0;
   for (int t = tid + 1; t <= m; t += NT) ssum += smooth_row(a, Ts, t, Pp, P, n);
   block_sum2(ssum, zero, red, red_parity);
   if (a.floating)
   {
      /* base quaternions back to unit length, after the costs (mod.cpp:2805-2808) */
      for (int t = tid + 1; t <= m; t += NT)
      {
         pose_normalize(Ts + t, Pp);
         for (int k = 3; k < 7; k++) traj[(size_t) t * n + k] = Ts[k * Pp + t];
      }
   }
   if (tid == 0)
   {
      const double cost_obs = csum * inv_m, cost_smooth = ssum + trC;
      a.costs[(size_t) run * 3 + 0] = cost_obs + cost_smooth;
      a.costs[(size_t) run * 3 + 1] = cost_obs;
      a.costs[(size_t) run * 3 + 2] = cost_smooth;
      a.iters_done[run] = iter + 1;
      if (a.trace_on)
      {
         double *tr = a.trace + ((size_t) run * a.n_iter + iter) * 3;
         tr[0] = cost_obs + cost_smooth;
         tr[1] = cost_obs;
         tr[2] = cost_smooth;
      }
   }
}

} /* namespace */

extern "C" size_t ocb_tile_smem_bytes(const OcbChompArgs *a, int tile_w)
{
   return (size_t) tile_layout(*a, tile_w).bytes;
}

extern "C" size_t ocb_run_update_smem_bytes(const OcbChompArgs *a)
{
   return (size_t) run_layout(*a).bytes;
}

/* one `iterate` call of the tiled path: 2 launches per iteration + 2 for the final cost pass */
extern "C" cudaError_t ocb_launch_chomp_tiled(const OcbChompArgs *args, size_t tile_smem, size_t run_smem,
                                              int run_threads, cudaStream_t st, long *launches)
{
   static OcbSmemOptIn optin_tile, optin_run; /* keyed by device inside */
   cudaError_t e = optin_tile.ensure(chomp_tile_cost_kernel, tile_smem);
   if (e != cudaSuccess) return e;
   e = optin_run.ensure(chomp_run_update_kernel, run_smem);
   if (e != cudaSuccess) return e;
   const int tile_grid = args->R * args->n_tiles;
   for (int iter = 0; iter <= args->n_iter; iter++)
   {
      const int final_pass = (iter == args->n_iter);
      chomp_tile_cost_kernel<<<tile_grid, TILE_THREADS, tile_smem, st>>>(*args, final_pass ? 0 : 1);
      chomp_run_update_kernel<<<args->R, run_threads, run_smem, st>>>(*args, iter, final_pass);
      if (launches) *launches += 2;
   }
   return cudaGetLastError();
}
