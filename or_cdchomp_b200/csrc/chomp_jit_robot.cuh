/* chomp_jit_robot.cuh -- the compiled robot as straight-line code.
 *
 * Only seen by the run-time compiler (ocb_jit.cpp, -DOCB_JIT_ROBOT): ocb_engine.cu writes the
 * kinematic tree that mod::create would pull out of OpenRAVE once per run (mod.cpp:2104-2300) as
 * constexpr tables (ocb_jit_robot.h, generated per robot) and the templates below turn them into
 * code without loops, table loads or index arithmetic:
 *
 *   - the forward sweep over the joint tree (what SetActiveDOFValues + GetTransform +
 *     CalculateJacobian evaluate inside sphere_cost_pre, mod.cpp:1022-1049) with every fixed
 *     transform folded in as literal operands; coefficients that are EXACTLY 0 or +-1 drop out
 *     (x*0 + y*1 + z*0 == y bit for bit; only the sign of a zero can differ), which is most of
 *     them for a robot whose link frames are axis aligned;
 *   - the self-collision range tests (mod.cpp:1251-1268) over the static list of sphere pairs on
 *     different links: one straight-line pass sets one bit per pair in range; pairs rigidly
 *     attached to the same joint frame have a constant distance and are decided once, on the
 *     host, when that distance is not within 1e-9 of the cut-off.
 *
 * Arithmetic per surviving term is that of the generic code in chomp_kernel.cu / chomp_device.cuh.
 * (The generic lambdas below carry no execution-space annotation: NVRTC is run with -default-device.)
 */
#ifndef OCB_CHOMP_JIT_ROBOT_CUH
#define OCB_CHOMP_JIT_ROBOT_CUH

#include "ocb_jit_robot.h" /* generated: JR_* sizes and jr_* tables */

namespace
{

template <int I>
struct JrIC
{
   static constexpr int v = I;
};

/* f(JrIC<B>), f(JrIC<B+1>), ... f(JrIC<E-1>): a loop whose index is a constant expression */
template <int B, int E, class F>
__device__ __forceinline__ void jr_for(F &&f)
{
   if constexpr (B < E)
   {
      f(JrIC<B>{});
      jr_for<B + 1, E>(f);
   }
}

/* r * x for a literal x, exact shortcuts for +-1 */
template <class T>
__device__ __forceinline__ double jr_scale(const double r, T)
{
   constexpr double x = T::x;
   if constexpr (x == 1.0) return r;
   else if constexpr (x == -1.0) return -r;
   else return r * x;
}

/* r0 x0 + r1 x1 + r2 x2 (+ add) with literal x: terms with x == 0 vanish; the dense case is the
 * generic code's own expression */
template <class X0, class X1, class X2, bool ADD>
__device__ __forceinline__ double jr_dot3(const double r0, const double r1, const double r2, const double add)
{
   constexpr double x0 = X0::x, x1 = X1::x, x2 = X2::x;
   constexpr bool n0 = (x0 != 0.0), n1 = (x1 != 0.0), n2 = (x2 != 0.0);
   if constexpr (!n0 && !n1 && !n2)
   {
      if constexpr (ADD) return add;
      else return 0.0;
   }
   else
   {
      double acc;
      if constexpr (n0 && n1 && n2) acc = r0 * x0 + r1 * x1 + r2 * x2;
      else if constexpr (n0 && n1) acc = jr_scale(r0, X0{}) + jr_scale(r1, X1{});
      else if constexpr (n0 && n2) acc = jr_scale(r0, X0{}) + jr_scale(r2, X2{});
      else if constexpr (n1 && n2) acc = jr_scale(r1, X1{}) + jr_scale(r2, X2{});
      else if constexpr (n0) acc = jr_scale(r0, X0{});
      else if constexpr (n1) acc = jr_scale(r1, X1{});
      else acc = jr_scale(r2, X2{});
      if constexpr (ADD) return acc + add;
      else return acc;
   }
}

/* literal carriers: entry K of joint J's fixed rotation / translation, coordinate K of sphere S */
template <int J, int K> struct JrXR { static constexpr double x = jr_XR[J][K]; };
template <int J, int K> struct JrXt { static constexpr double x = jr_Xt[J][K]; };
template <int S, int K> struct JrSP { static constexpr double x = jr_sph_pos[S][K]; };

/* One step of the forward sweep for joint J (fk_step of chomp_device.cuh with the tables folded in).
 * sc: this waypoint's sine / cosine of every revolute joint; COMPUTE_SC fills it, else it is read. */
template <int J, bool SAVE, bool FLOAT, bool COMPUTE_SC>
__device__ __forceinline__ void jr_fk_step(const double *__restrict__ Tt, const int Pp, double *__restrict__ slots,
                                           const int t, double R[9], double tr[3], double ax[3], double org[3],
                                           double *sc)
{
   double Rn[9], tn[3];
   constexpr int load = jr_load[J];
   if constexpr (FLOAT && load == OCB_LOAD_BASE)
   {
      const double qx = Tt[3 * Pp], qy = Tt[4 * Pp], qz = Tt[5 * Pp], qw = Tt[6 * Pp];
      const double qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz, qw2 = qw * qw;
      const double qxqy = qx * qy, qxqz = qx * qz, qxqw = qx * qw;
      const double qyqz = qy * qz, qyqw = qy * qw, qzqw = qz * qw;
      Rn[0] = qx2 - qy2 - qz2 + qw2; Rn[1] = 2 * (qxqy - qzqw);      Rn[2] = 2 * (qxqz + qyqw);
      Rn[3] = 2 * (qxqy + qzqw);     Rn[4] = -qx2 + qy2 - qz2 + qw2; Rn[5] = 2 * (qyqz - qxqw);
      Rn[6] = 2 * (qxqz - qyqw);     Rn[7] = 2 * (qyqz + qxqw);      Rn[8] = -qx2 - qy2 + qz2 + qw2;
      tn[0] = Tt[0]; tn[1] = Tt[Pp]; tn[2] = Tt[2 * Pp];
   }
   else if constexpr (load == OCB_LOAD_BASE)
   {
      Rn[0] = jr_XR[J][0]; Rn[1] = jr_XR[J][1]; Rn[2] = jr_XR[J][2];
      Rn[3] = jr_XR[J][3]; Rn[4] = jr_XR[J][4]; Rn[5] = jr_XR[J][5];
      Rn[6] = jr_XR[J][6]; Rn[7] = jr_XR[J][7]; Rn[8] = jr_XR[J][8];
      tn[0] = jr_Xt[J][0]; tn[1] = jr_Xt[J][1]; tn[2] = jr_Xt[J][2];
   }
   else
   {
      if constexpr (load >= 0)
      {
         const double *sl = slots + 12 * load * Pp + t;
#pragma unroll
         for (int k = 0; k < 9; k++) R[k] = sl[k * Pp];
#pragma unroll
         for (int k = 0; k < 3; k++) tr[k] = sl[(9 + k) * Pp];
      }
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
         const double r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
         Rn[3 * r] = jr_dot3<JrXR<J, 0>, JrXR<J, 3>, JrXR<J, 6>, false>(r0, r1, r2, 0.0);
         Rn[3 * r + 1] = jr_dot3<JrXR<J, 1>, JrXR<J, 4>, JrXR<J, 7>, false>(r0, r1, r2, 0.0);
         Rn[3 * r + 2] = jr_dot3<JrXR<J, 2>, JrXR<J, 5>, JrXR<J, 8>, false>(r0, r1, r2, 0.0);
         tn[r] = jr_dot3<JrXt<J, 0>, JrXt<J, 1>, JrXt<J, 2>, true>(r0, r1, r2, tr[r]);
      }
   }
   ax[0] = Rn[2]; ax[1] = Rn[5]; ax[2] = Rn[8];
   org[0] = tn[0]; org[1] = tn[1]; org[2] = tn[2];
   if constexpr (jr_type[J] == OCB_JOINT_REVOLUTE)
   {
      double s, c;
      if constexpr (COMPUTE_SC)
      {
         double v;
         if constexpr (jr_c0[J] == 1.0 && jr_c1[J] == 0.0) v = Tt[jr_dof[J] * Pp];
         else v = fma(jr_c0[J], Tt[jr_dof[J] * Pp], jr_c1[J]);
         sincos(v, &s, &c);
         if (sc) { sc[2 * J] = s; sc[2 * J + 1] = c; }
      }
      else { s = sc[2 * J]; c = sc[2 * J + 1]; }
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
      const double v = fma(jr_c0[J], Tt[jr_dof[J] * Pp], jr_c1[J]);
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
         R[3 * r] = Rn[3 * r];
         R[3 * r + 1] = Rn[3 * r + 1];
         R[3 * r + 2] = Rn[3 * r + 2];
         tr[r] = fma(v, Rn[3 * r + 2], tn[r]);
      }
   }
   if constexpr (SAVE && jr_save[J] >= 0)
   {
      double *sl = slots + 12 * jr_save[J] * Pp + t;
#pragma unroll
      for (int k = 0; k < 9; k++) sl[k * Pp] = R[k];
#pragma unroll
      for (int k = 0; k < 3; k++) sl[(9 + k) * Pp] = tr[k];
   }
}

/* world positions of the active spheres of waypoint t (fk_waypoint of chomp_kernel.cu) */
template <bool FLOAT>
__device__ __forceinline__ void jr_fk_waypoint(const double *__restrict__ Ts, double *__restrict__ ws, const int Pp,
                                               const int t)
{
   double *slots = ws + 3 * JR_NSA * Pp;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   jr_for<0, JR_NJ>([&](auto jc)
   {
      constexpr int J = decltype(jc)::v;
      jr_fk_step<J, true, FLOAT, true>(Ts + t, Pp, slots, t, R, tr, ax, org, nullptr);
      jr_for<jr_sph_begin[J], jr_sph_end[J]>([&](auto sc)
      {
         constexpr int S = decltype(sc)::v;
         double *o = ws + 3 * S * Pp + t;
         o[0] = jr_dot3<JrSP<S, 0>, JrSP<S, 1>, JrSP<S, 2>, true>(R[0], R[1], R[2], tr[0]);
         o[Pp] = jr_dot3<JrSP<S, 0>, JrSP<S, 1>, JrSP<S, 2>, true>(R[3], R[4], R[5], tr[1]);
         o[2 * Pp] = jr_dot3<JrSP<S, 0>, JrSP<S, 1>, JrSP<S, 2>, true>(R[6], R[7], R[8], tr[2]);
      });
   });
}

/* J^T f of waypoint t from the per-joint-frame wrenches (flush_wrenches of chomp_kernel.cu) */
template <bool FLOAT>
__device__ __forceinline__ void jr_flush_wrenches(const double *__restrict__ Ts, double *__restrict__ ws,
                                                  double *__restrict__ Gs, const int Pp, const int t)
{
   double *slots = ws + 3 * JR_NSA * Pp;
   const double *Wg = ws + (3 * JR_NSA + 12 * JR_NSLOTS) * Pp + t;
   double R[9], tr[3], ax[3], org[3];
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   jr_for<0, JR_NJ>([&](auto jc)
   {
      constexpr int J = decltype(jc)::v;
      jr_fk_step<J, false, FLOAT, true>(Ts + t, Pp, slots, t, R, tr, ax, org, nullptr);
      double F0 = 0.0, F1 = 0.0, F2 = 0.0, M0 = 0.0, M1 = 0.0, M2 = 0.0;
      jr_for<jr_desc_begin[J], jr_desc_end[J]>([&](auto dc)
      {
         constexpr int g = jr_desc[decltype(dc)::v];
         const double *Wo = Wg + 6 * g * Pp;
         F0 += Wo[0]; F1 += Wo[Pp]; F2 += Wo[2 * Pp];
         M0 += Wo[3 * Pp]; M1 += Wo[4 * Pp]; M2 += Wo[5 * Pp];
      });
      double val;
      if constexpr (jr_type[J] == OCB_JOINT_REVOLUTE)
      {
         const double mx = M0 - (org[1] * F2 - org[2] * F1);
         const double my = M1 - (org[2] * F0 - org[0] * F2);
         const double mz = M2 - (org[0] * F1 - org[1] * F0);
         val = ax[0] * mx + ax[1] * my + ax[2] * mz;
      }
      else
         val = ax[0] * F0 + ax[1] * F1 + ax[2] * F2;
      Gs[jr_dof[J] * Pp + t] = fma(jr_c0[J], val, Gs[jr_dof[J] * Pp + t]);
   });
   if constexpr (FLOAT)
   {
      double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int g = 0; g < JR_NG; g++)
      {
         const double *Wo = Wg + 6 * g * Pp;
         F[0] += Wo[0]; F[1] += Wo[Pp]; F[2] += Wo[2 * Pp];
         M[0] += Wo[3 * Pp]; M[1] += Wo[4 * Pp]; M[2] += Wo[5 * Pp];
      }
      pose_gradient(Ts + t, Pp, F, M, Gs + t, Pp);
   }
}

/* Self-collision range tests of waypoint t (mod.cpp:1260-1268): bit k of the result is set when
 * pair k of the static list is within its cut-off.  Pairs are numbered own sphere ascending, then
 * partner ascending: active partners first (JR_NPA pairs, each unordered pair once), then the
 * (own, inactive) pairs.  jr_pair_kind: 0 = test, 1 = always in range, 2 = never (rigid pairs
 * decided on the host).  Own-sphere coordinates are re-used across that sphere's partners. */
struct JrHits
{
   unsigned w[JR_HIT_WORDS];
};

__device__ __forceinline__ JrHits jr_pair_hits(const double *__restrict__ ws, const int Pp, const int t)
{
   JrHits h;
#pragma unroll
   for (int k = 0; k < JR_HIT_WORDS; k++) h.w[k] = jr_hits_always[k];
   jr_for<0, JR_NSA>([&](auto sc)
   {
      constexpr int S = decltype(sc)::v;
      if constexpr (jr_own_tests[S] > 0)
      {
         const double *ps = ws + 3 * S * Pp + t;
         const double px = ps[0], py = ps[Pp], pz = ps[2 * Pp];
         jr_for<jr_pair_begin[S], jr_pair_begin[S + 1]>([&](auto kc)
         {
            constexpr int K = decltype(kc)::v;
            if constexpr (jr_pair_kind[K] == 0)
            {
               const double *po = ws + 3 * jr_pair_o[K] * Pp + t;
               const double dx = px - po[0], dy = py - po[Pp], dz = pz - po[2 * Pp];
               const double d2 = dx * dx + dy * dy + dz * dz;
               if (d2 <= jr_pair_cut2[K]) h.w[K >> 5] |= 1u << (K & 31);
            }
         });
         jr_for<0, JR_NSI>([&](auto ic)
         {
            constexpr int I = decltype(ic)::v;
            constexpr int K = JR_NPA + S * JR_NSI + I;
            if constexpr (jr_pair_kind[K] == 0)
            {
               const double dx = px - jr_inactive_pos[I][0], dy = py - jr_inactive_pos[I][1], dz = pz - jr_inactive_pos[I][2];
               if (dx * dx + dy * dy + dz * dz <= jr_pair_cut2[K]) h.w[K >> 5] |= 1u << (K & 31);
            }
         });
      }
   });
   return h;
}

/* bits [begin, begin + count) of the hit set, count <= 32, begin a run-time value */
__device__ __forceinline__ unsigned jr_hit_bits(const JrHits &h, const int begin, const int count)
{
   const int wi = begin >> 5, sh = begin & 31;
   unsigned lo = 0, hi = 0;
#pragma unroll
   for (int k = 0; k < JR_HIT_WORDS; k++)
   {
      if (k == wi) lo = h.w[k];
      if (k == wi + 1) hi = h.w[k];
   }
   const unsigned v = __funnelshift_r(lo, hi, sh);
   return (count >= 32) ? v : (v & ((1u << count) - 1u));
}

/* the same with compile-time bounds: a couple of shifts */
template <int BEGIN, int COUNT>
__device__ __forceinline__ unsigned jr_hit_bits_static(const JrHits &h)
{
   if constexpr (COUNT <= 0) return 0u;
   else
   {
      constexpr int wi = BEGIN >> 5, sh = BEGIN & 31;
      unsigned v = h.w[wi] >> sh;
      if constexpr (sh + COUNT > 32) v |= h.w[wi + 1] << (32 - sh);
      if constexpr (COUNT < 32) v &= (1u << COUNT) - 1u;
      return v;
   }
}

/* Spheres that may lie inside some field: bit S is set unless sphere S is outside every grid beyond
 * any rounding (the first, conservative half of sdf_sample's range test, for all spheres at once in
 * straight-line code).  Only those spheres are probed. */
template <int NSDF>
__device__ __forceinline__ unsigned jr_grid_candidates(const OcbSdfDev *__restrict__ sdfs, const double *__restrict__ ws,
                                                       const int Pp, const int t)
{
   unsigned cand = 0;
   jr_for<0, JR_NSA>([&](auto sc)
   {
      constexpr int S = decltype(sc)::v;
      const double *ps = ws + 3 * S * Pp + t;
      const double px = ps[0], py = ps[Pp], pz = ps[2 * Pp];
      bool any = false;
#pragma unroll
      for (int k = 0; k < NSDF; k++)
      {
         const OcbSdfDev &G = sdfs[k];
         bool in = true;
#pragma unroll
         for (int r = 0; r < 3; r++)
         {
            const double g = G.Rgw[3 * r] * px + G.Rgw[3 * r + 1] * py + G.Rgw[3 * r + 2] * pz + G.tgw[r];
            in = in && (g >= -1e-290) && !(g * G.scale[r] > G.edge_hi[r]);
         }
         any = any || in;
      }
      if (any) cand |= 1u << S;
   });
   return cand;
}

/* cost (and, when want_grad, the configuration-space gradient row) of moving waypoint t: waypoint_cost
 * of chomp_kernel.cu (sphere_cost, mod.cpp:1134-1327) for the compiled robot.
 *   1. all self-collision range tests in one straight-line pass -> one bit per pair in range;
 *   2. a straight-line pass marks the spheres that may lie inside a field;
 *   3. only spheres with a partner in range or a field around them are visited, each lane walking
 *      its OWN list (a lane never idles through a sphere only its neighbours need); a sphere with
 *      neither has exactly zero cost and force (every term carries a factor that vanishes).
 * Forces are gathered as wrenches per joint frame, in registers while consecutive spheres share a frame. */
template <bool FLOAT>
__device__ __forceinline__ double jr_waypoint_cost(const OcbChompArgs &a, const OcbSdfDev *__restrict__ sdfs,
                                                   const double *__restrict__ Ts, double *__restrict__ ws,
                                                   double *__restrict__ Gs, const int Pp, const int t,
                                                   const bool want_grad PHASE_ARG)
{
   double *Wg = ws + (3 * JR_NSA + 12 * JR_NSLOTS) * Pp + t;
   const double inv2dt = 1.0 / (2.0 * a.dt);
   const double invdt2 = 1.0 / (a.dt * a.dt);
   const double es = a.eps_self, inv_es = 1.0 / es, half_inv_es = 0.5 / es;
   double cost = 0.0;

   if (want_grad)
   {
#pragma unroll
      for (int k = 0; k < 6 * JR_NG; k++) Wg[k * Pp] = 0.0;
   }
   const JrHits hits = jr_pair_hits(ws, Pp, t);
   const unsigned cand = jr_grid_candidates<OCB_JIT_nsdf>(sdfs, ws, Pp, t);
   /* spheres that take part in a pair in range, as the pair's first member (partners are reached from it) */
   unsigned need = cand;
   jr_for<0, JR_NSA>([&](auto sc)
   {
      constexpr int S = decltype(sc)::v;
      const unsigned b = jr_hit_bits_static<jr_pair_begin[S], jr_pair_begin[S + 1] - jr_pair_begin[S]>(hits) |
                         jr_hit_bits_static<JR_NPA + S * JR_NSI, JR_NSI>(hits);
      if (b) need |= 1u << S;
   });
   PHASE(2);

   int gcur = -1;
   double F[3] = {0.0, 0.0, 0.0}, M[3] = {0.0, 0.0, 0.0};
   auto flush_group = [&]()
   {
      if (gcur >= 0 && want_grad)
      {
         double *Wo = Wg + 6 * gcur * Pp;
         Wo[0] += F[0]; Wo[Pp] += F[1]; Wo[2 * Pp] += F[2];
         Wo[3 * Pp] += M[0]; Wo[4 * Pp] += M[1]; Wo[5 * Pp] += M[2];
      }
      F[0] = F[1] = F[2] = M[0] = M[1] = M[2] = 0.0;
   };
   while (need)
   {
      const int s = __ffs(need) - 1;
      need &= need - 1;
      const double *ps = ws + 3 * s * Pp + t;
      const double p[3] = {ps[0], ps[Pp], ps[2 * Pp]};
      const double radius = jr_radius[s];

      /* --- obstacle term, first step: which field, how far (mod.cpp:1169-1198) --- */
      double d_obs = 0.0, bg[3] = {0.0, 0.0, 0.0};
      int best = -1;
      if ((cand >> s) & 1u) best = obstacle_probe(a, sdfs, p, radius, OCB_JIT_nsdf, d_obs, bg);
      const bool obs = (best >= 0) && (d_obs < a.eps);
      const int pb = jr_pair_begin[s];
      unsigned hit_a = jr_hit_bits(hits, pb, jr_pair_begin[s + 1] - pb);
      unsigned hit_i = (JR_NSI > 0) ? jr_hit_bits(hits, JR_NPA + s * JR_NSI, JR_NSI) : 0u;
      if (!obs && !(hit_a | hit_i)) continue;
      const int g = jr_group[s];
      if (g != gcur)
      {
         flush_group();
         gcur = g;
      }

      double vel[3];
#pragma unroll
      for (int k = 0; k < 3; k++) vel[k] = (ps[k * Pp + 1] - ps[k * Pp - 1]) * inv2dt;
      const double vn2 = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2];
      const double rv = rsqrt(vn2);
      const double vn = (vn2 > 0.0) ? vn2 * rv : 0.0;
      const double iv2 = rv * rv; /* 1 / |v|^2, unguarded (inf at rest), as mod.cpp:1239 */
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
            for (int k = 0; k < 3; k++) acc[k] = (p[k] * -2.0 + ps[k * Pp - 1] + ps[k * Pp + 1]) * invdt2;
         }
         obstacle_apply(a, sdfs[best], d_obs, bg, vel, acc, vn, iv2, moving, want_grad, cost_s, f);
      }

      /* --- self collision, each unordered pair once (1251-1317); see waypoint_cost --- */
      const double ws_self = vn * a.obs_factor_self;
      auto in_range = [&](const double q[3], const double *po, const int o)
      {
         const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
         const double d2 = dx * dx + dy * dy + dz * dz;
         const double inv = rsqrt(d2);
         const double dist = d2 * inv;
         const double dd = dist - (radius + jr_radius[o]);
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
#pragma unroll
            for (int r = 0; r < 3; r++) x[r] -= y[r];
            const int go = jr_group[o];
            if (go == gcur)
            {
               /* partner on the same joint frame: its reaction joins the wrench held in registers */
               F[0] -= x[0]; F[1] -= x[1]; F[2] -= x[2];
               M[0] -= q[1] * x[2] - q[2] * x[1];
               M[1] -= q[2] * x[0] - q[0] * x[2];
               M[2] -= q[0] * x[1] - q[1] * x[0];
            }
            else
            {
               double *Wo = Wg + 6 * go * Pp;
               Wo[0] -= x[0];
               Wo[Pp] -= x[1];
               Wo[2 * Pp] -= x[2];
               Wo[3 * Pp] -= q[1] * x[2] - q[2] * x[1];
               Wo[4 * Pp] -= q[2] * x[0] - q[0] * x[2];
               Wo[5 * Pp] -= q[0] * x[1] - q[1] * x[0];
            }
         }
#pragma unroll
         for (int r = 0; r < 3; r++) f[r] += x[r];
      };
      /* active partners in ascending order, then the inactive ones (frozen in the world, mod.cpp:2332-2345) */
      while (hit_a)
      {
         const int k = __ffs(hit_a) - 1;
         hit_a &= hit_a - 1;
         const int o = jr_pair_o[pb + k];
         const double *po = ws + 3 * o * Pp + t;
         const double q[3] = {po[0], po[Pp], po[2 * Pp]};
         in_range(q, po, o);
      }
      while (hit_i)
      {
         const int i = __ffs(hit_i) - 1;
         hit_i &= hit_i - 1;
         const double q[3] = {jr_inactive_pos[i][0], jr_inactive_pos[i][1], jr_inactive_pos[i][2]};
         in_range(q, nullptr, JR_NSA + i);
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
   flush_group();
   PHASE(3);
   if (want_grad) jr_flush_wrenches<FLOAT>(Ts, ws, Gs, Pp, t);
   PHASE(4);
   return cost;
}

} /* namespace */

#endif
