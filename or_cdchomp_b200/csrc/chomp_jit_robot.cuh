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

/* ---- branch-free elementary functions ---------------------------------------------------------
 * CUDA's rsqrt() and sincos() each end in a branch to a special-case path (denormals, huge arguments);
 * fast_rsqrt (chomp_device.cuh) and jr_sincos below do without.
 * A branch closes the scheduler's window: the three square roots of a sphere pair, or the seven joint
 * angles of a sweep, are then evaluated one after the other although they do not depend on each
 * other, and a warp of this kernel has little else to overlap them with.  These versions are the
 * same algorithms without the branch; the callers keep their arguments in range. */

/* sin and cos of |x| <= 1e4: Cody-Waite reduction by pi/2 in three pieces, the fdlibm kernels
 * (k_sin.c, k_cos.c) on [-pi/4, pi/4], quadrant by selection; error below 1 ulp */
__device__ __forceinline__ void jr_sincos(const double x, double &s, double &c)
{
   const double q = rint(x * 0x1.45f306dc9c883p-1);
   const int iq = (int) q;
   double r = fma(-q, 0x1.921fb54442d18p+0, x);
   r = fma(-q, 0x1.1a62633145c07p-54, r);
   r = fma(-q, -0x1.f1976b7ed8fbcp-110, r);
   const double z = r * r;
   double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
   ps = fma(z, ps, 2.75573137070700676789e-06);
   ps = fma(z, ps, -1.98412698298579493134e-04);
   ps = fma(z, ps, 8.33333333332248946124e-03);
   ps = fma(z, ps, -1.66666666666666324348e-01);
   const double sn = fma(r * z, ps, r);
   double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
   pc = fma(z, pc, -2.75573143513906633035e-07);
   pc = fma(z, pc, 2.48015872894767294178e-05);
   pc = fma(z, pc, -1.38888888888741095749e-03);
   pc = fma(z, pc, 4.16666666666666019037e-02);
   const double hz = 0.5 * z, w = 1.0 - hz;
   const double cs = w + (((1.0 - w) - hz) + z * (z * pc));
   const bool swap = iq & 1;
   const double a0 = swap ? cs : sn, b0 = swap ? sn : cs;
   s = (iq & 2) ? -a0 : a0;
   c = ((iq + 1) & 2) ? -b0 : b0;
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
 * sc: this waypoint's sine / cosine of every joint angle (jr_trig). */
template <int J, bool SAVE, bool FLOAT>
__device__ __forceinline__ void jr_fk_step(const double *__restrict__ Tt, const int Pp, double *__restrict__ slots,
                                           const int t, double R[9], double tr[3], double ax[3], double org[3],
                                           const double *sc)
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
      const double s = sc[2 * J], c = sc[2 * J + 1];
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

/* sine and cosine of every revolute joint's angle at this waypoint, all at once: seven independent
 * evaluations in one straight line; an angle beyond the range of jr_sincos (never, for a joint with
 * limits) is redone with sincos() afterwards */
__device__ __forceinline__ void jr_trig(const double *__restrict__ Tt, const int Pp, double sc[2 * JR_NJ])
{
   bool big = false;
   jr_for<0, JR_NJ>([&](auto jc)
   {
      constexpr int J = decltype(jc)::v;
      if constexpr (jr_type[J] == OCB_JOINT_REVOLUTE)
      {
         double v;
         if constexpr (jr_c0[J] == 1.0 && jr_c1[J] == 0.0) v = Tt[jr_dof[J] * Pp];
         else v = fma(jr_c0[J], Tt[jr_dof[J] * Pp], jr_c1[J]);
         big = big || !(fabs(v) <= 1e4);
         jr_sincos(v, sc[2 * J], sc[2 * J + 1]);
      }
      else
      {
         sc[2 * J] = 0.0;
         sc[2 * J + 1] = 1.0;
      }
   });
   if (big)
   {
      jr_for<0, JR_NJ>([&](auto jc)
      {
         constexpr int J = decltype(jc)::v;
         if constexpr (jr_type[J] == OCB_JOINT_REVOLUTE)
            sincos(fma(jr_c0[J], Tt[jr_dof[J] * Pp], jr_c1[J]), &sc[2 * J], &sc[2 * J + 1]);
      });
   }
}

/* world positions of the active spheres of waypoint t (fk_waypoint of chomp_kernel.cu) */
template <bool FLOAT>
__device__ __forceinline__ void jr_fk_waypoint(const double *__restrict__ Ts, double *__restrict__ ws, const int Pp,
                                               const int t, double *__restrict__ trig)
{
   double *slots = ws + 3 * JR_NSA * Pp;
   double R[9], tr[3], ax[3], org[3], sc[2 * JR_NJ];
   jr_trig(Ts + t, Pp, sc);
#ifndef JR_NO_TRIG_CACHE
   /* the J^T sweep of this iteration needs the same sines and cosines: parked in global memory (L2) */
#pragma unroll
   for (int k = 0; k < 2 * JR_NJ; k++) trig[k * Pp + t] = sc[k];
#endif
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   jr_for<0, JR_NJ>([&](auto jc)
   {
      constexpr int J = decltype(jc)::v;
      jr_fk_step<J, true, FLOAT>(Ts + t, Pp, slots, t, R, tr, ax, org, sc);
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
                                                  double *__restrict__ Gs, const int Pp, const int t,
                                                  const double *__restrict__ trig)
{
   double *slots = ws + 3 * JR_NSA * Pp;
   const double *Wg = ws + (3 * JR_NSA + 12 * JR_NSLOTS) * Pp + t;
   double R[9], tr[3], ax[3], org[3], sc[2 * JR_NJ];
#ifndef JR_NO_TRIG_CACHE
#pragma unroll
   for (int k = 0; k < 2 * JR_NJ; k++) sc[k] = trig[k * Pp + t];
#else
   jr_trig(Ts + t, Pp, sc);
#endif
#pragma unroll
   for (int k = 0; k < 9; k++) R[k] = 0.0;
   tr[0] = tr[1] = tr[2] = 0.0;
   jr_for<0, JR_NJ>([&](auto jc)
   {
      constexpr int J = decltype(jc)::v;
      jr_fk_step<J, false, FLOAT>(Ts + t, Pp, slots, t, R, tr, ax, org, sc);
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
   /* three spheres at a time: independent work for the scheduler without the register pressure of all fifteen */
#pragma unroll 3
   for (int S = 0; S < JR_NSA; S++)
   {
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
   }
   return cand;
}

/* finite-difference velocity of a sphere column (mod.cpp:1099-1113) with its norm, 1/|v|^2 and the
 * "moving" predicate of sphere_cost (1226, 1302) */
struct JrVel
{
   double v[3], vn, iv2;
   bool moving;
};

__device__ __forceinline__ JrVel jr_velocity(const double *__restrict__ ps, const int Pp, const double inv2dt)
{
   JrVel k;
#pragma unroll
   for (int r = 0; r < 3; r++) k.v[r] = (ps[r * Pp + 1] - ps[r * Pp - 1]) * inv2dt;
   const double vn2 = k.v[0] * k.v[0] + k.v[1] * k.v[1] + k.v[2] * k.v[2];
   speed_terms(vn2, k.vn, k.iv2);
   k.moving = k.vn > 0.000001;
   return k;
}

/* cost (and, when WANT_GRAD, the configuration-space gradient row) of moving waypoint t: waypoint_cost
 * of chomp_kernel.cu (sphere_cost, mod.cpp:1134-1327) for the compiled robot, organised for
 * instruction-level parallelism -- a warp of this kernel is one long dependency chain, and three
 * warps per scheduler do not hide it:
 *   1. all self-collision range tests in one straight-line pass -> one bit per pair in range;
 *   2. a straight-line pass marks the spheres that may lie inside a field; only those are probed, and
 *      only a sphere with a field value below epsilon has an obstacle term at all;
 *   3. the pairs in range are then worked off PAIR by pair, every lane walking its own list, each pair a
 *      branch-free block that needs nothing from the previous one: both spheres' velocities, both directed terms
 *      (mod.cpp:1281-1317), the force on one sphere and its reaction on the other added straight to
 *      the wrenches of their joint frames (F, M about the world origin, in shared memory).
 * Sums therefore run in pair order instead of sphere order: results agree with waypoint_cost to rounding. */
template <bool FLOAT, bool WANT_GRAD>
__device__ __forceinline__ double jr_waypoint_cost(const OcbChompArgs &a, const OcbSdfDev *__restrict__ sdfs,
                                                   const double *__restrict__ Ts, double *__restrict__ ws,
                                                   double *__restrict__ Gs, const int Pp, const int t,
                                                   const double *__restrict__ trig PHASE_ARG)
{
   double *Wg = ws + (3 * JR_NSA + 12 * JR_NSLOTS) * Pp + t;
   const double inv2dt = 1.0 / (2.0 * a.dt);
   const double invdt2 = 1.0 / (a.dt * a.dt);
   const double es = a.eps_self, inv_es = 1.0 / es, half_inv_es = 0.5 / es;
   const double ofs = a.obs_factor_self;
   double cost = 0.0;

   if (WANT_GRAD)
   {
#pragma unroll
      for (int k = 0; k < 6 * JR_NG; k++) Wg[k * Pp] = 0.0;
   }
   const JrHits hits = jr_pair_hits(ws, Pp, t);
#ifdef EXP_NOCAND
   unsigned cand = (1u << JR_NSA) - 1u;
#else
   unsigned cand = jr_grid_candidates<OCB_JIT_nsdf>(sdfs, ws, Pp, t);
#endif
   PHASE(2);

   /* force f at point c on joint frame g: its wrench about the world origin */
   auto add_wrench = [&](const int g, const double c[3], const double f[3], const double sign)
   {
      double *Wo = Wg + 6 * g * Pp;
      Wo[0] = fma(sign, f[0], Wo[0]);
      Wo[Pp] = fma(sign, f[1], Wo[Pp]);
      Wo[2 * Pp] = fma(sign, f[2], Wo[2 * Pp]);
      Wo[3 * Pp] = fma(sign, c[1] * f[2] - c[2] * f[1], Wo[3 * Pp]);
      Wo[4 * Pp] = fma(sign, c[2] * f[0] - c[0] * f[2], Wo[4 * Pp]);
      Wo[5 * Pp] = fma(sign, c[0] * f[1] - c[1] * f[0], Wo[5 * Pp]);
   };

   /* --- obstacle terms (mod.cpp:1169-1249) --- */
   while (cand)
   {
      const int s = __ffs(cand) - 1;
      cand &= cand - 1;
      const double *ps = ws + 3 * s * Pp + t;
      const double p[3] = {ps[0], ps[Pp], ps[2 * Pp]};
      double d_obs, bg[3];
      const int best = obstacle_probe(a, sdfs, p, jr_radius[s], OCB_JIT_nsdf, d_obs, bg);
      if (best < 0 || !(d_obs < a.eps)) continue;
      const JrVel k = jr_velocity(ps, Pp, inv2dt);
      double acc[3] = {0.0, 0.0, 0.0};
      if (WANT_GRAD)
      {
#pragma unroll
         for (int r = 0; r < 3; r++) acc[r] = (p[r] * -2.0 + ps[r * Pp - 1] + ps[r * Pp + 1]) * invdt2;
      }
      double cost_s = 0.0, f[3] = {0.0, 0.0, 0.0};
      obstacle_apply(a, sdfs[best], d_obs, bg, k.v, acc, k.vn, k.iv2, k.moving, WANT_GRAD, cost_s, f);
      cost += cost_s;
      if (WANT_GRAD) add_wrench(jr_group[s], p, f, 1.0);
   }

   /* --- self collision (mod.cpp:1251-1317): one pair of active spheres in range (live: a weight of 1) --- */
   auto active_pair = [&](const int kk, const double live)
   {
      const unsigned info = jr_pair_info[kk]; /* s | o << 8 | group(s) << 16 | group(o) << 24 */
      const int s = info & 0xff, o = (info >> 8) & 0xff;
      const double *ps = ws + 3 * s * Pp + t, *po = ws + 3 * o * Pp + t;
      const double p[3] = {ps[0], ps[Pp], ps[2 * Pp]}, q[3] = {po[0], po[Pp], po[2 * Pp]};
      const JrVel ks = jr_velocity(ps, Pp, inv2dt), ko = jr_velocity(po, Pp, inv2dt);
      const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
      const double d2 = dx * dx + dy * dy + dz * dz;
      const double inv = fast_rsqrt(d2); /* two spheres in range of each other on different links: d2 is no denormal */
      const double dist = d2 * inv;
      const double dd = dist - jr_pair_rsum[kk];
      /* cost shape shared by both directions (1281-1289) */
      const double cshape = (dd < 0.0) ? (0.5 * es - dd) : half_inv_es * (dd - es) * (dd - es);
      const double w_s = ks.vn * ofs * live, w_o = ko.vn * ofs * live;
      cost += w_s * cshape;
      cost += w_o * cshape;
      if (WANT_GRAD)
      {
         const double sc = (dd < 0.0) ? -1.0 : ((dd < es) ? (dd * inv_es - 1.0) : 1.0);
         const double gh[3] = {dx * inv, dy * inv, dz * inv};
         double x[3], y[3];
         const double wa = sc * w_s, wb = -sc * w_o;
#pragma unroll
         for (int r = 0; r < 3; r++) { x[r] = gh[r] * wa; y[r] = gh[r] * wb; }
         /* each directed term loses its component along its own sphere's velocity (1302-1306) */
         const double pjx = ks.moving ? (x[0] * ks.v[0] + x[1] * ks.v[1] + x[2] * ks.v[2]) * ks.iv2 : 0.0;
         const double pjy = ko.moving ? (y[0] * ko.v[0] + y[1] * ko.v[1] + y[2] * ko.v[2]) * ko.iv2 : 0.0;
#pragma unroll
         for (int r = 0; r < 3; r++)
         {
            x[r] = fma(-pjx, ks.v[r], x[r]);
            y[r] = fma(-pjy, ko.v[r], y[r]);
            x[r] -= y[r]; /* (J_s - J_o)^T x + (J_o - J_s)^T y = J_s^T (x - y) - J_o^T (x - y) */
         }
         add_wrench((info >> 16) & 0xff, p, x, 1.0);
         add_wrench((info >> 24) & 0xff, q, x, -1.0);
      }
   };
   {
      /* every lane walks its OWN list of pairs in range, in ascending order across the words of the hit
       * set: the loop runs as often as the busiest lane has pairs */
      unsigned bits[JR_HIT_WORDS];
      unsigned any = 0;
#pragma unroll
      for (int w = 0; w < JR_HIT_WORDS; w++)
      {
         const unsigned amask = (32 * w >= JR_NPA) ? 0u : ((JR_NPA - 32 * w >= 32) ? 0xffffffffu : ((1u << ((JR_NPA - 32 * w) & 31)) - 1u));
         bits[w] = hits.w[w] & amask;
         any |= bits[w];
      }
      while (any)
      {
         int kk = 0;
         bool found = false;
         any = 0;
#pragma unroll
         for (int w = 0; w < JR_HIT_WORDS; w++)
         {
            const bool take = !found && bits[w] != 0;
            if (take)
            {
               kk = 32 * w + __ffs(bits[w]) - 1;
               bits[w] &= bits[w] - 1;
            }
            found = found || take;
            any |= bits[w];
         }
         active_pair(kk, 1.0);
      }
   }
   /* an active sphere against an inactive one, frozen in the world (mod.cpp:2332-2345): one directed term */
   if constexpr (JR_NSI > 0)
   {
#pragma unroll
      for (int w = 0; w < JR_HIT_WORDS; w++)
      {
         const unsigned imask = (32 * w + 32 <= JR_NPA) ? 0u : ((32 * w >= JR_NPA) ? 0xffffffffu : ~((1u << ((JR_NPA - 32 * w) & 31)) - 1u));
         unsigned bits = hits.w[w] & imask;
         while (bits)
         {
            const int kk = 32 * w + __ffs(bits) - 1;
            bits &= bits - 1;
            const int s = (kk - JR_NPA) / JR_NSI, i = (kk - JR_NPA) % JR_NSI;
            const double *ps = ws + 3 * s * Pp + t;
            const double p[3] = {ps[0], ps[Pp], ps[2 * Pp]};
            const JrVel ks = jr_velocity(ps, Pp, inv2dt);
            const double dx = p[0] - jr_inactive_pos[i][0], dy = p[1] - jr_inactive_pos[i][1], dz = p[2] - jr_inactive_pos[i][2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            const double inv = fast_rsqrt(d2);
            const double dist = d2 * inv;
            const double dd = dist - (jr_radius[s] + jr_radius[JR_NSA + i]);
            const double cshape = (dd < 0.0) ? (0.5 * es - dd) : half_inv_es * (dd - es) * (dd - es);
            const double w_s = ks.vn * ofs;
            cost += w_s * cshape;
            if (WANT_GRAD)
            {
               const double sc = (dd < 0.0) ? -1.0 : ((dd < es) ? (dd * inv_es - 1.0) : 1.0);
               const double wa = sc * w_s;
               double x[3] = {dx * inv * wa, dy * inv * wa, dz * inv * wa};
               const double pjx = ks.moving ? (x[0] * ks.v[0] + x[1] * ks.v[1] + x[2] * ks.v[2]) * ks.iv2 : 0.0;
#pragma unroll
               for (int r = 0; r < 3; r++) x[r] = fma(-pjx, ks.v[r], x[r]);
               add_wrench(jr_group[s], p, x, 1.0);
            }
         }
      }
   }
   PHASE(3);
   if (WANT_GRAD) jr_flush_wrenches<FLOAT>(Ts, ws, Gs, Pp, t, trig);
   PHASE(4);
   return cost;
}

} /* namespace */

#endif
