/* chomp_constraints.cuh -- hard end-effector constraints (task space regions) of the CHOMP update.
 *
 * What it replaces in the reference (paths relative to the reference root):
 *   con_tsr / con_everyn_tsr / con_start_tsr      src/orcdchomp_mod.cpp:1330-1497, 1500-1657, 1659-1784
 *     (value = chosen entries of [x y z yaw pitch roll] of  inv(T0w) * frame * inv(Twe);  Jacobian =
 *      xyzypr_J * pose_jac_inverse * xm(inv(T0w)) * [angular; linear-at-origin] world Jacobian)
 *   cd_kin_pose_to_xyzypr, cd_kin_pose_to_xyzypr_J src/libcd/kin.c:615-647, 680-718
 *   cd_spatial_xm_from_pose, cd_spatial_pose_jac, cd_spatial_pose_jac_inverse
 *                                                 src/libcd/spatial.c:71-102, 295-337, 339-375
 *   the projection in cd_chomp_iterate            src/libcd/chomp.c:553-600 (dgemv / dgemm / LAPACKE_dgesv)
 *
 * Differences of FORM: the frame's Jacobian is never stored in world coordinates -- every joint column
 * [axis; origin x axis] goes straight through the k x 6 map of its constraint; the three 6 x 6 products of
 * the reference collapse to  lin = R_A v + (t_A - p) x (R_A w),  ang = E (R_A w)  with E the 3 x 3 product of
 * the reference's yaw-pitch-roll and quaternion-rate tables.  With a tridiagonal metric (derivative = 1) the
 * dense system J A^-1 J^T is never formed: the correction comes from a block-tridiagonal sweep over the
 * waypoints (con_project_tridiag).  Wider metrics build the system and solve it by the block with the same
 * row-pivoted elimination LAPACK performs.
 */
#ifndef OCB_CHOMP_CONSTRAINTS_CUH
#define OCB_CHOMP_CONSTRAINTS_CUH

#include "chomp_device.cuh"

namespace
{

/* unit quaternion [x y z w] of a rotation matrix, largest component first (its sign is free: every use
 * below is even in q) */
__device__ inline void con_quat_of(const double R[9], double q[4])
{
   const double tr = R[0] + R[4] + R[8];
   if (tr >= R[0] && tr >= R[4] && tr >= R[8])
   {
      const double w = 0.5 * sqrt(1.0 + tr), s = 0.25 / w;
      q[3] = w; q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
   }
   else if (R[0] >= R[4] && R[0] >= R[8])
   {
      const double x = 0.5 * sqrt(1.0 + R[0] - R[4] - R[8]), s = 0.25 / x;
      q[0] = x; q[1] = (R[1] + R[3]) * s; q[2] = (R[2] + R[6]) * s; q[3] = (R[7] - R[5]) * s;
   }
   else if (R[4] >= R[8])
   {
      const double y = 0.5 * sqrt(1.0 - R[0] + R[4] - R[8]), s = 0.25 / y;
      q[1] = y; q[0] = (R[1] + R[3]) * s; q[2] = (R[5] + R[7]) * s; q[3] = (R[2] - R[6]) * s;
   }
   else
   {
      const double z = 0.5 * sqrt(1.0 - R[0] - R[4] + R[8]), s = 0.25 / z;
      q[2] = z; q[0] = (R[2] + R[6]) * s; q[1] = (R[5] + R[7]) * s; q[3] = (R[3] - R[1]) * s;
   }
}

/* what a constraint needs of its frame at one waypoint: the six candidate values, the point the linear
 * rows pivot about, and the angular map E */
struct ConFrame
{
   double xyzypr[6];
   double arm[3];  /* t_A - p: lin = R_A v + arm x (R_A w) */
   double E[9];    /* [yaw; pitch; roll] rates from the angular velocity in the TSR frame */
};

__device__ inline void con_frame(const OcbConDev &c, const double R[9], const double tr[3], ConFrame &f)
{
   /* frame in the world, then seen from the TSR:  A * (joint frame * C) */
   double Rf[9], tf[3], Rt[9], p[3];
#pragma unroll
   for (int r = 0; r < 3; r++)
   {
#pragma unroll
      for (int k = 0; k < 3; k++) Rf[3 * r + k] = R[3 * r] * c.CR[k] + R[3 * r + 1] * c.CR[3 + k] + R[3 * r + 2] * c.CR[6 + k];
      tf[r] = R[3 * r] * c.Ct[0] + R[3 * r + 1] * c.Ct[1] + R[3 * r + 2] * c.Ct[2] + tr[r];
   }
#pragma unroll
   for (int r = 0; r < 3; r++)
   {
#pragma unroll
      for (int k = 0; k < 3; k++) Rt[3 * r + k] = c.AR[3 * r] * Rf[k] + c.AR[3 * r + 1] * Rf[3 + k] + c.AR[3 * r + 2] * Rf[6 + k];
      p[r] = c.AR[3 * r] * tf[0] + c.AR[3 * r + 1] * tf[1] + c.AR[3 * r + 2] * tf[2] + c.At[r];
   }
   double q[4];
   con_quat_of(Rt, q);
   const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
   /* kin.c:615-647 */
   f.xyzypr[0] = p[0]; f.xyzypr[1] = p[1]; f.xyzypr[2] = p[2];
   const double half_sinp = qw * qy - qz * qx;
   const double quarter_turn = 1.5707963267948966;
   if (half_sinp > 0.49999)
   {
      f.xyzypr[3] = -2.0 * atan2(qx, qw); f.xyzypr[4] = quarter_turn; f.xyzypr[5] = 0.0;
   }
   else if (half_sinp < -0.49999)
   {
      f.xyzypr[3] = 2.0 * atan2(qx, qw); f.xyzypr[4] = -quarter_turn; f.xyzypr[5] = 0.0;
   }
   else
   {
      f.xyzypr[3] = atan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz));
      f.xyzypr[4] = asin(2.0 * half_sinp);
      f.xyzypr[5] = atan2(2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx * qx + qy * qy));
   }
#pragma unroll
   for (int r = 0; r < 3; r++) f.arm[r] = c.At[r] - p[r];
   /* Y = d[yaw pitch roll]/d[qx qy qz qw] (kin.c:694-716, general branch only, as the reference) */
   double Y[12];
   {
      double nu = 2.0 * (qw * qz + qx * qy), de = 1.0 - 2.0 * (qy * qy + qz * qz);
      double dn = de / (de * de + nu * nu), nn = nu / (de * de + nu * nu);
      Y[0] = dn * (2.0 * qy);
      Y[1] = dn * (2.0 * qx) + nn * (4.0 * qy);
      Y[2] = dn * (2.0 * qw) + nn * (4.0 * qz);
      Y[3] = dn * (2.0 * qz);
      const double as = 2.0 * half_sinp;
      const double s = 2.0 / sqrt(1.0 - as * as);
      Y[4] = -s * qz; Y[5] = s * qw; Y[6] = -s * qx; Y[7] = s * qy;
      nu = 2.0 * (qw * qx + qy * qz);
      de = 1.0 - 2.0 * (qx * qx + qy * qy);
      dn = de / (de * de + nu * nu);
      nn = nu / (de * de + nu * nu);
      Y[8] = dn * (2.0 * qw) + nn * (4.0 * qx);
      Y[9] = dn * (2.0 * qz) + nn * (4.0 * qy);
      Y[10] = dn * (2.0 * qy);
      Y[11] = dn * (2.0 * qx);
   }
   /* H = d[qx qy qz qw]/d(omega) (spatial.c:361-373) */
   const double hx = 0.5 * qx, hy = 0.5 * qy, hz = 0.5 * qz, hw = 0.5 * qw;
   const double H[12] = {hw, hz, -hy, -hz, hw, hx, hy, -hx, hw, -hx, -hy, -hz};
#pragma unroll
   for (int r = 0; r < 3; r++)
#pragma unroll
      for (int k = 0; k < 3; k++)
         f.E[3 * r + k] = Y[4 * r] * H[k] + Y[4 * r + 1] * H[3 + k] + Y[4 * r + 2] * H[6 + k] + Y[4 * r + 3] * H[9 + k];
}

/* one world Jacobian column [w; v] through the constraint's map: entries of d[x y z yaw pitch roll] */
__device__ inline void con_map_column(const OcbConDev &c, const ConFrame &f, const double w[3], const double v[3],
                                      double out[6])
{
   double wt[3], vt[3];
#pragma unroll
   for (int r = 0; r < 3; r++)
   {
      wt[r] = c.AR[3 * r] * w[0] + c.AR[3 * r + 1] * w[1] + c.AR[3 * r + 2] * w[2];
      vt[r] = c.AR[3 * r] * v[0] + c.AR[3 * r + 1] * v[1] + c.AR[3 * r + 2] * v[2];
   }
   out[0] = vt[0] + (f.arm[1] * wt[2] - f.arm[2] * wt[1]);
   out[1] = vt[1] + (f.arm[2] * wt[0] - f.arm[0] * wt[2]);
   out[2] = vt[2] + (f.arm[0] * wt[1] - f.arm[1] * wt[0]);
#pragma unroll
   for (int r = 0; r < 3; r++) out[3 + r] = f.E[3 * r] * wt[0] + f.E[3 * r + 1] * wt[1] + f.E[3 * r + 2] * wt[2];
}

__device__ inline bool con_applies(const OcbConDev &c, int i, int m)
{
   return c.where == OCB_CON_ALL || (c.where == OCB_CON_START && i == 0) || (c.where == OCB_CON_END && i == m - 1) ||
          (c.where == OCB_CON_START_TSR && i == 0);
}

/* Values h and Jacobian rows J (row length n) of every constraint on moving waypoint t (column t of the
 * [item][waypoint] arrays), written at the waypoint's rows of the run's stacked system; then
 * h += -1/lambda * J AG_t  (chomp.c:562-565).  The joint frames of waypoint t are those the forward
 * sweep of this iteration left in the workspace (or, SAVE_FRAMES, rebuilt here). */
template <bool FLOAT, bool SAVE_FRAMES>
__device__ inline void con_eval_waypoint(const OcbChompArgs &a, const double *__restrict__ Ts,
                                         double *__restrict__ slots, const double *__restrict__ AGc,
                                         int Pp, int t, int m, int n, double inv_lambda,
                                         double *__restrict__ Jc, double *__restrict__ hc)
{
   const int i = t - 1;
   int row = a.con_row0[i];
   for (int ci = 0; ci < a.n_con; ci++)
   {
      const OcbConDev &c = a.cons[ci];
      if (!con_applies(c, i, m)) continue;
      /* first walk: the frame */
      double R[9], tr[3], ax[3], org[3];
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = 0.0;
      R[0] = R[4] = R[8] = 1.0;
      tr[0] = tr[1] = tr[2] = 0.0;
      ConFrame f;
      if (c.joint < 0)
         con_frame(c, R, tr, f);
      else
         for (int j = 0; j <= c.joint; j++)
         {
            const OcbJointDev &J = a.joints[j];
            /* SAVE_FRAMES: no forward sweep has left the branch frames of this waypoint in `slots` (tiled path):
             * this walk saves them itself, the second one reloads them */
            fk_step<SAVE_FRAMES, FLOAT>(J, Ts[J.dof * Pp + t], slots, Pp, t, R, tr, ax, org, Ts + t, Pp);
            if (j == c.joint) con_frame(c, R, tr, f);
         }
      double *Jr = Jc + (size_t) row * n;
      for (int e = 0; e < c.k * n; e++) Jr[e] = 0.0;
      for (int r = 0; r < c.k; r++) hc[row + r] = f.xyzypr[c.rows[r]];
      /* second walk: the columns of the joints that move the frame */
      if (c.joint >= 0)
      {
         for (int j = 0; j <= c.joint; j++)
         {
            const OcbJointDev &J = a.joints[j];
            fk_step<false, FLOAT>(J, Ts[J.dof * Pp + t], slots, Pp, t, R, tr, ax, org, Ts + t, Pp);
            if (!((c.anc >> j) & 1ull) || J.c0 == 0.0) continue;
            double w[3], v[3], col[6];
            if (J.type == OCB_JOINT_REVOLUTE)
            {
               w[0] = ax[0]; w[1] = ax[1]; w[2] = ax[2];
               v[0] = org[1] * ax[2] - org[2] * ax[1];
               v[1] = org[2] * ax[0] - org[0] * ax[2];
               v[2] = org[0] * ax[1] - org[1] * ax[0];
            }
            else
            {
               w[0] = w[1] = w[2] = 0.0;
               v[0] = ax[0]; v[1] = ax[1]; v[2] = ax[2];
            }
            con_map_column(c, f, w, v, col);
            for (int r = 0; r < c.k; r++) Jr[r * n + J.dof] = fma(J.c0, col[c.rows[r]], Jr[r * n + J.dof]);
         }
         if (FLOAT)
         {
            /* the seven pose entries (spatial.c:295-337): columns of [omega; v at the origin] */
            const double x = Ts[t], y = Ts[Pp + t], z = Ts[2 * Pp + t];
            const double bx = 2.0 * Ts[3 * Pp + t], by = 2.0 * Ts[4 * Pp + t], bz = 2.0 * Ts[5 * Pp + t],
                         bw = 2.0 * Ts[6 * Pp + t];
            const double Jsp[6][7] = {
               {0, 0, 0, bw, -bz, by, -bx},
               {0, 0, 0, bz, bw, -bx, -by},
               {0, 0, 0, -by, bx, bw, -bz},
               {1, 0, 0, -z * bz - y * by, -z * bw + y * bx, z * bx + y * bw, z * by - y * bz},
               {0, 1, 0, z * bw + x * by, -z * bz - x * bx, z * by - x * bw, -z * bx + x * bz},
               {0, 0, 1, -y * bw + x * bz, y * bz + x * bw, -y * by - x * bx, y * bx - x * by}};
            for (int e = 0; e < 7; e++)
            {
               const double w[3] = {Jsp[0][e], Jsp[1][e], Jsp[2][e]}, v[3] = {Jsp[3][e], Jsp[4][e], Jsp[5][e]};
               double col[6];
               con_map_column(c, f, w, v, col);
               for (int r = 0; r < c.k; r++) Jr[r * n + e] = col[c.rows[r]];
            }
         }
      }
      for (int r = 0; r < c.k; r++)
      {
         double acc = 0.0;
         for (int j = 0; j < n; j++) acc += Jr[r * n + j] * AGc[j * Pp + t];
         hc[row + r] = fma(-inv_lambda, acc, hc[row + r]);
      }
      row += c.k;
   }
}

/* S = J Ainv J^T over the stacked rows (chomp.c:567-575): entry (r, q) couples the waypoints of the two rows */
__device__ inline void con_build_system(const OcbChompArgs &a, const double *__restrict__ Jc, double *__restrict__ S,
                                        int m, int n)
{
   const int K = a.con_K;
   for (int e = threadIdx.x; e < K * K; e += blockDim.x)
   {
      const int r = e / K, q = e - r * K;
      const double *Ja = Jc + (size_t) r * n, *Jb = Jc + (size_t) q * n;
      double acc = 0.0;
      for (int j = 0; j < n; j++) acc += Ja[j] * Jb[j];
      S[e] = __ldg(a.Ainv + (size_t) a.con_row_wp[r] * m + a.con_row_wp[q]) * acc;
   }
}

/* S x = b by the whole block: elimination with partial pivoting on rows (the factorisation LAPACKE_dgesv
 * performs, chomp.c:579-581), then back substitution by warp 0.  S (K x K, row-major) and b live in
 * global memory and are overwritten; x is left in b.  red / ired: block scratch in shared memory (>= 10).
 * Returns non-zero when a pivot is exactly zero (b is then not a solution). */
__device__ inline int con_solve(double *__restrict__ S, double *__restrict__ b, int K, double *red, int *ired)
{
   const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = NT >> 5;
   for (int k = 0; k < K; k++)
   {
      /* pivot: largest |S[i][k]|, i >= k; the first of equals */
      double best = -1.0;
      int at = K;
      for (int i = k + tid; i < K; i += NT)
      {
         const double v = fabs(S[(size_t) i * K + k]);
         if (v > best) { best = v; at = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
         const double ov = __shfl_xor_sync(0xffffffffu, best, o);
         const int oa = __shfl_xor_sync(0xffffffffu, at, o);
         if (ov > best || (ov == best && oa < at)) { best = ov; at = oa; }
      }
      if (lane == 0) { red[warp] = best; ired[warp] = at; }
      __syncthreads();
      best = red[0];
      at = ired[0];
      for (int w = 1; w < nw; w++)
         if (red[w] > best || (red[w] == best && ired[w] < at)) { best = red[w]; at = ired[w]; }
      if (!(best > 0.0)) return k + 1; /* uniform: every thread reads the same partials */
      if (at != k)
      {
         for (int j = k + tid; j < K; j += NT)
         {
            const double u = S[(size_t) k * K + j];
            S[(size_t) k * K + j] = S[(size_t) at * K + j];
            S[(size_t) at * K + j] = u;
         }
         if (tid == 0) { const double u = b[k]; b[k] = b[at]; b[at] = u; }
      }
      __syncthreads();
      /* eliminate column k below the pivot: one warp per row, lanes across the row */
      const double pinv = 1.0 / S[(size_t) k * K + k];
      const double bk = b[k];
      for (int i = k + 1 + warp; i < K; i += nw)
      {
         const double f = S[(size_t) i * K + k] * pinv;
         if (f != 0.0)
         {
            for (int j = k + 1 + lane; j < K; j += 32) S[(size_t) i * K + j] = fma(-f, S[(size_t) k * K + j], S[(size_t) i * K + j]);
            if (lane == 0) b[i] = fma(-f, bk, b[i]);
         }
      }
      __syncthreads();
   }
   if (warp == 0)
      for (int k = K - 1; k >= 0; k--)
      {
         double acc = 0.0;
         for (int j = k + 1 + lane; j < K; j += 32) acc += S[(size_t) k * K + j] * b[j];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
         if (lane == 0) b[k] = (b[k] - acc) / S[(size_t) k * K + k];
         __syncwarp();
      }
   __syncthreads();
   return 0;
}

/* ------------------------------------------------------------------------- */
/* Tridiagonal metric: the projection without the dense system.
 *
 * The correction the reference applies, d = -A^-1 J^T x with (J A^-1 J^T) x = h (chomp.c:567-599), is the
 * minimiser of 1/2 d^T A d under J_i d_i = -h_i on every constrained waypoint.  With A tridiagonal
 * (diagonal a_i, off-diagonal b_i, the same for every dof) its optimality conditions
 *     b_{i-1} d_{i-1} + a_i d_i + b_i d_{i+1} + J_i^T x_i = 0,     J_i d_i = -h_i
 * are block tridiagonal.  Eliminating waypoint after waypoint from one end leaves, at waypoint i,
 *     P_i = a_i I - b^2 N_prev,  r_i = -b v_prev                           (what the eliminated side contributes)
 *     W = P_i^-1 J_i^T,  Q = J_i W,  N_i = P_i^-1 - W Q^-1 W^T,  v_i = N_i r_i - W Q^-1 h_i
 *     d_i = -b N_i d_next + v_i
 * P_i and Q are symmetric positive definite (no pivoting).  O(m n^3) instead of O((k m)^3) work and
 * m (n^2 + n) instead of (k m)^2 doubles; same solution as the reference's up to rounding.
 *
 * Two warps eliminate from the two ends towards the middle waypoint, which sees both sides; the two
 * substitutions then run outwards, again side by side.  Within a warp every stage spreads the entries of
 * its small matrix over the lanes, one __syncwarp per stage; both inverses by Gauss-Jordan steps between
 * two buffers.  Every matrix lives in `scr` (shared memory when the run's workspace has room):
 * N[m][n][n], v[m][n] (d on return), then one work area per warp. */
struct ConWork
{
   double *B0, *B1, *W, *U, *Q0, *Q1, *rv;
};

__device__ inline ConWork con_work(double *base, int n, int kmax)
{
   ConWork w;
   w.B0 = base; w.B1 = w.B0 + n * n;
   w.W = w.B1 + n * n; w.U = w.W + n * kmax;
   w.Q0 = w.U + n * kmax; w.Q1 = w.Q0 + kmax * kmax;
   w.rv = w.Q1 + kmax * kmax;
   return w;
}

/* one elimination step at waypoint i, executed by one warp: N_i and v_i from up to two eliminated
 * neighbours (coupling b, their N and v; N == nullptr: none).  Returns 1 when the waypoint's constraint
 * rows were linearly dependent and have been skipped (the reference's dgesv fails on such a system).
 * NN / KK: the number of dofs / of rows on every waypoint when known at compile time (0: read at run
 * time) -- with literal bounds the small dot products unroll and their loads overlap. */
template <int NN, int KK>
__device__ __forceinline__ int con_sweep_step(const OcbChompArgs &a, const double *__restrict__ Jc,
                                              const double *__restrict__ hc, const ConWork &w, double *__restrict__ N,
                                              double *__restrict__ v, int i, int n_rt, double b1, const double *N1,
                                              const double *v1, double b2, const double *N2, const double *v2)
{
   const int lane = threadIdx.x & 31;
   const int n = NN ? NN : n_rt;
   const int kmax = KK ? KK : a.con_kmax, nn = n * n;
   const double ai = __ldg(a.Aband + 3 * i + 1);
   const int r0 = a.con_row0[i];
   const int k = KK ? KK : a.con_row0[i + 1] - r0;
   const double *J = Jc + (size_t) r0 * n, *h = hc + r0;
   /* entry e of a row-major matrix with `cols` columns -> (row, column); exact for the few hundred entries there are */
   auto split = [](int e, int cols, float inv, int &r, int &c) { r = (int) ((e + 0.5f) * inv); c = e - r * cols; };
   const float inv_n = 1.0f / n, inv_k = 1.0f / (k > 0 ? k : 1);
   const double bb1 = b1 * b1, bb2 = b2 * b2;
#pragma unroll
   for (int e = lane; e < nn; e += 32)
   {
      int r, c;
      split(e, n, inv_n, r, c);
      double p = (r == c) ? ai : 0.0;
      if (N1) p = fma(-bb1, N1[e], p);
      if (N2) p = fma(-bb2, N2[e], p);
      w.B0[e] = p;
   }
   if (lane < n)
   {
      double r = 0.0;
      if (N1) r = -b1 * v1[lane];
      if (N2) r = fma(-b2, v2[lane], r);
      w.rv[lane] = r;
   }
   __syncwarp();
   /* P^-1: Gauss-Jordan steps between two buffers */
   double *src = w.B0, *dst = w.B1;
#pragma unroll
   for (int j = 0; j < n; j++)
   {
      const double pj = src[j * n + j];
      const double piv = __drcp_rn(pj > 0.0 ? pj : 1e-300);
#pragma unroll
      for (int e = lane; e < nn; e += 32)
      {
         int r, c;
         split(e, n, inv_n, r, c);
         const double f = src[r * n + j] * piv, sjc = src[j * n + c], se = src[e];
         double out;
         if (r == j) out = (c == j) ? piv : se * piv;
         else out = (c == j) ? -f : fma(-f, sjc, se);
         dst[e] = out;
      }
      __syncwarp();
      double *t = src; src = dst; dst = t;
   }
   const double *Pi = src;
   bool constrained = k > 0;
   const double *Qi = w.Q0;
   if (constrained)
   {
#pragma unroll
      for (int e = lane; e < n * k; e += 32)
      {
         int r, q;
         split(e, k, inv_k, r, q);
         double acc = 0.0;
#pragma unroll
         for (int c = 0; c < n; c++) acc = fma(Pi[r * n + c], J[q * n + c], acc);
         w.W[r * kmax + q] = acc;
      }
      __syncwarp();
      for (int e = lane; e < k * k; e += 32)
      {
         int p, q;
         split(e, k, inv_k, p, q);
         double acc = 0.0;
#pragma unroll
         for (int c = 0; c < n; c++) acc = fma(J[p * n + c], w.W[c * kmax + q], acc);
         w.Q0[p * kmax + q] = acc;
      }
      __syncwarp();
      /* Q^-1; rows that depend on the others show up as a vanishing pivot */
      double scale = 0.0;
#pragma unroll
      for (int j = 0; j < k; j++) scale = fmax(scale, w.Q0[j * kmax + j]);
      double *qs = w.Q0, *qd = w.Q1;
#pragma unroll
      for (int j = 0; j < k; j++)
      {
         const double pj = qs[j * kmax + j];
         if (!(pj > 1e-13 * scale)) { constrained = false; break; } /* uniform: every lane reads the same entry */
         const double piv = __drcp_rn(pj);
         for (int e = lane; e < k * k; e += 32)
         {
            int r, c;
            split(e, k, inv_k, r, c);
            const double f = qs[r * kmax + j] * piv, sjc = qs[j * kmax + c], se = qs[r * kmax + c];
            double out;
            if (r == j) out = (c == j) ? piv : se * piv;
            else out = (c == j) ? -f : fma(-f, sjc, se);
            qd[r * kmax + c] = out;
         }
         __syncwarp();
         double *t = qs; qs = qd; qd = t;
      }
      Qi = qs;
   }
   const int skipped = (k > 0 && !constrained) ? 1 : 0;
   if (constrained)
   {
#pragma unroll
      for (int e = lane; e < n * k; e += 32)
      {
         int r, q;
         split(e, k, inv_k, r, q);
         double acc = 0.0;
#pragma unroll
         for (int p = 0; p < k; p++) acc = fma(w.W[r * kmax + p], Qi[p * kmax + q], acc);
         w.U[r * kmax + q] = acc;
      }
      __syncwarp();
   }
#pragma unroll
   for (int e = lane; e < nn; e += 32)
   {
      int r, c;
      split(e, n, inv_n, r, c);
      double acc = Pi[e];
      if (constrained)
      {
#pragma unroll
         for (int p = 0; p < k; p++) acc = fma(-w.U[r * kmax + p], w.W[c * kmax + p], acc);
      }
      N[e] = acc;
   }
   __syncwarp();
   if (lane < n)
   {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < n; c++) acc = fma(N[lane * n + c], w.rv[c], acc);
      if (constrained)
      {
#pragma unroll
         for (int p = 0; p < k; p++) acc = fma(-w.U[lane * kmax + p], h[p], acc);
      }
      v[lane] = acc;
   }
   __syncwarp();
   return skipped;
}

/* the whole projection, called by every thread of the block (it holds block barriers); d is left in the
 * v rows of scr.  Returns (to thread 0) the number of skipped waypoints. */
template <int NN, int KK>
__device__ __noinline__ int con_project_tridiag(const OcbChompArgs &a, const double *__restrict__ Jc,
                                                const double *__restrict__ hc, double *__restrict__ scr, int m, int n)
{
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int nn = n * n, kmax = a.con_kmax;
   double *Nall = scr, *vall = Nall + (size_t) m * nn;
   double *work0 = vall + (size_t) m * n;
   const size_t work_size = 2 * (size_t) nn + 2 * (size_t) n * kmax + 2 * (size_t) kmax * kmax + n;
   const bool two = (blockDim.x >= 64) && m >= 3;
   const int mid = two ? m / 2 : m - 1; /* the waypoint both eliminations stop at (one-sided: the last) */
   int skipped = 0;
   if (warp == 0)
   {
      const ConWork w = con_work(work0, n, kmax);
      for (int i = 0; i < mid; i++)
      {
         const double b = (i > 0) ? __ldg(a.Aband + 3 * i) : 0.0; /* A[i][i-1] */
         skipped += con_sweep_step<NN, KK>(a, Jc, hc, w, Nall + (size_t) i * nn, vall + (size_t) i * n, i, n, b,
                                   i > 0 ? Nall + (size_t) (i - 1) * nn : nullptr, vall + (size_t) (i - 1) * n, 0.0, nullptr, nullptr);
      }
   }
   else if (warp == 1 && two)
   {
      const ConWork w = con_work(work0 + work_size, n, kmax);
      for (int i = m - 1; i > mid; i--)
      {
         const double b = (i < m - 1) ? __ldg(a.Aband + 3 * i + 2) : 0.0; /* A[i][i+1] */
         skipped += con_sweep_step<NN, KK>(a, Jc, hc, w, Nall + (size_t) i * nn, vall + (size_t) i * n, i, n, b,
                                   i < m - 1 ? Nall + (size_t) (i + 1) * nn : nullptr, vall + (size_t) (i + 1) * n, 0.0, nullptr, nullptr);
      }
      if (lane == 0) work0[2 * work_size] = (double) skipped; /* handed to thread 0 below */
   }
   __syncthreads();
   if (warp == 0)
   {
      /* the middle waypoint: both sides eliminated, nothing left to couple to -> v is d */
      const ConWork w = con_work(work0, n, kmax);
      const int i = mid;
      const double bl = (i > 0) ? __ldg(a.Aband + 3 * i) : 0.0, br = (i < m - 1) ? __ldg(a.Aband + 3 * i + 2) : 0.0;
      skipped += con_sweep_step<NN, KK>(a, Jc, hc, w, Nall + (size_t) i * nn, vall + (size_t) i * n, i, n,
                                bl, i > 0 ? Nall + (size_t) (i - 1) * nn : nullptr, vall + (size_t) (i - 1) * n,
                                br, (two && i < m - 1) ? Nall + (size_t) (i + 1) * nn : nullptr, vall + (size_t) (i + 1) * n);
      if (two) skipped += (int) work0[2 * work_size];
   }
   __syncthreads();
   /* substitution outwards from the middle: d_i = -b N_i d_next + v_i, in place */
   if (warp == 0)
      for (int i = mid - 1; i >= 0; i--)
      {
         const double b = __ldg(a.Aband + 3 * i + 2); /* A[i][i+1] */
         const double *N = Nall + (size_t) i * nn, *dn = vall + (size_t) (i + 1) * n;
         if (lane < n)
         {
            double acc = 0.0;
            for (int c = 0; c < n; c++) acc = fma(N[lane * n + c], dn[c], acc);
            vall[(size_t) i * n + lane] = fma(-b, acc, vall[(size_t) i * n + lane]);
         }
         __syncwarp();
      }
   else if (warp == 1 && two)
      for (int i = mid + 1; i < m; i++)
      {
         const double b = __ldg(a.Aband + 3 * i); /* A[i][i-1] */
         const double *N = Nall + (size_t) i * nn, *dn = vall + (size_t) (i - 1) * n;
         if (lane < n)
         {
            double acc = 0.0;
            for (int c = 0; c < n; c++) acc = fma(N[lane * n + c], dn[c], acc);
            vall[(size_t) i * n + lane] = fma(-b, acc, vall[(size_t) i * n + lane]);
         }
         __syncwarp();
      }
   __syncthreads();
   return skipped;
}

/* The constrained update of one run by its whole block (chomp.c:525-605 with constraints): on entry Gs holds the
 * complete gradient G (barrier done).  AG = A^-1 G (or the momentum form, leapfrog_first is cleared); constraint
 * values h and Jacobians J at the current T; the correction -A^-1 J^T x with (J A^-1 J^T) x = h - J AG / lambda;
 * T -= AG / lambda + A^-1 J^T x, which zeroes the linearised constraints at the new T.  Tridiagonal metric:
 * con_project_tridiag; else the dense system in global scratch, solved by the block.  ws: the run's shared
 * workspace when J / h / the sweep's matrices may live there (con_jh_smem / con_rec_smem), else unused;
 * slots: the saved branch frames of the forward sweep ([12 slot][Pp]; SAVE_FRAMES: scratch for them).
 * Returns this thread's joint-limit violation flag. */
template <bool FLOAT, bool SAVE_FRAMES>
__device__ __forceinline__ int con_update(const OcbChompArgs &a, const int run, double *__restrict__ Ts,
                                          double *__restrict__ Gs, double *__restrict__ AGs, double *ws, double *slots,
                                          double *red, int *ired, const int Pp, const int m, const int n,
                                          const double inv_lambda, int &leapfrog_first)
{
   const int tid = threadIdx.x, NT = blockDim.x;
   const int K = a.con_K;
   int violated = 0;
   double *gs = a.con_scratch + (size_t) run * a.con_stride; /* global: J, h, saved h, then S or the sweep's matrices */
   double *Jc = a.con_jh_smem ? ws : gs, *hc = Jc + (size_t) K * n, *h0 = gs + (size_t) K * (n + 1), *S = h0 + K;
   block_band_solve(a, Gs, Pp, m, n);
   __syncthreads();
   const double *AGc = Gs;
   if (a.use_momentum)
   {
      const double coef = (leapfrog_first ? 0.5 : 1.0) * inv_lambda;
      for (int t = tid + 1; t <= m; t += NT)
         for (int j = 0; j < n; j++) AGs[j * Pp + t] = fma(coef, Gs[j * Pp + t], AGs[j * Pp + t]);
      leapfrog_first = 0;
      AGc = AGs;
   }
   for (int t = tid + 1; t <= m && K > 0; t += NT)
      if (a.con_row0[t] > a.con_row0[t - 1])
         con_eval_waypoint<FLOAT, SAVE_FRAMES>(a, Ts, slots, AGc, Pp, t, m, n, inv_lambda, Jc, hc);
   __syncthreads();
   if (K > 0 && a.con_fast)
   {
      /* tridiagonal metric: d = -A^-1 J^T x straight from two sweeps over the waypoints */
      double *scr = a.con_rec_smem ? ws + a.con_rec_off : S;
      /* the common shapes with their sizes as literals (7 dofs; the same 3 or 6 rows on every waypoint) */
      const int ku = a.con_kuniform;
      const int skipped = (n == 7 && ku == 3) ? con_project_tridiag<7, 3>(a, Jc, hc, scr, m, n)
                        : (n == 7 && ku == 6) ? con_project_tridiag<7, 6>(a, Jc, hc, scr, m, n)
                        : (n == 7)            ? con_project_tridiag<7, 0>(a, Jc, hc, scr, m, n)
                                              : con_project_tridiag<0, 0>(a, Jc, hc, scr, m, n);
      if (tid == 0 && skipped) a.con_singular[run] += skipped;
      const double *d = scr + (size_t) m * n * n;
      for (int t = tid + 1; t <= m; t += NT)
         for (int j = 0; j < n; j++)
         {
            const double q = fma(-inv_lambda, AGc[j * Pp + t], Ts[j * Pp + t]) + d[(t - 1) * n + j];
            Ts[j * Pp + t] = q;
            violated |= (q < __ldg(a.lim_lo + j)) | (q > __ldg(a.lim_hi + j));
         }
      return violated;
   }
   if (K > 0)
   {
      con_build_system(a, Jc, S, m, n);
      for (int e = tid; e < K; e += NT) h0[e] = hc[e];
   }
   __syncthreads();
   if (K > 0 && con_solve(S, hc, K, red, ired))
   {
      /* zero pivot: dgesv leaves the right-hand side as it was and the reference carries on with it */
      __syncthreads();
      for (int e = tid; e < K; e += NT) hc[e] = h0[e];
      if (tid == 0) a.con_singular[run]++;
      __syncthreads();
   }
   for (int t = tid + 1; t <= m; t += NT)
   {
      const int r0 = K > 0 ? a.con_row0[t - 1] : 0, r1 = K > 0 ? a.con_row0[t] : 0;
      for (int j = 0; j < n; j++)
      {
         Ts[j * Pp + t] = fma(-inv_lambda, AGc[j * Pp + t], Ts[j * Pp + t]);
         double d = 0.0;
         for (int r = r0; r < r1; r++) d += Jc[(size_t) r * n + j] * hc[r];
         Gs[j * Pp + t] = d;
      }
   }
   __syncthreads();
   block_band_solve(a, Gs, Pp, m, n);
   __syncthreads();
   for (int t = tid + 1; t <= m; t += NT)
      for (int j = 0; j < n; j++)
      {
         const double q = Ts[j * Pp + t] - Gs[j * Pp + t];
         Ts[j * Pp + t] = q;
         violated |= (q < __ldg(a.lim_lo + j)) | (q > __ldg(a.lim_hi + j));
      }
   return violated;
}

} /* namespace */

#endif
