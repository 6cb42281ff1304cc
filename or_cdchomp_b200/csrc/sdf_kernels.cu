/* sdf_kernels.cu -- environment signed-distance-field build on sm_100a.
 *
 * Replaces, for 3-d grids of doubles (paths relative to the reference root):
 *   cd_grid_double_dt_sqeuc + sedt_onedim   src/libcd/grid.c:462-569, 269-329
 *   cd_grid_double_bin_sdf                  src/libcd/grid.c:637-687
 *   cd_grid_flood_fill + 1.0->HUGE_VAL      src/libcd/grid_flood.c:30-111,
 *                                           src/orcdchomp_mod.cpp:143-151, 543-548
 *   the occupancy loop                      src/orcdchomp_mod.cpp:498-525
 *                                           (+ cd_grid_center_index, grid.c:172-189)
 *
 * This file is compiled with -fmad=false: every product and sum rounds on its
 * own, so the general distance transform reproduces the reference's arithmetic
 * operation for operation and the occupancy predicate agrees bit for bit with
 * the CPU oracle (oracle/orcdchomp_port.c, built with -ffp-contract=off).
 *
 * General path (any finite sample heights, any cell pitch): one thread per grid
 * line runs the lower-envelope scan; the envelope stack lives in HBM scratch laid
 * out [entry][line] so a warp's accesses coalesce while depths stay similar.
 */
#include <math.h>
#include "ocb_internal.h"
#include "../../include/orcdchomp_b200.h"

namespace
{

/* ------------------------------------------------------------------ EDT pass */
struct EdtPass
{
   const double *src; /* read from here ... */
   double *dst;       /* ... write here (may alias src) */
   int len;           /* cells along the transformed dimension */
   size_t stride;     /* element stride along it = product of faster dimensions */
   size_t inner;      /* == stride */
   size_t nlines;
   double pitch2;     /* (length/size)^2, grid.c:518 */
   int flip;          /* 1: treat the input as (x == 0 ? HUGE_VAL : 0), grid.c:657-663 */
   int *stk_v;        /* [len][nlines]   */
   double *stk_z;     /* [len+1][nlines] */
   double *stk_f;     /* [len][nlines]   */
};

__global__ void __launch_bounds__(128) edt_pass_kernel(const EdtPass p)
{
   const size_t line = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
   if (line >= p.nlines) return;
   const size_t o = line / p.inner, i = line % p.inner;
   const size_t base = o * p.inner * (size_t) p.len + i;
   const double inf = HUGE_VAL;
   const size_t L = p.nlines;

   /* build the lower envelope (grid.c:280-312); the top entry is kept in registers */
   int np = 0;
   int v_top = 0;
   double z_top = 0.0, f_top = 0.0;
   for (int q = 0; q < p.len; q++)
   {
      double x = p.src[base + (size_t) q * p.stride];
      if (p.flip) x = (x == 0.0) ? inf : 0.0;
      const double fq = x / p.pitch2;
      if (fq == inf) continue;
      if (np == 0)
      {
         np = 1;
         v_top = q; z_top = -inf; f_top = fq;
         continue;
      }
      double s;
      for (;;)
      {
         s = fq + (double) (q * q);
         s -= f_top + (double) (v_top * v_top);
         s /= 2.0 * (double) (q - v_top);
         if (s <= z_top)
         {
            np--; /* z_top == -inf for the last entry, so np never reaches 0 here */
            v_top = p.stk_v[(size_t) (np - 1) * L + line];
            z_top = p.stk_z[(size_t) (np - 1) * L + line];
            f_top = p.stk_f[(size_t) (np - 1) * L + line];
         }
         else
            break;
      }
      p.stk_v[(size_t) (np - 1) * L + line] = v_top;
      p.stk_z[(size_t) (np - 1) * L + line] = z_top;
      p.stk_f[(size_t) (np - 1) * L + line] = f_top;
      np++;
      v_top = q; z_top = s; f_top = fq;
   }
   if (np == 0)
   {
      for (int q = 0; q < p.len; q++) p.dst[base + (size_t) q * p.stride] = inf; /* inf * pitch2 = inf */
      return;
   }
   p.stk_v[(size_t) (np - 1) * L + line] = v_top;
   p.stk_z[(size_t) (np - 1) * L + line] = z_top;
   p.stk_f[(size_t) (np - 1) * L + line] = f_top;

   /* read the envelope back (grid.c:321-326) */
   int k = 0;
   int v = p.stk_v[line];
   double f = p.stk_f[line];
   double z_next = (np > 1) ? p.stk_z[L + line] : inf;
   for (int q = 0; q < p.len; q++)
   {
      while (z_next < (double) q)
      {
         k++;
         v = p.stk_v[(size_t) k * L + line];
         f = p.stk_f[(size_t) k * L + line];
         z_next = (k + 1 < np) ? p.stk_z[(size_t) (k + 1) * L + line] : inf;
      }
      const double d = (double) (q - v);
      double out = d * d + f;
      out *= p.pitch2;
      p.dst[base + (size_t) q * p.stride] = out;
   }
}

/* sqrt(sedt_obs) - sqrt(sedt_emp), grid.c:674-679 */
__global__ void sdf_combine_kernel(const double *__restrict__ d_obs, const double *__restrict__ d_emp,
                                   double *__restrict__ out, size_t n)
{
   for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
      out[i] = sqrt(d_obs[i]) - sqrt(d_emp[i]);
}

/* ---------------------------------------------------------------- occupancy */
struct PrimDev
{
   int type;
   double c[3];   /* centre in the grid frame */
   double e[3];   /* half extents (box) / e[0] = radius (sphere) */
   double R[9];   /* box axes = columns */
   double A[9];   /* |R| + 1e-12 */
   double V[9];   /* triangle vertices (OCB_PRIM_TRIANGLE) */
   double lo[3], hi[3]; /* axis-aligned bounds of the primitive in the grid frame (conservative) */
};

__global__ void prep_prims_kernel(const ocb_prim *prims, PrimDev *out, int n)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const ocb_prim &p = prims[i];
   PrimDev d;
   d.type = p.type;
   for (int k = 0; k < 3; k++) { d.c[k] = p.pose[k]; d.e[k] = p.extents[k]; }
   const double qx = p.pose[3], qy = p.pose[4], qz = p.pose[5], qw = p.pose[6];
   const double qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz, qw2 = qw * qw;
   const double qxqy = qx * qy, qxqz = qx * qz, qxqw = qx * qw;
   const double qyqz = qy * qz, qyqw = qy * qw, qzqw = qz * qw;
   d.R[0] = qx2 - qy2 - qz2 + qw2; d.R[1] = 2 * (qxqy - qzqw);      d.R[2] = 2 * (qxqz + qyqw);
   d.R[3] = 2 * (qxqy + qzqw);     d.R[4] = -qx2 + qy2 - qz2 + qw2; d.R[5] = 2 * (qyqz - qxqw);
   d.R[6] = 2 * (qxqz - qyqw);     d.R[7] = 2 * (qyqz + qxqw);      d.R[8] = -qx2 - qy2 + qz2 + qw2;
   for (int k = 0; k < 9; k++) d.A[k] = fabs(d.R[k]) + 1e-12;
   for (int k = 0; k < 7; k++) d.V[k] = p.pose[k];
   d.V[7] = p.extents[0];
   d.V[8] = p.extents[1];
   /* bounds used only to choose which voxels to test (the exact predicate decides every hit) */
   for (int k = 0; k < 3; k++)
   {
      if (p.type == OCB_PRIM_TRIANGLE)
      {
         d.lo[k] = fmin(d.V[k], fmin(d.V[3 + k], d.V[6 + k]));
         d.hi[k] = fmax(d.V[k], fmax(d.V[3 + k], d.V[6 + k]));
      }
      else
      {
         const double w = (p.type == OCB_PRIM_SPHERE) ? p.extents[0]
                          : d.A[3 * k] * p.extents[0] + d.A[3 * k + 1] * p.extents[1] + d.A[3 * k + 2] * p.extents[2];
         d.lo[k] = p.pose[k] - w;
         d.hi[k] = p.pose[k] + w;
      }
   }
   out[i] = d;
}

/* cube (centre c, half extent h, axes = grid axes) against one primitive; the
 * operation order is the contract shared with oracle/orcdchomp_port.c */
/* cube against a triangle: 9 edge cross products, 3 cube face normals, the triangle plane
 * (Akenine-Moller 2001); strict inequalities, touching is a hit */
__device__ __forceinline__ bool cube_hits_triangle(const double c[3], double h, const double *V)
{
   double v[3][3], e[3][3], n[3], vmin[3], vmax[3];
#pragma unroll
   for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) v[i][k] = V[3 * i + k] - c[k];
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      e[0][k] = v[1][k] - v[0][k];
      e[1][k] = v[2][k] - v[1][k];
      e[2][k] = v[0][k] - v[2][k];
   }
#pragma unroll
   for (int i = 0; i < 3; i++)
   {
      const double ex = e[i][0], ey = e[i][1], ez = e[i][2];
      const double fx = fabs(ex), fy = fabs(ey), fz = fabs(ez);
      double p0, p1, p2, lo, hi, rad;
      p0 = ey * v[0][2] - ez * v[0][1]; p1 = ey * v[1][2] - ez * v[1][1]; p2 = ey * v[2][2] - ez * v[2][1];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fz * h + fy * h;
      if (lo > rad || hi < -rad) return false;
      p0 = ez * v[0][0] - ex * v[0][2]; p1 = ez * v[1][0] - ex * v[1][2]; p2 = ez * v[2][0] - ex * v[2][2];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fz * h + fx * h;
      if (lo > rad || hi < -rad) return false;
      p0 = ex * v[0][1] - ey * v[0][0]; p1 = ex * v[1][1] - ey * v[1][0]; p2 = ex * v[2][1] - ey * v[2][0];
      lo = p0; hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      if (p2 < lo) lo = p2; if (p2 > hi) hi = p2;
      rad = fy * h + fx * h;
      if (lo > rad || hi < -rad) return false;
   }
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      double lo = v[0][k], hi = v[0][k];
      if (v[1][k] < lo) lo = v[1][k]; if (v[1][k] > hi) hi = v[1][k];
      if (v[2][k] < lo) lo = v[2][k]; if (v[2][k] > hi) hi = v[2][k];
      if (lo > h || hi < -h) return false;
   }
   n[0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
   n[1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
   n[2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      if (n[k] > 0.0) { vmin[k] = -h - v[0][k]; vmax[k] = h - v[0][k]; }
      else { vmin[k] = h - v[0][k]; vmax[k] = -h - v[0][k]; }
   }
   if (n[0] * vmin[0] + n[1] * vmin[1] + n[2] * vmin[2] > 0.0) return false;
   if (n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2] >= 0.0) return true;
   return false;
}

__device__ __forceinline__ bool cube_hits(const double c[3], double h, const PrimDev &p)
{
   if (p.type == OCB_PRIM_TRIANGLE) return cube_hits_triangle(c, h, p.V);
   if (p.type == OCB_PRIM_SPHERE)
   {
      double d2 = 0.0;
      const double r = p.e[0];
#pragma unroll
      for (int k = 0; k < 3; k++)
      {
         double d = fabs(p.c[k] - c[k]) - h;
         if (d < 0.0) d = 0.0;
         d2 = d2 + d * d;
      }
      return d2 <= r * r;
   }
   const double *R = p.R, *A = p.A, *e = p.e;
   double t[3];
#pragma unroll
   for (int i = 0; i < 3; i++) t[i] = p.c[i] - c[i];
#pragma unroll
   for (int i = 0; i < 3; i++)
   {
      const double ra = h;
      const double rb = e[0] * A[3 * i + 0] + e[1] * A[3 * i + 1] + e[2] * A[3 * i + 2];
      if (fabs(t[i]) > ra + rb) return false;
   }
#pragma unroll
   for (int j = 0; j < 3; j++)
   {
      const double ra = h * A[0 + j] + h * A[3 + j] + h * A[6 + j];
      const double rb = e[j];
      const double tl = t[0] * R[0 + j] + t[1] * R[3 + j] + t[2] * R[6 + j];
      if (fabs(tl) > ra + rb) return false;
   }
#pragma unroll
   for (int i = 0; i < 3; i++)
   {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
      for (int j = 0; j < 3; j++)
      {
         const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
         const double ra = h * A[3 * i1 + j] + h * A[3 * i2 + j];
         const double rb = e[j1] * A[3 * i + j2] + e[j2] * A[3 * i + j1];
         const double tl = t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j];
         if (fabs(tl) > ra + rb) return false;
      }
   }
   return true;
}

/* every voxel starts free (mod.cpp:398) */
__global__ void fill_kernel(double *__restrict__ grid, size_t n, double value)
{
   for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
      grid[i] = value;
}

/* The reference asks OpenRAVE, voxel by voxel, whether a cube collides with anything
 * (mod.cpp:498-525).  Here each primitive rasterises itself: blockIdx.x = primitive,
 * blockIdx.y = a slice of its bounding range of voxels; every voxel in range runs the exact
 * predicate and a hit stores HUGE_VAL (idempotent, so overlapping primitives need no
 * atomics and the result does not depend on scheduling). */
__global__ void __launch_bounds__(256)
rasterize_kernel(const PrimDev *__restrict__ prims, int sx, int sy, int sz,
                 double lx, double ly, double lz, double h, double *__restrict__ grid)
{
   __shared__ PrimDev P;
   {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(prims + blockIdx.x);
      uint32_t *dst = reinterpret_cast<uint32_t *>(&P);
      for (int w = threadIdx.x; w < (int) (sizeof(PrimDev) / 4); w += blockDim.x) dst[w] = src[w];
   }
   __syncthreads();
   const int size[3] = {sx, sy, sz};
   const double len[3] = {lx, ly, lz};
   int lo[3], cnt[3];
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      /* voxel centres within (bounds +- h), widened by one voxel against rounding */
      const double pitch = len[k] / size[k];
      int a = (int) floor((P.lo[k] - h) / pitch) - 1;
      int b = (int) floor((P.hi[k] + h) / pitch) + 1;
      if (a < 0) a = 0;
      if (b > size[k] - 1) b = size[k] - 1;
      lo[k] = a;
      cnt[k] = b - a + 1;
   }
   if (cnt[0] <= 0 || cnt[1] <= 0 || cnt[2] <= 0) return;
   const long long vol = (long long) cnt[0] * cnt[1] * cnt[2];
   for (long long lin = blockIdx.y * (long long) blockDim.x + threadIdx.x; lin < vol;
        lin += (long long) gridDim.y * blockDim.x)
   {
      const int z = lo[2] + (int) (lin % cnt[2]);
      const int y = lo[1] + (int) ((lin / cnt[2]) % cnt[1]);
      const int x = lo[0] + (int) (lin / ((long long) cnt[2] * cnt[1]));
      double c[3];
      /* cd_grid_center_index: ((0.5 + sub) / size) * length   (grid.c:184-187) */
      c[0] = (0.5 + x) / sx * lx;
      c[1] = (0.5 + y) / sy * ly;
      c[2] = (0.5 + z) / sz * lz;
      if (cube_hits(c, h, P)) grid[((size_t) x * sy + y) * sz + z] = HUGE_VAL; /* mod.cpp:522 */
   }
}

/* --------------------------------------------------------------- flood fill */
/* cd_grid_flood_fill(g, start, no wrap, replace_1_to_0) + relabel (grid_flood.c:30-111,
 * mod.cpp:143-151, 536-548) on bit masks.  32 consecutive z cells share one word of
 *   open[(x NY + y) NZW + w]    the cell holds exactly 1.0 (it conducts the fill)
 *   fill[...]                   the fill has reached the cell
 * The fill spreads by whole-line sweeps: inside a word through the carry chain of an integer
 * add, across words and along x / y by one AND + OR per 32 cells.  Both masks of a 400^3 grid
 * are 8 MB each and stay in L2, so a round of three sweeps costs tens of microseconds; rounds
 * repeat until nothing changes (6-connected reachability, independent of the sweep order). */

/* all open bits connected, inside the word, to a seed bit (seeds must be open) */
__device__ __forceinline__ unsigned ripple(unsigned open, unsigned seeds)
{
   const unsigned up = (((open + seeds) ^ open) & open) | seeds;
   const unsigned ro = __brev(open), rs = __brev(seeds);
   const unsigned down = __brev((((ro + rs) ^ ro) & ro) | rs);
   return up | down;
}

/* one warp per z line: 32 cells -> one word by ballot */
__global__ void __launch_bounds__(256)
flood_pack_kernel(const double *__restrict__ grid, unsigned *__restrict__ open, unsigned *__restrict__ fill,
                  size_t nlines, int nz, int nzw, size_t start)
{
   const int lane = threadIdx.x & 31;
   const size_t warp = (blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5;
   const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
   for (size_t line = warp; line < nlines; line += nwarps)
   {
      const double *src = grid + line * (size_t) nz;
      for (int w = 0; w < nzw; w++)
      {
         const int z = 32 * w + lane;
         const bool is_open = (z < nz) && (src[z] == 1.0);
         const unsigned m = __ballot_sync(0xffffffffu, is_open);
         if (lane == 0)
         {
            unsigned f = 0;
            const size_t first = line * (size_t) nz + 32 * (size_t) w;
            if (start >= first && start < first + 32) f = (1u << (unsigned) (start - first)) & m;
            open[line * nzw + w] = m;
            fill[line * nzw + w] = f;
         }
      }
   }
}

/* along z: one thread per (x, y) line, carry between its words */
__global__ void __launch_bounds__(128)
flood_z_kernel(const unsigned *__restrict__ open, unsigned *__restrict__ fill, size_t nlines, int nzw, int *changed)
{
   const size_t line = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
   if (line >= nlines) return;
   const unsigned *o = open + line * nzw;
   unsigned *f = fill + line * nzw;
   bool any = false;
   unsigned carry = 0;
   for (int w = 0; w < nzw; w++)
   {
      const unsigned m = o[w], old = f[w];
      const unsigned now = ripple(m, (old | carry) & m);
      if (now != old) { f[w] = now; any = true; }
      carry = now >> 31;
   }
   carry = 0;
   for (int w = nzw - 1; w >= 0; w--)
   {
      const unsigned m = o[w], old = f[w];
      const unsigned now = ripple(m, (old | carry) & m);
      if (now != old) { f[w] = now; any = true; }
      carry = (now & 1u) << 31;
   }
   if (any) *changed = 1;
}

/* along x or y: one thread per (other axis, word); `len` steps of `stride` words, there and back */
__global__ void __launch_bounds__(128)
flood_xy_kernel(const unsigned *__restrict__ open, unsigned *__restrict__ fill, int len, size_t stride,
                size_t outer_stride, int n_outer, int nzw, int *changed)
{
   const size_t id = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
   if (id >= (size_t) n_outer * nzw) return;
   const size_t base = (id / nzw) * outer_stride + (id % nzw);
   bool any = false;
   /* a thread owns its words for the whole launch, so a batch of them can be fetched ahead of the
    * serial dependency (only ~5 k threads exist: the loads must overlap inside the thread) */
   constexpr int B = 16;
   unsigned prev = 0;
   for (int q0 = 0; q0 < len; q0 += B)
   {
      unsigned m[B], old[B];
#pragma unroll
      for (int k = 0; k < B; k++)
      {
         const int q = min(q0 + k, len - 1);
         m[k] = open[base + (size_t) q * stride];
         old[k] = fill[base + (size_t) q * stride];
      }
#pragma unroll
      for (int k = 0; k < B; k++)
      {
         if (q0 + k >= len) break;
         unsigned now = old[k];
         const unsigned in = prev & m[k] & ~now;
         if (in)
         {
            now = ripple(m[k], (now | in) & m[k]);
            fill[base + (size_t) (q0 + k) * stride] = now;
            any = true;
         }
         prev = now;
      }
   }
   prev = 0;
   for (int q0 = len - 1; q0 >= 0; q0 -= B)
   {
      unsigned m[B], old[B];
#pragma unroll
      for (int k = 0; k < B; k++)
      {
         const int q = max(q0 - k, 0);
         m[k] = open[base + (size_t) q * stride];
         old[k] = fill[base + (size_t) q * stride];
      }
#pragma unroll
      for (int k = 0; k < B; k++)
      {
         if (q0 - k < 0) break;
         unsigned now = old[k];
         const unsigned in = prev & m[k] & ~now;
         if (in)
         {
            now = ripple(m[k], (now | in) & m[k]);
            fill[base + (size_t) (q0 - k) * stride] = now;
            any = true;
         }
         prev = now;
      }
   }
   if (any) *changed = 1;
}

/* reached -> 0.0 (replace_1_to_0, mod.cpp:143-151); open but not reached -> HUGE_VAL (mod.cpp:546-548) */
__global__ void __launch_bounds__(256)
flood_unpack_kernel(double *__restrict__ grid, const unsigned *__restrict__ open, const unsigned *__restrict__ fill,
                    size_t nlines, int nz, int nzw)
{
   const int lane = threadIdx.x & 31;
   const size_t warp = (blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5;
   const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
   for (size_t line = warp; line < nlines; line += nwarps)
   {
      double *dst = grid + line * (size_t) nz;
      for (int w = 0; w < nzw; w++)
      {
         const unsigned m = open[line * nzw + w], f = fill[line * nzw + w];
         const int z = 32 * w + lane;
         if (z < nz && ((m >> lane) & 1u)) dst[z] = ((f >> lane) & 1u) ? 0.0 : HUGE_VAL;
      }
   }
}

int grid_blocks(size_t n, int threads)
{
   size_t b = (n + threads - 1) / threads;
   const size_t cap = 148 * 16;
   if (b > cap) b = cap;
   if (b < 1) b = 1;
   return (int) b;
}

} /* namespace */

/* ------------------------------------------------------------ host launchers */
extern "C" size_t ocb_dt_scratch_bytes(const int sizes[3])
{
   const size_t n = (size_t) sizes[0] * sizes[1] * sizes[2];
   int maxlen = sizes[0];
   if (sizes[1] > maxlen) maxlen = sizes[1];
   if (sizes[2] > maxlen) maxlen = sizes[2];
   /* per cell: int apex + double bound + double height, plus one extra bound row per line */
   return n * (sizeof(int) + 2 * sizeof(double)) + (n / 2 + 1024) * sizeof(double) + 4096;
}

static cudaError_t run_edt(const double *src, double *dst, const int sizes[3], const double lengths[3],
                           int flip, void *scratch, cudaStream_t st, long *launches)
{
   const size_t n = (size_t) sizes[0] * sizes[1] * sizes[2];
   char *sp = (char *) scratch;
   double *stk_z = (double *) sp;
   size_t maxlines = 0;
   for (int d = 0; d < 3; d++) maxlines = (n / sizes[d] > maxlines) ? n / sizes[d] : maxlines;
   double *stk_f = stk_z + n + maxlines;
   int *stk_v = (int *) (stk_f + n);
   for (int d = 0; d < 3; d++)
   {
      EdtPass p;
      p.src = (d == 0) ? src : dst;
      p.dst = dst;
      p.len = sizes[d];
      p.stride = 1;
      for (int d2 = d + 1; d2 < 3; d2++) p.stride *= (size_t) sizes[d2];
      p.inner = p.stride;
      p.nlines = n / (size_t) sizes[d];
      const double pitch = lengths[d] / sizes[d];
      p.pitch2 = pitch * pitch; /* pow(x, 2.0) */
      p.flip = (d == 0) ? flip : 0;
      p.stk_v = stk_v;
      p.stk_z = stk_z;
      p.stk_f = stk_f;
      const int threads = 128;
      const unsigned blocks = (unsigned) ((p.nlines + threads - 1) / threads);
      edt_pass_kernel<<<blocks, threads, 0, st>>>(p);
      if (launches) (*launches)++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
   }
   return cudaSuccess;
}

extern "C" cudaError_t ocb_launch_dt_sqeuc(const double *d_func, double *d_out, const int sizes[3],
                                           const double lengths[3], void *scratch, size_t scratch_bytes,
                                           cudaStream_t st, long *launches)
{
   if (scratch_bytes < ocb_dt_scratch_bytes(sizes)) return cudaErrorInvalidValue;
   return run_edt(d_func, d_out, sizes, lengths, 0, scratch, st, launches);
}

extern "C" size_t ocb_sdf_scratch_bytes(const int sizes[3], const double lengths[3])
{
   (void) lengths;
   const size_t n = (size_t) sizes[0] * sizes[1] * sizes[2];
   return ocb_dt_scratch_bytes(sizes) + 2 * n * sizeof(double) + 512;
}

extern "C" cudaError_t ocb_launch_bin_sdf(const double *d_obs, double *d_sdf, const int sizes[3],
                                          const double lengths[3], void *scratch, size_t scratch_bytes,
                                          cudaStream_t st, long *launches)
{
   if (scratch_bytes < ocb_sdf_scratch_bytes(sizes, lengths)) return cudaErrorInvalidValue;
   const size_t n = (size_t) sizes[0] * sizes[1] * sizes[2];
   double *g_emp = (double *) scratch;
   double *g_obs = g_emp + n;
   void *stk = (void *) (g_obs + n + 32);
   cudaError_t e = run_edt(d_obs, g_emp, sizes, lengths, 0, stk, st, launches);
   if (e != cudaSuccess) return e;
   e = run_edt(d_obs, g_obs, sizes, lengths, 1, stk, st, launches);
   if (e != cudaSuccess) return e;
   sdf_combine_kernel<<<grid_blocks(n, 256), 256, 0, st>>>(g_obs, g_emp, d_sdf, n);
   if (launches) (*launches)++;
   return cudaGetLastError();
}

extern "C" cudaError_t ocb_launch_occupancy(const void *d_prims, int n_prims, const int sizes[3],
                                            const double lengths[3], double cube_extent, double *d_grid,
                                            int slices, cudaStream_t st)
{
   /* the raw ocb_prim array sits at the start of the scratch; the prepared form follows it */
   const size_t raw = ((size_t) (n_prims > 0 ? n_prims : 1) * sizeof(ocb_prim) + 255) & ~(size_t) 255;
   PrimDev *prep = nullptr;
   cudaError_t e = cudaMallocAsync((void **) &prep, (size_t) (n_prims > 0 ? n_prims : 1) * sizeof(PrimDev), st);
   if (e != cudaSuccess) return e;
   (void) raw;
   const size_t n = (size_t) sizes[0] * sizes[1] * sizes[2];
   fill_kernel<<<grid_blocks(n, 256), 256, 0, st>>>(d_grid, n, 1.0);
   if (n_prims > 0)
   {
      prep_prims_kernel<<<(n_prims + 127) / 128, 128, 0, st>>>((const ocb_prim *) d_prims, prep, n_prims);
      /* `slices` blocks share one primitive's voxel range (large boxes vs. small mesh triangles) */
      if (slices < 1) slices = 1;
      if (slices > 64) slices = 64;
      for (int p0 = 0; p0 < n_prims; p0 += 32768)
      {
         const int cntp = (n_prims - p0 < 32768) ? n_prims - p0 : 32768;
         rasterize_kernel<<<dim3(cntp, slices), 256, 0, st>>>(prep + p0, sizes[0], sizes[1], sizes[2], lengths[0],
                                                          lengths[1], lengths[2], cube_extent, d_grid);
      }
   }
   e = cudaGetLastError();
   cudaFreeAsync(prep, st);
   return e;
}

extern "C" size_t ocb_flood_scratch_bytes(const int sizes[3])
{
   const size_t nzw = ((size_t) sizes[2] + 31) / 32;
   return 256 + 2 * (size_t) sizes[0] * sizes[1] * nzw * sizeof(unsigned);
}

extern "C" cudaError_t ocb_launch_flood_relabel(double *d_grid, const int sizes[3], size_t index_start,
                                                void *scratch, size_t scratch_bytes, cudaStream_t st,
                                                long *launches)
{
   if (scratch_bytes < ocb_flood_scratch_bytes(sizes)) return cudaErrorInvalidValue;
   const int nx = sizes[0], ny = sizes[1], nz = sizes[2], nzw = (nz + 31) / 32;
   const size_t nlines = (size_t) nx * ny, nwords = nlines * nzw;
   int *changed = (int *) scratch;
   unsigned *open = (unsigned *) ((char *) scratch + 256);
   unsigned *fill = open + nwords;
   const int pack_blocks = grid_blocks(nlines * 32, 256);
   flood_pack_kernel<<<pack_blocks, 256, 0, st>>>(d_grid, open, fill, nlines, nz, nzw, index_start);
   if (launches) (*launches)++;
   for (int round = 0; round < 100000; round++)
   {
      cudaError_t e = cudaMemsetAsync(changed, 0, sizeof(int), st);
      if (e != cudaSuccess) return e;
      flood_z_kernel<<<(unsigned) ((nlines + 127) / 128), 128, 0, st>>>(open, fill, nlines, nzw, changed);
      /* along y: lines indexed by (x, w), step nzw words;  along x: by (y, w), step ny * nzw words */
      flood_xy_kernel<<<(unsigned) (((size_t) nx * nzw + 127) / 128), 128, 0, st>>>(open, fill, ny, (size_t) nzw,
                                                                                   (size_t) ny * nzw, nx, nzw, changed);
      flood_xy_kernel<<<(unsigned) (((size_t) ny * nzw + 127) / 128), 128, 0, st>>>(open, fill, nx, (size_t) ny * nzw,
                                                                                   (size_t) nzw, ny, nzw, changed);
      if (launches) (*launches) += 3;
      int h = 0;
      e = cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return e;
      e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
      if (!h) break;
   }
   flood_unpack_kernel<<<pack_blocks, 256, 0, st>>>(d_grid, open, fill, nlines, nz, nzw);
   if (launches) (*launches)++;
   return cudaGetLastError();
}
