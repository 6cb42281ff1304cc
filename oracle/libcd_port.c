/* oracle/libcd_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the libcd functions that sit on the CHOMP hot path of
 * personalrobotics/or_cdchomp, with the same names, signatures and struct
 * layouts (oracle/cd_abi.h) so that oracle/orcdchomp_port.c links against
 * either this file ("port") or the real reference objects ("_ref").  It is the
 * checker for the CUDA engine; nothing under or_cdchomp_b200/ may call it.
 *
 * PINNING: tests/test_oracle_vs_ref.py compares every function here against
 * the unmodified reference sources compiled into oracle/_ref/liboracle_ref.so
 * (same inputs, results equal to <= 1e-12 or bit-exact where stated).  The
 * reference itself ships no tests or golden vectors for this path.
 *
 * Differences that are deliberate and documented:
 *   - BLAS/LAPACK calls are replaced by the straightforward triple loops /
 *     Gauss-Jordan inverse below (the reference links an unpinned system BLAS).
 *   - a constraint with k = 0 rows (a TSR none of whose bounds is [0, 0]) adds nothing here.  In the
 *     reference cblas_dgemv(CblasTrans, k = 0, ...) at chomp.c:594-596 returns before scaling its
 *     output ("quick return if possible"), so cons_delta keeps the previous constraint's
 *     correction -- or, for the first constraint of the list, uninitialised memory -- and chomp.c:597-598
 *     applies it once more at that constraint's waypoint.  Not reproduced (and not by the engine).
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference root).
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "cd_abi.h"

/* ------------------------------------------------------------------ helpers */

static void zero_fill(double *a, size_t count)
{
   size_t i;
   for (i = 0; i < count; i++) a[i] = 0.0;
}

/* C(MxN) = alpha * op(A) * op(B) + beta * C, row-major; stands in for
 * cblas_dgemm at chomp.c:450,515,529,540,545,640,665-669 */
static void gemm_rm(int ta, int tb, int M, int N, int K, double alpha,
                    const double *A, int lda, const double *B, int ldb,
                    double beta, double *C, int ldc)
{
   int i, j, k;
   for (i = 0; i < M; i++)
      for (j = 0; j < N; j++)
      {
         double acc = 0.0;
         for (k = 0; k < K; k++)
         {
            double a = ta ? A[k * lda + i] : A[i * lda + k];
            double b = tb ? B[j * ldb + k] : B[k * ldb + j];
            acc += a * b;
         }
         C[i * ldc + j] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * C[i * ldc + j];
      }
}

/* in-place inverse with partial pivoting; stands in for LAPACKE_dgetrf +
 * LAPACKE_dgetri at chomp.c:393-403.  returns non-zero when singular. */
static int invert_rm(double *a, int n)
{
   int i, j, k, piv;
   int *perm = (int *) malloc(n * sizeof(int));
   double *inv = (double *) malloc((size_t) n * n * sizeof(double));
   if (!perm || !inv) { free(perm); free(inv); return -1; }
   for (i = 0; i < n; i++)
      for (j = 0; j < n; j++) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
   for (k = 0; k < n; k++)
   {
      double best = fabs(a[k * n + k]);
      double d;
      piv = k;
      for (i = k + 1; i < n; i++)
         if (fabs(a[i * n + k]) > best) { best = fabs(a[i * n + k]); piv = i; }
      if (best == 0.0) { free(perm); free(inv); return 1; }
      if (piv != k)
         for (j = 0; j < n; j++)
         {
            double t = a[k * n + j]; a[k * n + j] = a[piv * n + j]; a[piv * n + j] = t;
            t = inv[k * n + j]; inv[k * n + j] = inv[piv * n + j]; inv[piv * n + j] = t;
         }
      d = 1.0 / a[k * n + k];
      for (j = 0; j < n; j++) { a[k * n + j] *= d; inv[k * n + j] *= d; }
      for (i = 0; i < n; i++)
      {
         double f;
         if (i == k) continue;
         f = a[i * n + k];
         if (f == 0.0) continue;
         for (j = 0; j < n; j++)
         {
            a[i * n + j] -= f * a[k * n + j];
            inv[i * n + j] -= f * inv[k * n + j];
         }
      }
   }
   memcpy(a, inv, (size_t) n * n * sizeof(double));
   free(perm);
   free(inv);
   return 0;
}

/* ===================================================================== grid */

/* grid.c:61-98 (create_sizeown via create_sizearray 50-59): C-ordered n-d grid,
 * every cell initialised from cell_init, lengths default to 1 */
int cd_grid_create_sizearray(struct cd_grid **gp, void *cell_init, int cell_size, int n, int *sizes)
{
   struct cd_grid *g;
   size_t i;
   int d;
   g = (struct cd_grid *) malloc(sizeof(struct cd_grid));
   if (!g) return -1;
   g->n = n;
   g->cell_size = cell_size;
   g->data = 0;
   g->lengths = 0;
   g->sizes = (int *) malloc(n * sizeof(int));
   if (!g->sizes) { free(g); return -1; }
   g->ncells = 1;
   for (d = 0; d < n; d++) { g->sizes[d] = sizes[d]; g->ncells *= (size_t) sizes[d]; }
   if (cell_size > 0)
   {
      g->data = (char *) malloc(g->ncells * cell_size);
      if (!g->data) { cd_grid_destroy(g); return -2; }
      for (i = 0; i < g->ncells; i++) memcpy(g->data + i * cell_size, cell_init, cell_size);
   }
   g->lengths = (double *) malloc(n * sizeof(double));
   if (!g->lengths) { cd_grid_destroy(g); return -1; }
   for (d = 0; d < n; d++) g->lengths[d] = 1.0;
   *gp = g;
   return 0;
}

/* grid.c:100-132: deep copy incl. lengths */
int cd_grid_create_copy(struct cd_grid **gp, struct cd_grid *gsrc)
{
   struct cd_grid *g = (struct cd_grid *) malloc(sizeof(struct cd_grid));
   if (!g) return -1;
   g->n = gsrc->n;
   g->ncells = gsrc->ncells;
   g->cell_size = gsrc->cell_size;
   g->data = 0;
   g->lengths = 0;
   g->sizes = (int *) malloc(gsrc->n * sizeof(int));
   if (!g->sizes) { free(g); return -1; }
   memcpy(g->sizes, gsrc->sizes, gsrc->n * sizeof(int));
   if (g->cell_size > 0)
   {
      g->data = (char *) malloc(g->ncells * g->cell_size);
      if (!g->data) { cd_grid_destroy(g); return -2; }
      memcpy(g->data, gsrc->data, g->ncells * g->cell_size);
   }
   g->lengths = (double *) malloc(gsrc->n * sizeof(double));
   if (!g->lengths) { cd_grid_destroy(g); return -1; }
   memcpy(g->lengths, gsrc->lengths, gsrc->n * sizeof(double));
   *gp = g;
   return 0;
}

/* grid.c:134-143 */
int cd_grid_destroy(struct cd_grid *g)
{
   if (!g) return 0;
   free(g->data);
   free(g->sizes);
   free(g->lengths);
   free(g);
   return 0;
}

/* grid.c:145-158: last subscript varies fastest */
int cd_grid_index_to_subs(struct cd_grid *g, size_t index, int *subs)
{
   int d;
   for (d = g->n - 1; d >= 0; d--)
   {
      subs[d] = (int) (index % (size_t) g->sizes[d]);
      index /= (size_t) g->sizes[d];
   }
   return 0;
}

/* grid.c:160-170 */
int cd_grid_index_from_subs(struct cd_grid *g, size_t *index, int *subs)
{
   int d;
   size_t acc = (size_t) subs[0];
   for (d = 1; d < g->n; d++) acc = acc * (size_t) g->sizes[d] + (size_t) subs[d];
   *index = acc;
   return 0;
}

/* grid.c:172-189: centre = ((0.5+sub)/size) * length, evaluated in that order */
int cd_grid_center_index(struct cd_grid *g, size_t index, double *center)
{
   int d;
   for (d = g->n - 1; d >= 0; d--)
   {
      int sub = (int) (index % (size_t) g->sizes[d]);
      index /= (size_t) g->sizes[d];
      center[d] = (0.5 + sub) / g->sizes[d];
   }
   for (d = 0; d < g->n; d++) center[d] *= g->lengths[d];
   return 0;
}

/* grid.c:191-209: returns 1 when p is outside [0,length] on any axis;
 * x == 1 maps onto the last cell */
int cd_grid_lookup_index(struct cd_grid *g, double *p, size_t *index)
{
   int d;
   size_t acc = 0;
   *index = 0;
   for (d = 0; d < g->n; d++)
   {
      double x = p[d] / g->lengths[d];
      int sub;
      if (x < 0.0) return 1;
      if (x > 1.0) return 1;
      sub = (int) floor(x * g->sizes[d]);
      if (sub == g->sizes[d]) sub--;
      acc = acc * (size_t) g->sizes[d] + (size_t) sub;
      *index = acc;
   }
   return 0;
}

/* grid.c:229-232 */
void *cd_grid_get_index(struct cd_grid *g, size_t index)
{
   return g->data + index * g->cell_size;
}

/* picks the neighbour used for the one-sided difference along one axis:
 * grid.c:352-360 (grad) and 414-423 (interp).  returns +1 (next) or -1 (prev) */
static int side_rule(int sub, int size, double p, double center)
{
   if (sub == 0) return +1;
   if (sub == size - 1) return -1;
   return (p < center) ? -1 : +1;
}

/* grid.c:331-384: piecewise-constant one-sided finite-difference gradient,
 * axes visited last-to-first while the stride accumulates; no HUGE_VAL test */
int cd_grid_double_grad(struct cd_grid *g, double *p, double *grad)
{
   size_t index, rest, stride;
   int d;
   int err = cd_grid_lookup_index(g, p, &index);
   if (err) return err;
   stride = 1;
   rest = index;
   for (d = g->n - 1; d >= 0; d--)
   {
      int size = g->sizes[d];
      int sub = (int) (rest % (size_t) size);
      double center, diff;
      rest /= (size_t) size;
      center = (0.5 + sub) / size * g->lengths[d];
      if (side_rule(sub, size, p[d], center) < 0)
      {
         diff = *(double *) cd_grid_get_index(g, index);
         diff -= *(double *) cd_grid_get_index(g, index - stride);
      }
      else
      {
         diff = *(double *) cd_grid_get_index(g, index + stride);
         diff -= *(double *) cd_grid_get_index(g, index);
      }
      grad[d] = diff * size / g->lengths[d];
      stride *= (size_t) size;
   }
   return 0;
}

/* grid.c:386-454: value at the containing cell's centre plus, per axis, the
 * one-sided slope times the offset from the centre; HUGE_VAL anywhere in the
 * stencil gives HUGE_VAL */
int cd_grid_double_interp(struct cd_grid *g, double *p, double *valuep)
{
   size_t index, rest, stride;
   int d;
   double value;
   int err = cd_grid_lookup_index(g, p, &index);
   if (err) return err;
   value = *(double *) cd_grid_get_index(g, index);
   if (value == HUGE_VAL) { *valuep = HUGE_VAL; return 0; }
   stride = 1;
   rest = index;
   for (d = g->n - 1; d >= 0; d--)
   {
      int size = g->sizes[d];
      int sub = (int) (rest % (size_t) size);
      double center, after, before, diff, slope;
      rest /= (size_t) size;
      center = (0.5 + sub) / size * g->lengths[d];
      if (side_rule(sub, size, p[d], center) < 0)
      {
         after = *(double *) cd_grid_get_index(g, index);
         before = *(double *) cd_grid_get_index(g, index - stride);
      }
      else
      {
         after = *(double *) cd_grid_get_index(g, index + stride);
         before = *(double *) cd_grid_get_index(g, index);
      }
      if (after == HUGE_VAL || before == HUGE_VAL) { *valuep = HUGE_VAL; return 0; }
      diff = after;
      diff -= before;
      slope = diff * size / g->lengths[d];
      value += slope * (p[d] - center);
      stride *= (size_t) size;
   }
   *valuep = value;
   return 0;
}

/* grid.c:269-329: one-dimensional squared distance transform by the lower
 * envelope of parabolas (Felzenszwalb & Huttenlocher).  func has unit spacing.
 * apex[] holds the parabola positions on the envelope, bound[] the abscissae
 * where the envelope switches parabola.  Infinite samples contribute nothing;
 * a line with no finite sample is all HUGE_VAL. */
static void envelope_1d(int n, const double *func, double *out, size_t out_stride,
                        int *apex, double *bound)
{
   int q, k, count = 0;
   for (q = 0; q < n; q++)
   {
      double s;
      if (func[q] == HUGE_VAL) continue;
      if (count == 0)
      {
         apex[0] = q;
         bound[0] = -HUGE_VAL;
         bound[1] = HUGE_VAL;
         count = 1;
         continue;
      }
      for (;;)
      {
         int v = apex[count - 1];
         s = func[q] + q * q;
         s -= func[v] + v * v;
         s /= 2.0 * (q - v);
         if (s <= bound[count - 1]) count--;
         else break;
      }
      apex[count] = q;
      bound[count] = s;
      bound[count + 1] = HUGE_VAL;
      count++;
   }
   if (count == 0)
   {
      for (q = 0; q < n; q++) out[(size_t) q * out_stride] = HUGE_VAL;
      return;
   }
   k = 0;
   for (q = 0; q < n; q++)
   {
      while (bound[k + 1] < q) k++;
      out[(size_t) q * out_stride] = pow(q - apex[k], 2.0) + func[apex[k]];
   }
}

/* grid.c:462-569: exact squared Euclidean distance transform of a sampled
 * function, one separable pass per dimension starting with dimension 0; each
 * line is divided by the squared cell pitch on the way in and multiplied back
 * on the way out (532-534) */
int cd_grid_double_dt_sqeuc(struct cd_grid **gp_dt, struct cd_grid *g_func)
{
   struct cd_grid *g = 0;
   int n = g_func->n;
   int d;
   if (cd_grid_create_copy(&g, g_func)) return -1;
   for (d = 0; d < n; d++)
   {
      int len = g->sizes[d];
      size_t stride = 1, outer, inner, o, i;
      int d2, q;
      double pitch2 = pow(g_func->lengths[d] / g->sizes[d], 2.0);
      int *apex = (int *) malloc(len * sizeof(int));
      double *bound = (double *) malloc((len + 1) * sizeof(double));
      double *line = (double *) malloc(len * sizeof(double));
      if (!apex || !bound || !line)
      {
         free(apex); free(bound); free(line); cd_grid_destroy(g); return -1;
      }
      for (d2 = d + 1; d2 < n; d2++) stride *= (size_t) g->sizes[d2];
      inner = stride;
      outer = g->ncells / (inner * (size_t) len);
      for (o = 0; o < outer; o++)
         for (i = 0; i < inner; i++)
         {
            double *base = (double *) g->data + o * inner * (size_t) len + i;
            for (q = 0; q < len; q++) line[q] = base[(size_t) q * stride] / pitch2;
            envelope_1d(len, line, base, stride, apex, bound);
            for (q = 0; q < len; q++) base[(size_t) q * stride] *= pitch2;
         }
      free(apex);
      free(bound);
      free(line);
   }
   *gp_dt = g;
   return 0;
}

/* grid.c:637-687: signed distance field from a grid that is 0.0 in free space
 * and HUGE_VAL inside obstacles: sqrt(sedt of the complement) - sqrt(sedt of
 * the input), i.e. positive outside obstacles and negative inside */
int cd_grid_double_bin_sdf(struct cd_grid **gp_dt, struct cd_grid *g_emp)
{
   struct cd_grid *g_obs = 0, *d_emp = 0, *d_obs = 0, *out = 0;
   size_t i;
   int ret = 0;
   if (g_emp->cell_size != (int) sizeof(double)) return -2;
   if (cd_grid_create_copy(&g_obs, g_emp)) return -1;
   for (i = 0; i < g_emp->ncells; i++)
      ((double *) g_obs->data)[i] = (((double *) g_emp->data)[i] == 0.0) ? HUGE_VAL : 0.0;
   if (cd_grid_double_dt_sqeuc(&d_emp, g_emp)) { ret = -1; goto done; }
   if (cd_grid_double_dt_sqeuc(&d_obs, g_obs)) { ret = -1; goto done; }
   if (cd_grid_create_copy(&out, d_obs)) { ret = -1; goto done; }
   for (i = 0; i < out->ncells; i++)
      ((double *) out->data)[i] = sqrt(((double *) out->data)[i]) - sqrt(((double *) d_emp->data)[i]);
done:
   cd_grid_destroy(g_obs);
   cd_grid_destroy(d_emp);
   cd_grid_destroy(d_obs);
   if (ret == 0) *gp_dt = out;
   return ret;
}

/* grid_flood.c:30-111: flood fill over face neighbours (2n-connected, despite
 * the header's remark about diagonals), optional wrap per dimension; a cell is
 * expanded only when replace() accepts it.  The visiting order does not affect
 * the result for a pure predicate/relabel callback, so an array stack is used. */
int cd_grid_flood_fill(struct cd_grid *g, size_t index_start, int *wrap_dim,
                       int (*replace)(void *, void *), void *rptr)
{
   size_t cap = 1024, top = 0;
   size_t *stack = (size_t *) malloc(cap * sizeof(size_t));
   int *subs = (int *) malloc(g->n * sizeof(int));
   if (!stack || !subs) { free(stack); free(subs); return -1; }
   stack[top++] = index_start;
   while (top)
   {
      size_t index = stack[--top];
      int d, dir;
      if (!replace(cd_grid_get_index(g, index), rptr)) continue;
      cd_grid_index_to_subs(g, index, subs);
      for (d = 0; d < g->n; d++)
         for (dir = -1; dir <= 1; dir += 2)
         {
            int keep = subs[d];
            int s = keep + dir;
            size_t nb;
            if (s < 0 || s >= g->sizes[d])
            {
               if (!wrap_dim || !wrap_dim[d]) continue;
               if (s < 0) s += g->sizes[d];
               else s -= g->sizes[d];
            }
            subs[d] = s;
            cd_grid_index_from_subs(g, &nb, subs);
            subs[d] = keep;
            if (top == cap)
            {
               size_t *bigger = (size_t *) realloc(stack, 2 * cap * sizeof(size_t));
               if (!bigger) { free(stack); free(subs); return -1; }
               stack = bigger;
               cap *= 2;
            }
            stack[top++] = nb;
         }
   }
   free(stack);
   free(subs);
   return 0;
}

/* ====================================================================== kin */

/* kin.c:42-52 */
int cd_kin_pose_identity(double pose[7])
{
   int i;
   for (i = 0; i < 6; i++) pose[i] = 0.0;
   pose[6] = 1.0;
   return 0;
}

/* kin.c:64-70 (dnrm2 + dscal on the quaternion part) */
int cd_kin_pose_normalize(double pose[7])
{
   double len = sqrt(pose[3] * pose[3] + pose[4] * pose[4] + pose[5] * pose[5] + pose[6] * pose[6]);
   double inv = 1.0 / len;
   int i;
   for (i = 3; i < 7; i++) pose[i] *= inv;
   return 0;
}

/* rotate v by the (not re-normalised) quaternion q = [x y z w] using the
 * expanded products exactly as written at kin.c:204-210 / 262-268 */
static void quat_rotate(const double *q, const double in[3], double out[3])
{
   double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
   double x = in[0], y = in[1], z = in[2];
   double qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz, qw2 = qw * qw;
   double qxqy = qx * qy, qxqz = qx * qz, qxqw = qx * qw;
   double qyqz = qy * qz, qyqw = qy * qw, qzqw = qz * qw;
   out[0] = x * (qx2 - qy2 - qz2 + qw2) + 2 * y * (qxqy - qzqw) + 2 * z * (qxqz + qyqw);
   out[1] = 2 * x * (qxqy + qzqw) + y * (-qx2 + qy2 - qz2 + qw2) + 2 * z * (qyqz - qxqw);
   out[2] = 2 * x * (qxqz - qyqw) + 2 * y * (qyqz + qxqw) + z * (-qx2 - qy2 + qz2 + qw2);
}

/* kin.c:180-212: point transform p_a = R(q_ab) p_b + t_ab */
int cd_kin_pose_compos(const double pose_ab[7], const double pos_bc[3], double pos_ac[3])
{
   double r[3];
   quat_rotate(pose_ab + 3, pos_bc, r);
   pos_ac[0] = r[0] + pose_ab[0];
   pos_ac[1] = r[1] + pose_ab[1];
   pos_ac[2] = r[2] + pose_ab[2];
   return 0;
}

/* kin.c:244-271: rotate a free vector (output may alias the input) */
int cd_kin_pose_compose_vec(const double pose_ab[7], const double vec_bc[3], double vec_ac[3])
{
   double r[3];
   quat_rotate(pose_ab + 3, vec_bc, r);
   vec_ac[0] = r[0];
   vec_ac[1] = r[1];
   vec_ac[2] = r[2];
   return 0;
}

/* kin.c:136-178: pose composition; quaternion product first (from the inputs),
 * then the rotated translation (output may alias either input) */
int cd_kin_pose_compose(const double pose_ab[7], const double pose_bc[7], double pose_ac[7])
{
   double ax = pose_ab[3], ay = pose_ab[4], az = pose_ab[5], aw = pose_ab[6];
   double bx = pose_bc[3], by = pose_bc[4], bz = pose_bc[5], bw = pose_bc[6];
   double tab[3], qab[4], r[3], q[4];
   tab[0] = pose_ab[0]; tab[1] = pose_ab[1]; tab[2] = pose_ab[2];
   qab[0] = ax; qab[1] = ay; qab[2] = az; qab[3] = aw;
   q[0] = aw * bx + ax * bw + ay * bz - az * by;
   q[1] = aw * by - ax * bz + ay * bw + az * bx;
   q[2] = aw * bz + ax * by - ay * bx + az * bw;
   q[3] = aw * bw - ax * bx - ay * by - az * bz;
   quat_rotate(qab, pose_bc, r);
   pose_ac[0] = r[0] + tab[0];
   pose_ac[1] = r[1] + tab[1];
   pose_ac[2] = r[2] + tab[2];
   pose_ac[3] = q[0];
   pose_ac[4] = q[1];
   pose_ac[5] = q[2];
   pose_ac[6] = q[3];
   return 0;
}

/* kin.c:288-326: inverse of a pose with a unit quaternion */
int cd_kin_pose_invert(const double pose_in[7], double pose_out[7])
{
   double q[4], r[3];
   q[0] = -pose_in[3];
   q[1] = -pose_in[4];
   q[2] = -pose_in[5];
   q[3] = pose_in[6];
   quat_rotate(q, pose_in, r);
   pose_out[0] = -r[0];
   pose_out[1] = -r[1];
   pose_out[2] = -r[2];
   pose_out[3] = q[0];
   pose_out[4] = q[1];
   pose_out[5] = q[2];
   pose_out[6] = q[3];
   return 0;
}

/* kin.c:347-370: rotation matrix of a quaternion [x y z w], the 1 - 2(..) form */
int cd_kin_quat_to_R(const double quat[4], double R[3][3])
{
   double x = quat[0], y = quat[1], z = quat[2], w = quat[3];
   double xx = x * x, xy = x * y, xz = x * z, xw = x * w;
   double yy = y * y, yz = y * z, yw = y * w, zz = z * z, zw = z * w;
   R[0][0] = 1 - 2 * (yy + zz); R[0][1] = 2 * (xy - zw);     R[0][2] = 2 * (xz + yw);
   R[1][0] = 2 * (xy + zw);     R[1][1] = 1 - 2 * (xx + zz); R[1][2] = 2 * (yz - xw);
   R[2][0] = 2 * (xz - yw);     R[2][1] = 2 * (yz + xw);     R[2][2] = 1 - 2 * (xx + yy);
   return 0;
}

/* kin.c:615-647: [x y z yaw pitch roll] of a pose; near the poles (|sin pitch| > 0.99998) the yaw
 * carries the whole in-plane rotation and the roll is zero */
int cd_kin_pose_to_xyzypr(const double pose[7], double xyzypr[6])
{
   const double quarter_turn = 0.25 * 6.2831853071795864769252867665590057683943387987502116;
   double qx = pose[3], qy = pose[4], qz = pose[5], qw = pose[6];
   double half_sinp = qw * qy - qz * qx;
   xyzypr[0] = pose[0];
   xyzypr[1] = pose[1];
   xyzypr[2] = pose[2];
   if (half_sinp > 0.49999)
   {
      xyzypr[3] = -2.0 * atan2(qx, qw);
      xyzypr[4] = quarter_turn;
      xyzypr[5] = 0.0;
   }
   else if (half_sinp < -0.49999)
   {
      xyzypr[3] = 2.0 * atan2(qx, qw);
      xyzypr[4] = -quarter_turn;
      xyzypr[5] = 0.0;
   }
   else
   {
      xyzypr[3] = atan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz));
      xyzypr[4] = asin(2.0 * half_sinp);
      xyzypr[5] = atan2(2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx * qx + qy * qy));
   }
   return 0;
}

/* kin.c:680-718: derivative of that vector with respect to the seven pose entries (the general
 * branch only: the reference does not treat the poles) */
int cd_kin_pose_to_xyzypr_J(const double pose[7], double J[6][7])
{
   double qx = pose[3], qy = pose[4], qz = pose[5], qw = pose[6];
   double nu, de, den, dn, nn, as, s;
   int i, j;
   for (i = 0; i < 6; i++) for (j = 0; j < 7; j++) J[i][j] = 0.0;
   J[0][0] = J[1][1] = J[2][2] = 1.0;
   /* an angle atan2(nu, de) has the gradient dn grad(nu) - nn grad(de), dn = de / (de^2 + nu^2), nn = nu / (de^2 + nu^2);
    * the factors are applied in the reference's order, so the entries round as its own do */
   nu = 2.0 * (qw * qz + qx * qy);          /* yaw */
   de = 1.0 - 2.0 * (qy * qy + qz * qz);
   den = de * de + nu * nu;
   dn = de / den;
   nn = nu / den;
   J[3][3] = dn * (2.0 * qy);
   J[3][4] = dn * (2.0 * qx) - nn * (-4.0 * qy);
   J[3][5] = dn * (2.0 * qw) - nn * (-4.0 * qz);
   J[3][6] = dn * (2.0 * qz);
   as = 2.0 * (qw * qy - qz * qx);          /* pitch = asin(as) */
   s = 1.0 / sqrt(1.0 - as * as);
   J[4][3] = s * 2.0 * (-qz);
   J[4][4] = s * 2.0 * (qw);
   J[4][5] = s * 2.0 * (-qx);
   J[4][6] = s * 2.0 * (qy);
   nu = 2.0 * (qw * qx + qy * qz);          /* roll */
   de = 1.0 - 2.0 * (qx * qx + qy * qy);
   den = de * de + nu * nu;
   dn = de / den;
   nn = nu / den;
   J[5][3] = dn * (2.0 * qw) - nn * (-4.0 * qx);
   J[5][4] = dn * (2.0 * qz) - nn * (-4.0 * qy);
   J[5][5] = dn * (2.0 * qy);
   J[5][6] = dn * (2.0 * qx);
   return 0;
}

/* spatial.c:71-102: 6 x 6 motion transform [R 0; [t]x R  R] of a pose (angular rows first) */
int cd_spatial_xm_from_pose(double xm[6][6], double pose[7])
{
   double R[3][3], tx[3][3];
   int i, j, k;
   for (i = 0; i < 6; i++) for (j = 0; j < 6; j++) xm[i][j] = 0.0;
   cd_kin_quat_to_R(pose + 3, R);
   for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) xm[i][j] = xm[3 + i][3 + j] = R[i][j];
   tx[0][0] = 0.0;      tx[0][1] = -pose[2]; tx[0][2] = pose[1];
   tx[1][0] = pose[2];  tx[1][1] = 0.0;      tx[1][2] = -pose[0];
   tx[2][0] = -pose[1]; tx[2][1] = pose[0];  tx[2][2] = 0.0;
   for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++)
      {
         double acc = 0.0;
         for (k = 0; k < 3; k++) acc += tx[i][k] * R[k][j];
         xm[3 + i][j] = acc;
      }
   return 0;
}

/* spatial.c:295-337: pose rates -> spatial velocity [omega; v of the point at the origin] */
int cd_spatial_pose_jac(double pose[7], double jac[6][7])
{
   double x = pose[0], y = pose[1], z = pose[2];
   double ax = 2.0 * pose[3], ay = 2.0 * pose[4], az = 2.0 * pose[5], aw = 2.0 * pose[6];
   int i, j;
   for (i = 0; i < 6; i++) for (j = 0; j < 7; j++) jac[i][j] = 0.0;
   jac[3][0] = jac[4][1] = jac[5][2] = 1.0;
   jac[0][3] = aw;  jac[0][4] = -az; jac[0][5] = ay;  jac[0][6] = -ax;
   jac[1][3] = az;  jac[1][4] = aw;  jac[1][5] = -ax; jac[1][6] = -ay;
   jac[2][3] = -ay; jac[2][4] = ax;  jac[2][5] = aw;  jac[2][6] = -az;
   jac[3][3] = -z * az - y * ay; jac[3][4] = -z * aw + y * ax; jac[3][5] = z * ax + y * aw;  jac[3][6] = z * ay - y * az;
   jac[4][3] = z * aw + x * ay;  jac[4][4] = -z * az - x * ax; jac[4][5] = z * ay - x * aw;  jac[4][6] = -z * ax + x * az;
   jac[5][3] = -y * aw + x * az; jac[5][4] = y * az + x * aw;  jac[5][5] = -y * ay - x * ax; jac[5][6] = y * ax - x * ay;
   return 0;
}

/* spatial.c:339-375: spatial velocity -> pose rates: p' = v + omega x p, q' = 1/2 omega (x) q */
int cd_spatial_pose_jac_inverse(double pose[7], double ji[7][6])
{
   double x = pose[0], y = pose[1], z = pose[2];
   double hx = 0.5 * pose[3], hy = 0.5 * pose[4], hz = 0.5 * pose[5], hw = 0.5 * pose[6];
   int i, j;
   for (i = 0; i < 7; i++) for (j = 0; j < 6; j++) ji[i][j] = 0.0;
   ji[0][1] = z;  ji[0][2] = -y;
   ji[1][0] = -z; ji[1][2] = x;
   ji[2][0] = y;  ji[2][1] = -x;
   ji[0][3] = ji[1][4] = ji[2][5] = 1.0;
   ji[3][0] = hw;  ji[3][1] = hz;  ji[3][2] = -hy;
   ji[4][0] = -hz; ji[4][1] = hw;  ji[4][2] = hx;
   ji[5][0] = hy;  ji[5][1] = -hx; ji[5][2] = hw;
   ji[6][0] = -hx; ji[6][1] = -hy; ji[6][2] = -hz;
   return 0;
}

/* ==================================================================== chomp */

/* chomp.c:180-217 */
void cd_chomp_free(struct cd_chomp *c)
{
   if (!c) return;
   free(c->cons_h); free(c->cons_Jcol); free(c->cons_JAJT); free(c->cons_ipiv); free(c->cons_delta);
   while (c->cons)
   {
      struct cd_chomp_con *con = c->cons;
      c->cons = con->next;
      free(con);
   }
   free(c->wds); free(c->initsfinals); free(c->inits); free(c->finals);
   free(c->A); free(c->Ainv); free(c->B);
   free(c->jlimit_lower); free(c->jlimit_upper);
   free(c->Gjlimit); free(c->GjlimitAinv);
   free(c->cost_nxn); free(c->cost_mxn);
   free(c->Kvels); free(c->Evels); free(c->vels);
   free(c->G); free(c->G_points); free(c->AG); free(c->AG_points); free(c->T_points);
   free(c);
}

/* chomp.c:40-178: allocate, defaults lambda=1, dt=1/(m+1), wds=[0..0,1],
 * limits +-inf, AG=0, inits/finals point at zero vectors, leapfrog_first=1 */
int cd_chomp_create(struct cd_chomp **cp, int m, int n, int D, double *T, int ldt)
{
   int i;
   struct cd_chomp *c = (struct cd_chomp *) calloc(1, sizeof(struct cd_chomp));
   if (!c) return -1;
   c->n = n;
   c->m = m;
   c->lambda = 1.0;
   c->dt = 1.0 / (m + 1);
   c->T = T;
   c->ldt = ldt;
   c->D = D;
   c->leapfrog_first = 1;
   c->T_points = (double **) malloc(m * sizeof(double *));
   c->G = (double *) malloc((size_t) m * n * sizeof(double));
   c->G_points = (double **) malloc(m * sizeof(double *));
   c->AG = (double *) malloc((size_t) m * n * sizeof(double));
   c->AG_points = (double **) malloc(m * sizeof(double *));
   if (D) c->wds = (double *) malloc(D * sizeof(double));
   c->initsfinals = (double *) malloc((size_t) (2 * D) * n * sizeof(double *));
   c->inits = (double **) malloc(D * sizeof(double *));
   c->finals = (double **) malloc(D * sizeof(double *));
   c->A = (double *) malloc((size_t) m * m * sizeof(double));
   c->Ainv = (double *) malloc((size_t) m * m * sizeof(double));
   c->B = (double *) malloc((size_t) m * n * sizeof(double));
   c->cost_nxn = (double *) malloc((size_t) n * n * sizeof(double));
   c->cost_mxn = (double *) malloc((size_t) m * n * sizeof(double));
   c->vels = (double *) malloc((size_t) m * n * sizeof(double));
   c->jlimit_lower = (double *) malloc(n * sizeof(double));
   c->jlimit_upper = (double *) malloc(n * sizeof(double));
   c->Gjlimit = (double *) malloc((size_t) m * n * sizeof(double));
   c->GjlimitAinv = (double *) malloc((size_t) m * n * sizeof(double));
   c->Kvels = (double *) malloc((size_t) m * m * sizeof(double));
   c->Evels = (double *) malloc((size_t) m * n * sizeof(double));
   if (!c->T_points || !c->G || !c->G_points || !c->AG || !c->AG_points || (D && !c->wds) ||
       !c->initsfinals || !c->inits || !c->finals || !c->A || !c->Ainv || !c->B ||
       !c->cost_nxn || !c->cost_mxn || !c->vels || !c->jlimit_lower || !c->jlimit_upper ||
       !c->Gjlimit || !c->GjlimitAinv || !c->Kvels || !c->Evels)
   {
      cd_chomp_free(c);
      return -1;
   }
   for (i = 0; i < m; i++)
   {
      c->T_points[i] = &T[i * ldt];
      c->G_points[i] = &c->G[i * n];
      c->AG_points[i] = &c->AG[i * n];
   }
   zero_fill(c->AG, (size_t) m * n);
   for (i = 0; i < D; i++) c->wds[i] = (i < D - 1) ? 0.0 : 1.0;
   zero_fill(c->initsfinals, (size_t) 2 * D * n);
   for (i = 0; i < D; i++)
   {
      c->inits[i] = &c->initsfinals[(2 * i) * n];
      c->finals[i] = &c->initsfinals[(2 * i + 1) * n];
   }
   zero_fill(c->A, (size_t) m * m);
   zero_fill(c->B, (size_t) m * n);
   c->trC = 0.0;
   for (i = 0; i < n; i++)
   {
      c->jlimit_lower[i] = -HUGE_VAL;
      c->jlimit_upper[i] = HUGE_VAL;
   }
   *cp = c;
   return 0;
}

/* chomp.c:239-340: smoothness metric from stacked finite-difference operators.
 * K_d (N_d x m) and E_d (N_d x n) give the d-th derivative samples K_d T + E_d;
 * A = sum_d w_d/N_d K_d^T K_d, B = sum_d w_d/N_d K_d^T E_d,
 * trC = 1/2 tr(sum_d w_d/N_d E_d^T E_d).  A row for the start (end) boundary is
 * present when inits[d] (finals[d]) is non-null. */
static int add_default_metric(struct cd_chomp *c)
{
   int m = c->m, n = c->n, D = c->D;
   int d, i, ret = 0;
   int *rows = (int *) malloc((D + 1) * sizeof(int));
   double **K = (double **) calloc(D ? D : 1, sizeof(double *));
   double **E = (double **) calloc(D ? D : 1, sizeof(double *));
   if (!rows || !K || !E) { free(rows); free(K); free(E); return -1; }
   rows[0] = m; /* rows[d+1] = N_d */
   for (d = 0; d < D; d++)
   {
      int prev = rows[d];
      int has_i = c->inits[d] ? 1 : 0;
      int has_f = c->finals[d] ? 1 : 0;
      int cur = prev - 1 + has_i + has_f;
      double *diff;
      rows[d + 1] = cur;
      K[d] = (double *) malloc((size_t) cur * m * sizeof(double));
      E[d] = (double *) malloc((size_t) cur * n * sizeof(double));
      diff = (double *) malloc((size_t) cur * prev * sizeof(double));
      if (!K[d] || !E[d] || !diff) { free(diff); ret = -1; goto done; }
      zero_fill(diff, (size_t) cur * prev);
      zero_fill(E[d], (size_t) cur * n);
      if (has_i)
      {
         diff[0] = 1.0 / c->dt;
         for (i = 0; i < n; i++) E[d][i] += (-1.0 / c->dt) * c->inits[d][i];
      }
      for (i = 0; i < prev - 1; i++)
      {
         diff[(has_i + i) * prev + i] = -1.0 / c->dt;
         diff[(has_i + i) * prev + i + 1] = 1.0 / c->dt;
      }
      if (has_f)
      {
         diff[(cur - 1) * prev + (prev - 1)] = -1.0 / c->dt;
         for (i = 0; i < n; i++) E[d][(cur - 1) * n + i] += (1.0 / c->dt) * c->finals[d][i];
      }
      if (d == 0)
         memcpy(K[d], diff, (size_t) cur * prev * sizeof(double));
      else
      {
         gemm_rm(0, 0, cur, m, prev, 1.0, diff, prev, K[d - 1], m, 0.0, K[d], m);
         gemm_rm(0, 0, cur, n, prev, 1.0, diff, prev, E[d - 1], n, 1.0, E[d], n);
      }
      free(diff);
   }
   zero_fill(c->A, (size_t) m * m);
   zero_fill(c->B, (size_t) m * n);
   zero_fill(c->cost_nxn, (size_t) n * n);
   for (d = 0; d < D; d++)
   {
      double w = c->wds[d] / rows[d + 1];
      gemm_rm(1, 0, m, m, rows[d + 1], w, K[d], m, K[d], m, 1.0, c->A, m);
      gemm_rm(1, 0, m, n, rows[d + 1], w, K[d], m, E[d], n, 1.0, c->B, n);
      gemm_rm(1, 0, n, n, rows[d + 1], w, E[d], n, E[d], n, 1.0, c->cost_nxn, n);
   }
   c->trC = 0.0;
   for (i = 0; i < n; i++) c->trC += c->cost_nxn[i * n + i];
   c->trC *= 0.5;
done:
   for (d = 0; d < D; d++) { free(K[d]); free(E[d]); }
   free(K);
   free(E);
   free(rows);
   return ret;
}

/* chomp.c:342-428: velocity operator (central differences, one-sided at a free
 * boundary), default metric, explicit inverse of A, constraint buffers. */
int cd_chomp_init(struct cd_chomp *c)
{
   int m = c->m, n = c->n, i, j;
   struct cd_chomp_con *con;
   zero_fill(c->Kvels, (size_t) m * m);
   zero_fill(c->Evels, (size_t) m * n);
   for (i = 0; i < m; i++)
   {
      if (i == 0)
      {
         if (c->inits[0])
         {
            c->Kvels[1] = 0.5 / c->dt;
            for (j = 0; j < n; j++) c->Evels[j] = c->inits[0][j] * (-0.5 / c->dt);
         }
         else
         {
            c->Kvels[1] = 1.0 / c->dt;
            c->Kvels[0] = -1.0 / c->dt;
         }
      }
      else if (i < m - 1)
      {
         c->Kvels[i * m + i + 1] = 0.5 / c->dt;
         c->Kvels[i * m + i - 1] = -0.5 / c->dt;
      }
      else
      {
         if (c->finals[0])
         {
            for (j = 0; j < n; j++) c->Evels[i * n + j] = c->finals[0][j] * (0.5 / c->dt);
            c->Kvels[i * m + i - 1] = -0.5 / c->dt;
         }
         else
         {
            c->Kvels[i * m + i] = 1.0 / c->dt;
            c->Kvels[i * m + i - 1] = -1.0 / c->dt;
         }
      }
   }
   if (add_default_metric(c)) return -1;
   memcpy(c->Ainv, c->A, (size_t) m * m * sizeof(double));
   if (invert_rm(c->Ainv, m)) return -2;
   /* chomp.c:405-426: one stacked h and J for the whole constraint list, in list order */
   c->cons_k = 0;
   for (con = c->cons; con; con = con->next) c->cons_k += con->k;
   if (c->cons_k)
   {
      int K = c->cons_k;
      c->cons_h = (double *) malloc(K * sizeof(double));
      c->cons_Jcol = (double *) malloc((size_t) K * n * sizeof(double));
      c->cons_JAJT = (double *) malloc((size_t) K * K * sizeof(double));
      c->cons_ipiv = (int *) malloc(K * sizeof(int));
      c->cons_delta = (double *) malloc(n * sizeof(double));
      if (!c->cons_h || !c->cons_Jcol || !c->cons_JAJT || !c->cons_ipiv || !c->cons_delta) return -1;
      K = 0;
      for (con = c->cons; con; con = con->next)
      {
         con->h = c->cons_h + K;
         con->J = c->cons_Jcol + (size_t) K * n;
         K += con->k;
      }
   }
   return 0;
}

/* chomp.c:219-236: new constraints go to the front of the list */
int cd_chomp_add_constraint(struct cd_chomp *c, int k, int i, void *cptr,
   int (*con_eval)(void *cptr, struct cd_chomp *c, int i, double *point, double *con_val, double *con_jacobian))
{
   struct cd_chomp_con *con = (struct cd_chomp_con *) malloc(sizeof(struct cd_chomp_con));
   if (!con) return -1;
   con->k = k;
   con->i = i;
   con->cptr = cptr;
   con->con_eval = con_eval;
   con->h = 0;
   con->J = 0;
   con->next = c->cons;
   c->cons = con;
   return 0;
}

/* stands in for LAPACKE_dgesv with one right-hand side (chomp.c:579-581): LU with partial
 * pivoting on rows, then the two triangular solves; a (K x K, row-major) is overwritten */
static int solve_rm(double *a, int K, double *b)
{
   int i, j, k;
   for (k = 0; k < K; k++)
   {
      int piv = k;
      double best = fabs(a[k * K + k]);
      for (i = k + 1; i < K; i++)
         if (fabs(a[i * K + k]) > best) { best = fabs(a[i * K + k]); piv = i; }
      if (best == 0.0) return k + 1;
      if (piv != k)
      {
         double t;
         for (j = 0; j < K; j++) { t = a[k * K + j]; a[k * K + j] = a[piv * K + j]; a[piv * K + j] = t; }
         t = b[k]; b[k] = b[piv]; b[piv] = t;
      }
      for (i = k + 1; i < K; i++)
      {
         double f = a[i * K + k] / a[k * K + k];
         if (f == 0.0) continue;
         for (j = k + 1; j < K; j++) a[i * K + j] -= f * a[k * K + j];
         b[i] -= f * b[k];
      }
   }
   for (k = K - 1; k >= 0; k--)
   {
      double acc = b[k];
      for (j = k + 1; j < K; j++) acc -= a[k * K + j] * b[j];
      b[k] = acc / a[k * K + k];
   }
   return 0;
}

/* chomp.c:430-683.
 *   vels = Kvels T + Evels (449-451; unused by the sphere cost)
 *   cost_pre, then per moving waypoint cost(); cost_obs = sum/m; G /= m (463-492)
 *   G += A T + B (515-522); AG = Ainv G, or the leapfrog momentum form (525-548)
 *   T -= AG/lambda (604-605); joint-limit projection loop (608-655)
 *   cost_smooth = tr(1/2 T^T A T + B^T T) + trC on the updated T (660-671) */
/* test hook (port only): largest number of joint-limit rounds any iteration needed */
int orc_debug_max_limit_rounds = 0;

int cd_chomp_iterate(struct cd_chomp *c, int do_iteration, double *costp_total,
                     double *costp_obs, double *costp_smooth)
{
   int m = c->m, n = c->n, i, j, round;
   double cost_point = 0.0, cost_obs = 0.0, cost_smooth = 0.0;
   int want_cost = (costp_total || costp_obs);

   memcpy(c->vels, c->Evels, (size_t) m * n * sizeof(double));
   gemm_rm(0, 0, m, n, m, 1.0, c->Kvels, m, c->T, c->ldt, 1.0, c->vels, n);

   if (c->cost_pre) c->cost_pre(c->cptr, c, m, c->T_points);
   if (do_iteration) zero_fill(c->G, (size_t) m * n);
   if (c->cost)
      for (i = 0; i < m; i++)
      {
         c->cost(c->cptr, c, i, c->T_points[i], &c->vels[i * n],
                 want_cost ? &cost_point : 0, do_iteration ? c->G_points[i] : 0);
         if (want_cost) cost_obs += cost_point;
      }
   if (want_cost) cost_obs /= m;
   for (i = 0; i < m * n; i++) c->G[i] *= 1.0 / m;
   if (c->cost_extra)
   {
      c->cost_extra(c->cptr, c, c->T, want_cost ? &cost_point : 0, do_iteration ? c->G : 0);
      cost_obs += cost_point;
   }

   if (do_iteration)
   {
      gemm_rm(0, 0, m, n, m, 1.0, c->A, m, c->T, c->ldt, 1.0, c->G, n);
      for (i = 0; i < m * n; i++) c->G[i] += c->B[i];
      if (!c->use_momentum)
         gemm_rm(0, 0, m, n, m, 1.0, c->Ainv, m, c->G, n, 0.0, c->AG, n);
      else if (c->leapfrog_first)
      {
         gemm_rm(0, 0, m, n, m, 0.5 / c->lambda, c->Ainv, m, c->G, n, 1.0, c->AG, n);
         c->leapfrog_first = 0;
      }
      else
         gemm_rm(0, 0, m, n, m, 1.0 / c->lambda, c->Ainv, m, c->G, n, 1.0, c->AG, n);

      /* hard constraints (chomp.c:553-600): with h the constraint values at T, J their Jacobians and
       * S = J Ainv J^T, solve S x = h - J AG / lambda and move T by -Ainv J^T x, so that the
       * linearised constraints vanish after the update below */
      if (c->cons_k)
      {
         struct cd_chomp_con *c1, *c2;
         int K = c->cons_k, r, q;
         for (c1 = c->cons; c1; c1 = c1->next)
            c1->con_eval(c1->cptr, c, c1->i, c->T_points[c1->i], c1->h, c1->J);
         for (c1 = c->cons; c1; c1 = c1->next)
            for (r = 0; r < c1->k; r++)
            {
               double acc = 0.0;
               for (j = 0; j < n; j++) acc += c1->J[r * n + j] * c->AG_points[c1->i][j];
               c1->h[r] += (-1.0 / c->lambda) * acc;
            }
         for (c1 = c->cons; c1; c1 = c1->next)
            for (c2 = c->cons; c2; c2 = c2->next)
            {
               double ainv = c->Ainv[c1->i * m + c2->i];
               double *blk = &c->cons_JAJT[(c1->h - c->cons_h) * K + (c2->h - c->cons_h)];
               for (r = 0; r < c1->k; r++)
                  for (q = 0; q < c2->k; q++)
                  {
                     double acc = 0.0;
                     for (j = 0; j < n; j++) acc += c1->J[r * n + j] * c2->J[q * n + j];
                     blk[r * K + q] = ainv * acc;
                  }
            }
         if (solve_rm(c->cons_JAJT, K, c->cons_h)) printf("constraint inversion error!\n");
         for (c1 = c->cons; c1; c1 = c1->next)
         {
            for (j = 0; j < n; j++)
            {
               double acc = 0.0;
               for (r = 0; r < c1->k; r++) acc += c1->J[r * n + j] * c1->h[r];
               c->cons_delta[j] = acc;
            }
            for (i = 0; i < m; i++)
               for (j = 0; j < n; j++) c->T[i * c->ldt + j] -= c->Ainv[i * m + c1->i] * c->cons_delta[j];
         }
      }

      for (i = 0; i < m; i++)
         for (j = 0; j < n; j++) c->T[i * c->ldt + j] += (-1.0 / c->lambda) * c->AG[i * n + j];

      for (round = 0; round < 1000; round++)
      {
         double worst = 0.0, scale;
         size_t worst_at = 0;
         zero_fill(c->Gjlimit, (size_t) m * n);
         for (i = 0; i < m; i++)
            for (j = 0; j < n; j++)
            {
               double q = c->T_points[i][j];
               if (q < c->jlimit_lower[j])
               {
                  c->Gjlimit[i * n + j] = c->jlimit_lower[j] - q;
                  if (fabs(c->Gjlimit[i * n + j]) > worst)
                  { worst = fabs(c->Gjlimit[i * n + j]); worst_at = (size_t) i * n + j; }
               }
               if (q > c->jlimit_upper[j])
               {
                  c->Gjlimit[i * n + j] = c->jlimit_upper[j] - q;
                  if (fabs(c->Gjlimit[i * n + j]) > worst)
                  { worst = fabs(c->Gjlimit[i * n + j]); worst_at = (size_t) i * n + j; }
               }
            }
         if (worst == 0.0) break;
         gemm_rm(0, 0, m, n, m, 1.0, c->Ainv, m, c->Gjlimit, n, 0.0, c->GjlimitAinv, n);
         scale = 1.01 * c->Gjlimit[worst_at] / c->GjlimitAinv[worst_at];
         /* the reference applies daxpy over m*n contiguous doubles of T (647-649),
          * i.e. it assumes ldt == n */
         for (i = 0; i < m * n; i++) c->T[i] += scale * c->GjlimitAinv[i];
      }
      if (round > orc_debug_max_limit_rounds) orc_debug_max_limit_rounds = round;
      if (!(round < 1000))
      {
         printf("ran too many joint limit fixes! aborting ...\n");
         return -1;
      }
   }

   if (costp_total || costp_smooth)
   {
      gemm_rm(0, 0, m, n, m, 1.0, c->A, m, c->T, c->ldt, 0.0, c->cost_mxn, n);
      gemm_rm(1, 0, n, n, m, 0.5, c->T, c->ldt, c->cost_mxn, n, 0.0, c->cost_nxn, n);
      gemm_rm(1, 0, n, n, m, 1.0, c->B, n, c->T, c->ldt, 1.0, c->cost_nxn, n);
      cost_smooth = 0.0;
      for (i = 0; i < n; i++) cost_smooth += c->cost_nxn[i * n + i];
      cost_smooth += c->trC;
   }
   if (costp_total) *costp_total = cost_obs + cost_smooth;
   if (costp_obs) *costp_obs = cost_obs;
   if (costp_smooth) *costp_smooth = cost_smooth;
   return 0;
}
