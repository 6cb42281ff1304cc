/* ocb_module.cpp -- host-side mirror of the reference's OpenRAVE module (C++).
 *
 * Same nine commands, same `key value` grammar, same defaults, same error texts
 * as src/orcdchomp_mod.cpp in the reference; every numeric step is delegated to
 * the engine through the public C ABI (include/orcdchomp_b200.h) -- this file
 * never touches CUDA.  See include/orcdchomp_b200_module.h for the mapping and
 * for what differs because OpenRAVE is not part of the build.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <map>
#include <mutex>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orcdchomp_b200_module.h"

namespace
{

/* the openrave_exception of the reference */
struct module_error : std::runtime_error
{
   explicit module_error(const std::string &m) : std::runtime_error(m) {}
};

[[noreturn]] void fail_engine(const char *what)
{
   throw module_error(std::string(what) + ": " + ocb_last_error());
}

/* --- tokeniser: the quoting rules of cd_util_shparse (util_shparse.c:37-128) ---
 * blanks separate arguments; ' and " open quotes that are closed by the same
 * character; outside quotes a backslash protects the next character; inside
 * double quotes it protects only " and backslash; backslash-newline vanishes. */
std::vector<std::string> shell_split(const std::string &in)
{
   std::vector<std::string> out;
   std::string cur;
   bool inarg = false;
   char quot = 0;
   for (size_t i = 0; i < in.size(); i++)
   {
      const char c = in[i];
      if (!inarg)
      {
         if (isspace((unsigned char) c)) continue;
         inarg = true;
         cur.clear();
      }
      if (!quot && isspace((unsigned char) c))
      {
         out.push_back(cur);
         inarg = false;
         continue;
      }
      if (!quot && (c == '"' || c == '\'')) { quot = c; continue; }
      if (quot && c == quot) { quot = 0; continue; }
      if ((!quot || quot == '"') && c == '\\' && i + 1 < in.size())
      {
         const char nx = in[i + 1];
         if (nx == '\n') { i++; continue; }
         if (!quot || nx == '"' || nx == '\\')
         {
            cur.push_back(nx);
            i++;
            continue;
         }
      }
      cur.push_back(c);
   }
   if (inarg) out.push_back(cur);
   return out;
}

std::vector<double> parse_doubles(const std::string &s)
{
   std::vector<double> v;
   for (const auto &tok : shell_split(s)) v.push_back(atof(tok.c_str()));
   return v;
}

/* tsr_create_parse (mod.cpp:3068-3110): "manipindex bodyandlink" then T0w's rotation matrix column by
 * column and its translation, the same for Twe, then the six [min max] bounds; the two transforms become
 * poses [x y z qx qy qz qw] */
struct Tsr
{
   double T0w[7], Twe[7], Bw[6][2];
};

void pose_from_columns(const double *v /* 9 rotation entries by column, 3 translation */, double pose[7])
{
   double R[3][3];
   for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) R[r][c] = v[3 * c + r];
   /* unit quaternion of R, the largest of the four components taken from the diagonal
    * (cd_kin_quat_from_R, src/libcd/kin.c:426-467) */
   const double x4 = 1.0 + R[0][0] - R[1][1] - R[2][2], y4 = 1.0 - R[0][0] + R[1][1] - R[2][2];
   const double z4 = 1.0 - R[0][0] - R[1][1] + R[2][2], w4 = 1.0 + R[0][0] + R[1][1] + R[2][2];
   double *q = pose + 3;
   if (x4 > y4 && x4 > z4 && x4 > w4)
   {
      q[0] = sqrt(0.25 * x4);
      const double f = 0.25 / q[0];
      q[1] = f * (R[1][0] + R[0][1]); q[2] = f * (R[0][2] + R[2][0]); q[3] = f * (R[2][1] - R[1][2]);
   }
   else if (y4 > z4 && y4 > w4)
   {
      q[1] = sqrt(0.25 * y4);
      const double f = 0.25 / q[1];
      q[0] = f * (R[1][0] + R[0][1]); q[2] = f * (R[2][1] + R[1][2]); q[3] = f * (R[0][2] - R[2][0]);
   }
   else if (z4 > w4)
   {
      q[2] = sqrt(0.25 * z4);
      const double f = 0.25 / q[2];
      q[0] = f * (R[0][2] + R[2][0]); q[1] = f * (R[2][1] + R[1][2]); q[3] = f * (R[1][0] - R[0][1]);
   }
   else
   {
      q[3] = sqrt(0.25 * w4);
      const double f = 0.25 / q[3];
      q[0] = f * (R[2][1] - R[1][2]); q[1] = f * (R[0][2] - R[2][0]); q[2] = f * (R[1][0] - R[0][1]);
   }
   pose[0] = v[9]; pose[1] = v[10]; pose[2] = v[11];
}

bool parse_tsr(const std::string &text, Tsr &t)
{
   int manipindex = 0, used = 0;
   char bodyandlink[32];
   if (sscanf(text.c_str(), "%d %31s%n", &manipindex, bodyandlink, &used) != 2) return false;
   const std::vector<double> v = parse_doubles(text.substr(used));
   if (v.size() != 36) return false;
   pose_from_columns(&v[0], t.T0w);
   pose_from_columns(&v[12], t.Twe);
   for (int i = 0; i < 6; i++) { t.Bw[i][0] = v[24 + 2 * i]; t.Bw[i][1] = v[24 + 2 * i + 1]; }
   return true;
}

/* cd_kin_pose_compose (kin.c:136-178) on the host: snapshot poses only */
void pose_compose(const double ab[7], const double bc[7], double ac[7])
{
   const double ax = ab[3], ay = ab[4], az = ab[5], aw = ab[6];
   const double bx = bc[3], by = bc[4], bz = bc[5], bw = bc[6];
   const double x = bc[0], y = bc[1], z = bc[2];
   const double qx2 = ax * ax, qy2 = ay * ay, qz2 = az * az, qw2 = aw * aw;
   const double qxqy = ax * ay, qxqz = ax * az, qxqw = ax * aw, qyqz = ay * az, qyqw = ay * aw, qzqw = az * aw;
   double r[7];
   r[0] = x * (qx2 - qy2 - qz2 + qw2) + 2 * y * (qxqy - qzqw) + 2 * z * (qxqz + qyqw) + ab[0];
   r[1] = 2 * x * (qxqy + qzqw) + y * (-qx2 + qy2 - qz2 + qw2) + 2 * z * (qyqz - qxqw) + ab[1];
   r[2] = 2 * x * (qxqz - qyqw) + 2 * y * (qyqz + qxqw) + z * (-qx2 - qy2 + qz2 + qw2) + ab[2];
   r[3] = aw * bx + ax * bw + ay * bz - az * by;
   r[4] = aw * by - ax * bz + ay * bw + az * bx;
   r[5] = aw * bz + ax * by - ay * bx + az * bw;
   r[6] = aw * bw - ax * bx - ay * by - az * bz;
   memcpy(ac, r, sizeof(r));
}

void pose_invert(const double in[7], double out[7])
{
   const double inv[7] = {0, 0, 0, -in[3], -in[4], -in[5], in[6]};
   const double neg[7] = {-in[0], -in[1], -in[2], 0, 0, 0, 1};
   pose_compose(inv, neg, out); /* R^-1 (-t) */
}

void pose_identity(double p[7])
{
   for (int i = 0; i < 6; i++) p[i] = 0.0;
   p[6] = 1.0;
}

void quat_rows(const double *q, double R[9])
{
   const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
   R[0] = qx * qx - qy * qy - qz * qz + qw * qw; R[1] = 2 * (qx * qy - qz * qw); R[2] = 2 * (qx * qz + qy * qw);
   R[3] = 2 * (qx * qy + qz * qw); R[4] = -qx * qx + qy * qy - qz * qz + qw * qw; R[5] = 2 * (qy * qz - qx * qw);
   R[6] = 2 * (qx * qz - qy * qw); R[7] = 2 * (qy * qz + qx * qw); R[8] = -qx * qx - qy * qy + qz * qz + qw * qw;
}

/* apply a rigid frame to a primitive: boxes carry a pose, spheres a centre, triangles three vertices */
void pose_apply(const double frame[7], const double p[3], double out[3])
{
   const double in[7] = {p[0], p[1], p[2], 0, 0, 0, 1};
   double r[7];
   pose_compose(frame, in, r);
   out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

ocb_prim prim_transform(const double frame[7], const ocb_prim &p)
{
   ocb_prim q = p;
   if (p.type == OCB_PRIM_TRIANGLE)
   {
      double v[9] = {p.pose[0], p.pose[1], p.pose[2], p.pose[3], p.pose[4], p.pose[5], p.pose[6], p.extents[0], p.extents[1]};
      double w[9];
      for (int i = 0; i < 3; i++) pose_apply(frame, v + 3 * i, w + 3 * i);
      for (int k = 0; k < 7; k++) q.pose[k] = w[k];
      q.extents[0] = w[7];
      q.extents[1] = w[8];
   }
   else if (p.type == OCB_PRIM_SPHERE)
   {
      double c[3];
      pose_apply(frame, p.pose, c);
      q.pose[0] = c[0]; q.pose[1] = c[1]; q.pose[2] = c[2];
   }
   else
      pose_compose(frame, p.pose, q.pose);
   return q;
}

struct Kinbody
{
   std::string name;
   double pose[7];
   bool enabled = true;
   std::vector<ocb_prim> prims; /* kinbody frame */
};

/* deep copy of an ocb_robot */
struct Robot
{
   std::string name;
   std::vector<int> parent, joint_type, dof_index, sphere_link;
   std::vector<double> pose_parent, axis, dof_coeff, limit_lower, limit_upper, sphere_pos, sphere_radius, q;
   ocb_robot desc;
   /* what con_tsr / start_tsr / everyn_tsr look up on the robot (mod.cpp:1957-1977): link names and
    * manipulators (end-effector link + local tool transform); the first manipulator added is the active one */
   struct Manip { std::string name; int link; double tool[7]; };
   std::vector<std::string> link_names;
   std::vector<Manip> manips;
   int active_manip = 0;

   void assign(const ocb_robot *r, const double *values)
   {
      const int nl = r->n_links, n = r->n_dof, ns = r->n_spheres;
      parent.assign(r->parent, r->parent + nl);
      joint_type.assign(r->joint_type, r->joint_type + nl);
      dof_index.assign(r->dof_index, r->dof_index + nl);
      pose_parent.assign(r->pose_parent, r->pose_parent + 7 * nl);
      axis.assign(r->axis, r->axis + 3 * nl);
      dof_coeff.assign(r->dof_coeff, r->dof_coeff + 2 * nl);
      limit_lower.assign(r->limit_lower, r->limit_lower + n);
      limit_upper.assign(r->limit_upper, r->limit_upper + n);
      sphere_link.assign(r->sphere_link, r->sphere_link + ns);
      sphere_pos.assign(r->sphere_pos, r->sphere_pos + 3 * ns);
      sphere_radius.assign(r->sphere_radius, r->sphere_radius + ns);
      q.assign(values, values + n);
      desc = *r;
      desc.parent = parent.data(); desc.joint_type = joint_type.data(); desc.dof_index = dof_index.data();
      desc.pose_parent = pose_parent.data(); desc.axis = axis.data(); desc.dof_coeff = dof_coeff.data();
      desc.limit_lower = limit_lower.data(); desc.limit_upper = limit_upper.data();
      desc.sphere_link = sphere_link.data(); desc.sphere_pos = sphere_pos.data(); desc.sphere_radius = sphere_radius.data();
   }
};

} /* namespace */

struct ocb_env
{
   std::map<std::string, Kinbody> kinbodies;
   std::map<std::string, std::unique_ptr<Robot>> robots;
   /* the reference holds the environment's recursive mutex for the whole of every command
    * (EnvironmentMutex::scoped_lock, mod.cpp:179, 323, 612, 806, 1829, 2745, 2894) */
   std::recursive_mutex mutex;
};

namespace
{

/* struct sdf of the reference (mod.cpp:148-153) */
struct Field
{
   std::string kinbody_name;
   double pose[7]; /* grid frame in the kinbody frame */
   int sizes[3];
   double lengths[3];
   int sdf_id = -1; /* the grid, resident in HBM (engine SDF slot); runs alias it with their own pose */
};

/* struct run of the reference (mod.cpp:887-966), R runs behind one handle */
struct Run
{
   ocb_batch *batch = nullptr;
   bool floating = false; /* rows start with the base pose (floating_base) */
   std::vector<int> sdf_ids;
   std::string robot_name;
   int n_runs = 1, n_points = 0, n_dof = 0;
   FILE *fp_dat = nullptr;
};

} /* namespace */

struct ocb_module
{
   ocb_env *env = nullptr;
   ocb_engine *engine = nullptr;
   std::vector<Field> sdfs;
   std::set<Run *> runs;

   ~ocb_module()
   {
      for (Run *r : runs)
      {
         ocb_batch_destroy(r->batch);
         for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
         if (r->fp_dat) fclose(r->fp_dat);
         delete r;
      }
      for (const Field &f : sdfs) ocb_sdf_remove(engine, f.sdf_id);
      if (engine) ocb_engine_destroy(engine);
   }

   Kinbody &kinbody(const std::string &name)
   {
      auto it = env->kinbodies.find(name);
      if (it == env->kinbodies.end()) throw module_error("Could not find kinbody with that name!");
      return it->second;
   }

   Run *run_from_handle(const char *text)
   {
      void *p = nullptr;
      if (sscanf(text, "%p", &p) != 1) throw module_error("Could not parse r!");
      Run *r = (Run *) p;
      if (!runs.count(r)) throw module_error("you must pass a created run!");
      return r;
   }

   static void bad_args(const std::vector<std::string> &argv, size_t i)
   {
      std::string rest;
      for (; i < argv.size(); i++) rest += " " + argv[i];
      throw module_error("Bad arguments!" + (rest.empty() ? std::string() : " (argument" + rest + " not known)"));
   }

   /* ---------------------------------------------------------------- commands */
   int viewspheres(const std::vector<std::string> &argv, std::ostream &sout)
   {
      std::string robot;
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "robot" && i + 1 < argv.size()) robot = argv[++i];
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      auto it = env->robots.find(robot);
      if (it == env->robots.end()) throw module_error("Did not pass all required args!");
      const Robot &r = *it->second;
      for (size_t s = 0; s < r.sphere_radius.size(); s++)
         sout << "orcdchomp_sphere_" << s << " link " << r.sphere_link[s] << " pos " << r.sphere_pos[3 * s] << " "
              << r.sphere_pos[3 * s + 1] << " " << r.sphere_pos[3 * s + 2] << " radius " << r.sphere_radius[s] << "\n";
      return 0;
   }

   int viewfields(const std::vector<std::string> &argv, std::ostream &sout)
   {
      if (argv.size() > 1) bad_args(argv, 1);
      for (const Field &f : sdfs)
         sout << f.kinbody_name << " " << f.sizes[0] << " " << f.sizes[1] << " " << f.sizes[2] << "\n";
      return 0;
   }

   void check_no_field(const std::string &name)
   {
      for (const Field &f : sdfs)
         if (f.kinbody_name == name) throw module_error("We already have an sdf for this kinbody!");
   }

   /* mod.cpp:297-589 */
   int computedistancefield(const std::vector<std::string> &argv, std::ostream &)
   {
      std::string name, cache_filename;
      bool have_kb = false, require_cache = false;
      double cube_extent = 0.02, aabb_padding = 0.2; /* mod.cpp:325-326 */
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "kinbody" && i + 1 < argv.size())
         {
            if (have_kb) throw module_error("Only one kinbody can be passed!");
            name = argv[++i];
            kinbody(name);
            have_kb = true;
         }
         else if (argv[i] == "aabb_padding" && i + 1 < argv.size()) aabb_padding = atof(argv[++i].c_str());
         else if (argv[i] == "cube_extent" && i + 1 < argv.size()) cube_extent = atof(argv[++i].c_str());
         else if (argv[i] == "cache_filename" && i + 1 < argv.size()) cache_filename = argv[++i];
         else if (argv[i] == "require_cache") require_cache = true;
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!have_kb) throw module_error("Did not pass all required args!");
      check_no_field(name);
      if (name.size() + 1 > 256) throw module_error("ugh, orcdchomp currently doesn't support long kinbody names!");
      Kinbody &kb = kinbody(name);

      /* AABB of the body at the world origin (KinBodyComputeEnabledAABB, mod.cpp:88-140, 377-381);
       * a body without geometry of its own (e.g. a disabled robot, scripts/test_wam7.py:78) gets
       * the AABB of everything enabled, expressed in its frame */
      double lo[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, hi[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
      double kb_inv[7];
      pose_invert(kb.pose, kb_inv);
      auto grow = [&](const ocb_prim &p0, const double frame[7])
      {
         const ocb_prim p = prim_transform(frame, p0);
         if (p.type == OCB_PRIM_TRIANGLE)
         {
            const double v[9] = {p.pose[0], p.pose[1], p.pose[2], p.pose[3], p.pose[4], p.pose[5], p.pose[6], p.extents[0], p.extents[1]};
            for (int i = 0; i < 3; i++)
               for (int k = 0; k < 3; k++)
               {
                  if (v[3 * i + k] < lo[k]) lo[k] = v[3 * i + k];
                  if (v[3 * i + k] > hi[k]) hi[k] = v[3 * i + k];
               }
            return;
         }
         double R[9];
         quat_rows(p.pose + 3, R);
         for (int k = 0; k < 3; k++)
         {
            double e = (p.type == OCB_PRIM_SPHERE)
                          ? p.extents[0]
                          : fabs(R[3 * k]) * p.extents[0] + fabs(R[3 * k + 1]) * p.extents[1] + fabs(R[3 * k + 2]) * p.extents[2];
            if (p.pose[k] - e < lo[k]) lo[k] = p.pose[k] - e;
            if (p.pose[k] + e > hi[k]) hi[k] = p.pose[k] + e;
         }
      };
      double ident[7];
      pose_identity(ident);
      if (!kb.prims.empty())
         for (const ocb_prim &p : kb.prims) grow(p, ident);
      else
         for (auto &kv : env->kinbodies)
         {
            if (!kv.second.enabled) continue;
            double rel[7];
            pose_compose(kb_inv, kv.second.pose, rel);
            for (const ocb_prim &p : kv.second.prims) grow(p, rel);
         }
      if (!(lo[0] <= hi[0])) throw module_error("kinbody has no geometry to compute a field for!");
      Field f;
      f.kinbody_name = name;
      double aabb_pos[3], aabb_ext[3];
      for (int k = 0; k < 3; k++)
      {
         aabb_pos[k] = 0.5 * (lo[k] + hi[k]);
         aabb_ext[k] = 0.5 * (hi[k] - lo[k]);
         f.sizes[k] = (int) ceil((aabb_ext[k] + aabb_padding) / cube_extent); /* mod.cpp:386-393 */
         f.lengths[k] = f.sizes[k] * 2.0 * cube_extent;                       /* mod.cpp:402-403 */
      }
      pose_identity(f.pose);
      for (int k = 0; k < 3; k++) f.pose[k] = aabb_pos[k] - 0.5 * f.lengths[k]; /* mod.cpp:407-409 */
      const size_t ncells = (size_t) f.sizes[0] * f.sizes[1] * f.sizes[2];
      /* the grid lives in HBM; a host copy exists only while a cache file is read or written */
      std::vector<double> host;

      bool loaded = false;
      if (!cache_filename.empty())
      {
         /* raw doubles, validated by size only (mod.cpp:416-444) */
         FILE *fp = fopen(cache_filename.c_str(), "rb");
         if (fp)
         {
            fseek(fp, 0L, SEEK_END);
            if ((size_t) ftell(fp) == ncells * sizeof(double))
            {
               fseek(fp, 0L, SEEK_SET);
               host.resize(ncells);
               loaded = fread(host.data(), sizeof(double), ncells, fp) == ncells;
            }
            fclose(fp);
         }
         if (loaded)
         {
            ocb_sdf d;
            for (int k = 0; k < 3; k++) { d.sizes[k] = f.sizes[k]; d.lengths[k] = f.lengths[k]; }
            memcpy(d.pose_world_gsdf, f.pose, sizeof(f.pose));
            d.data = host.data();
            if (ocb_sdf_upload(engine, &d, &f.sdf_id) != OCB_OK) fail_engine("Not enough memory for distance field!");
         }
      }
      if (!loaded)
      {
         if (require_cache) throw module_error("Field not found from cache, but require_cache flag set!");
         /* all enabled bodies, expressed in the grid frame (mod.cpp:483-518) */
         double world_gsdf[7], gsdf_world[7];
         pose_compose(kb.pose, f.pose, world_gsdf);
         pose_invert(world_gsdf, gsdf_world);
         std::vector<ocb_prim> prims;
         for (auto &kv : env->kinbodies)
         {
            if (!kv.second.enabled) continue;
            double rel[7];
            pose_compose(gsdf_world, kv.second.pose, rel);
            for (const ocb_prim &p : kv.second.prims) prims.push_back(prim_transform(rel, p));
         }
         if (ocb_computedistancefield_resident(engine, prims.data(), (int) prims.size(), f.sizes, f.lengths, cube_extent,
                                               f.pose, &f.sdf_id) != OCB_OK)
            fail_engine("Not enough memory for distance field!");
         if (!cache_filename.empty())
         {
            FILE *fp = fopen(cache_filename.c_str(), "wb"); /* mod.cpp:571-580 */
            if (fp)
            {
               host.resize(ncells);
               if (ocb_sdf_download(engine, f.sdf_id, host.data()) == OCB_OK) fwrite(host.data(), sizeof(double), ncells, fp);
               fclose(fp);
            }
         }
      }
      sdfs.push_back(std::move(f));
      return 0;
   }

   /* mod.cpp:592-722: takes ownership of the malloc'ed obstacle array */
   int addfield_fromobsarray(const std::vector<std::string> &argv, std::ostream &)
   {
      std::string name;
      bool have_kb = false;
      double *obsarray = nullptr;
      int sizes[3] = {0, 0, 0};
      double lengths[3] = {0, 0, 0}, pose[7];
      pose_identity(pose);
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "kinbody" && i + 1 < argv.size())
         {
            if (have_kb) throw module_error("Only one kinbody can be passed!");
            name = argv[++i];
            kinbody(name);
            have_kb = true;
         }
         else if (argv[i] == "obsarray" && i + 1 < argv.size())
         {
            void *p = nullptr;
            sscanf(argv[++i].c_str(), "%p", &p);
            obsarray = (double *) p;
         }
         else if (argv[i] == "sizes" && i + 1 < argv.size())
         {
            const auto v = shell_split(argv[++i]);
            if (v.size() != 3) throw module_error("sizes must be length 3!");
            for (int k = 0; k < 3; k++) sizes[k] = atoi(v[k].c_str());
         }
         else if (argv[i] == "lengths" && i + 1 < argv.size())
         {
            const auto v = parse_doubles(argv[++i]);
            if (v.size() != 3) throw module_error("lengths must be length 3!");
            for (int k = 0; k < 3; k++) lengths[k] = v[k];
         }
         else if (argv[i] == "pose" && i + 1 < argv.size())
         {
            const auto v = parse_doubles(argv[++i]);
            if (v.size() != 7) throw module_error("pose must be length 7!");
            for (int k = 0; k < 7; k++) pose[k] = v[k];
         }
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!have_kb) throw module_error("Did not pass a kinbody!");
      if (!obsarray) throw module_error("Did not pass an obsarray!");
      for (int k = 0; k < 3; k++)
         if (sizes[k] <= 0) throw module_error("Didn't pass non-zero sizes!");
      for (int k = 0; k < 3; k++)
         if (lengths[k] <= 0.0) throw module_error("Didn't pass non-zero lengths!");
      {
         /* cd_kin_pose_normalize (kin.c:64-70) */
         const double len = sqrt(pose[3] * pose[3] + pose[4] * pose[4] + pose[5] * pose[5] + pose[6] * pose[6]);
         for (int k = 3; k < 7; k++) pose[k] *= 1.0 / len;
      }
      check_no_field(name);
      Field f;
      f.kinbody_name = name;
      memcpy(f.pose, pose, sizeof(pose));
      for (int k = 0; k < 3; k++) { f.sizes[k] = sizes[k]; f.lengths[k] = lengths[k]; }
      const int rc = ocb_sdf_build_resident(engine, obsarray, sizes, lengths, f.pose, &f.sdf_id);
      free(obsarray); /* the reference frees it through cd_grid_destroy (mod.cpp:703-704, 714) */
      if (rc != OCB_OK) fail_engine("Not enough memory for distance field!");
      sdfs.push_back(std::move(f));
      return 0;
   }

   /* mod.cpp:799-847 */
   int removefield(const std::vector<std::string> &argv, std::ostream &)
   {
      std::string name;
      bool have_kb = false;
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "kinbody" && i + 1 < argv.size())
         {
            if (have_kb) throw module_error("Only one kinbody can be passed!");
            name = argv[++i];
            kinbody(name);
            have_kb = true;
         }
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!have_kb) throw module_error("Did not pass all required args!");
      for (size_t k = 0; k < sdfs.size(); k++)
         if (sdfs[k].kinbody_name == name)
         {
            /* runs created earlier hold aliases of this grid: they must be destroyed first, as the
             * reference's runs hold raw grid pointers (mod.cpp:2347-2369, 836) */
            ocb_sdf_remove(engine, sdfs[k].sdf_id);
            sdfs.erase(sdfs.begin() + k);
            return 0;
         }
      throw module_error("No field with that kinbody name exists!");
   }

   /* `starttraj`: an OpenRAVE trajectory in its XML serialisation (what gettraj returns).  The
    * reference deserialises it and samples the active-dof group at n_points evenly spaced times
    * over its duration (mod.cpp:2005-2012, 2375-2415); here the two groups that matter are read
    * directly -- "joint_values <robot> i0 i1 ..." (linear interpolation) and "deltatime" -- and
    * sampled the same way.  out: [n_points][n]. */
   static bool xml_attr(const std::string &tag, const char *name, std::string &val)
   {
      const std::string key = std::string(name) + "=\"";
      size_t a = tag.find(key);
      if (a == std::string::npos) return false;
      a += key.size();
      const size_t b = tag.find('"', a);
      if (b == std::string::npos) return false;
      val = tag.substr(a, b - a);
      return true;
   }

   static void sample_starttraj(const std::string &xml, int n, int n_points, std::vector<double> &out)
   {
      int jv_off = -1, jv_dof = 0, dt_off = -1, stride = 0;
      for (size_t pos = 0; (pos = xml.find("<group", pos)) != std::string::npos; pos++)
      {
         const size_t end = xml.find('>', pos);
         if (end == std::string::npos) break;
         const std::string tag = xml.substr(pos, end - pos);
         std::string name, off, dof;
         if (!xml_attr(tag, "name", name) || !xml_attr(tag, "offset", off) || !xml_attr(tag, "dof", dof)) continue;
         const int o = atoi(off.c_str()), d = atoi(dof.c_str());
         stride = std::max(stride, o + d);
         if (name.compare(0, 12, "joint_values") == 0 && jv_off < 0) { jv_off = o; jv_dof = d; }
         else if (name == "deltatime") dt_off = o;
      }
      if (jv_off < 0) throw module_error("starttraj has no joint_values group!");
      if (jv_dof != n) throw module_error("starttraj joint_values group does not match the robot's active dofs!");
      if (dt_off < 0) throw module_error("starttraj has no deltatime group (an untimed trajectory cannot be sampled)!");
      const size_t d0 = xml.find("<data");
      const size_t d1 = (d0 == std::string::npos) ? d0 : xml.find('>', d0);
      const size_t d2 = (d1 == std::string::npos) ? d1 : xml.find("</data>", d1);
      if (d2 == std::string::npos) throw module_error("starttraj has no data block!");
      std::string cnt;
      if (!xml_attr(xml.substr(d0, d1 - d0), "count", cnt)) throw module_error("starttraj data block has no count!");
      const int count = atoi(cnt.c_str());
      const std::vector<double> v = parse_doubles(xml.substr(d1 + 1, d2 - d1 - 1));
      if (count < 2 || (size_t) count * stride != v.size()) throw module_error("starttraj data block has the wrong size!");
      std::vector<double> tau(count);
      double acc = 0.0;
      for (int k = 0; k < count; k++)
      {
         const double dt = v[(size_t) k * stride + dt_off];
         if (dt < 0.0) throw module_error("starttraj has a negative deltatime!");
         acc += (k == 0) ? 0.0 : dt; /* the first waypoint's deltatime is not part of the duration */
         tau[k] = acc;
      }
      const double duration = acc;
      if (!(duration > 0.0)) throw module_error("starttraj has zero duration!");
      out.assign((size_t) n_points * n, 0.0);
      int seg = 0;
      for (int i = 0; i < n_points; i++)
      {
         const double t = i * duration / (n_points - 1); /* mod.cpp:2411 */
         while (seg < count - 2 && t > tau[seg + 1]) seg++;
         const double span = tau[seg + 1] - tau[seg];
         double w = (span > 0.0) ? (t - tau[seg]) / span : 1.0;
         if (w < 0.0) w = 0.0;
         if (w > 1.0) w = 1.0;
         for (int j = 0; j < n; j++)
         {
            const double a = v[(size_t) seg * stride + jv_off + j], b = v[(size_t) (seg + 1) * stride + jv_off + j];
            out[(size_t) i * n + j] = (w == 0.0) ? a : ((w == 1.0) ? b : a + (b - a) * w);
         }
      }
   }

   /* mod.cpp:1800-2688 (single run) + the createbatch extension */
   int create(const std::vector<std::string> &argv, std::ostream &sout, bool batch_form)
   {
      std::string robot_name, dat_filename, starttraj;
      bool have_starttraj = false;
      std::vector<double> adofgoal, basegoal;
      ocb_params pr;
      ocb_params_default(&pr);
      unsigned int seed = 0;
      int n_runs = 1;
      const double *goals_ptr = nullptr, *starts_ptr = nullptr;
      const unsigned int *seeds_ptr = nullptr;
      const char *unsupported = nullptr;
      std::vector<ocb_constraint> start_tsrs, everyn_tsrs, con_tsrs; /* a repeated start_tsr / everyn_tsr replaces the earlier one */
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         const std::string &a = argv[i];
         const bool has1 = i + 1 < argv.size();
         if (a == "robot" && has1)
         {
            if (!robot_name.empty()) throw module_error("Only one robot can be passed!");
            robot_name = argv[++i];
            if (!env->robots.count(robot_name)) throw module_error("Could not find robot that matches the given name!");
         }
         else if (a == "adofgoal" && has1)
         {
            if (!adofgoal.empty()) throw module_error("Only one adofgoal can be passed!");
            if (have_starttraj) throw module_error("Cannot pass both adofgoal and starttraj!");
            adofgoal = parse_doubles(argv[++i]);
         }
         else if (a == "starttraj" && has1)
         {
            if (have_starttraj) throw module_error("Only one starttraj can be passed!");
            if (!adofgoal.empty()) throw module_error("Cannot pass both adofgoal and starttraj!");
            if (!basegoal.empty()) throw module_error("Cannot pass both basegoal and starttraj!");
            starttraj = argv[++i];
            have_starttraj = true;
         }
         else if (a == "lambda" && has1) pr.lambda = atof(argv[++i].c_str());
         else if (a == "n_points" && has1) pr.n_points = atoi(argv[++i].c_str());
         else if (a == "derivative" && has1) pr.derivative = atoi(argv[++i].c_str());
         else if (a == "use_momentum") pr.use_momentum = 1;
         else if (a == "use_hmc") pr.use_hmc = 1;
         else if (a == "hmc_resample_lambda" && has1) pr.hmc_resample_lambda = atof(argv[++i].c_str());
         else if (a == "seed" && has1) sscanf(argv[++i].c_str(), "%u", &seed);
         else if (a == "epsilon" && has1) pr.epsilon = atof(argv[++i].c_str());
         else if (a == "epsilon_self" && has1) pr.epsilon_self = atof(argv[++i].c_str());
         else if (a == "obs_factor" && has1) pr.obs_factor = atof(argv[++i].c_str());
         else if (a == "obs_factor_self" && has1) pr.obs_factor_self = atof(argv[++i].c_str());
         else if (a == "dat_filename" && has1) dat_filename = argv[++i];
         else if ((a == "ee_force" || a == "ee_force_at" || a == "ee_torque_weights") && has1) ++i; /* dead parameters, mod.cpp:1323 */
         else if (a == "floating_base") pr.floating_base = 1;
         else if (a == "basegoal" && has1)
         {
            if (!basegoal.empty()) throw module_error("Only one basegoal can be passed!");
            if (have_starttraj) throw module_error("Cannot pass both basegoal and starttraj!");
            basegoal = parse_doubles(argv[++i]);
            if (basegoal.size() != 7) throw module_error("basegoal argument must be length 7!");
         }
         else if (a == "start_cost" && has1)
         {
            unsupported = argv[i].c_str(); /* a host callback per iteration: not on the device path */
            ++i;
         }
         else if ((a == "start_tsr" || a == "everyn_tsr") && has1)
         {
            /* mod.cpp:1988-1997: on the active manipulator's end effector */
            Tsr t;
            if (!parse_tsr(argv[++i], t)) throw module_error("Cannot parse " + a + " TSR!");
            ocb_constraint c;
            memset(&c, 0, sizeof(c));
            c.where = (a == "start_tsr") ? OCB_CON_START_TSR : OCB_CON_ALL;
            c.link = -1; /* the active manipulator, resolved once the robot is known */
            memcpy(c.T0w, t.T0w, sizeof(c.T0w)); memcpy(c.Twe, t.Twe, sizeof(c.Twe)); memcpy(c.Bw, t.Bw, sizeof(c.Bw));
            (a == "start_tsr" ? start_tsrs : everyn_tsrs).push_back(c);
         }
         else if (a == "con_tsr" && i + 2 < argv.size())
         {
            /* mod.cpp:1930-1987: 'start|end|all [manipee NAME | link NAME]' 'TSR' */
            if (robot_name.empty()) throw module_error("You must pass robot before any con_tsrs!");
            const std::vector<std::string> sub = shell_split(argv[++i]);
            if (sub.size() != 1 && sub.size() != 3) throw module_error("con_tsr first argument must be length 1 or 3!");
            ocb_constraint c;
            memset(&c, 0, sizeof(c));
            if (sub[0] == "all") c.where = OCB_CON_ALL;
            else if (sub[0] == "start") c.where = OCB_CON_START;
            else if (sub[0] == "end") c.where = OCB_CON_END;
            else throw module_error("con_tsr first arg must be start, end, or all!");
            Robot &rb0 = *env->robots[robot_name];
            c.pose_link_ee[6] = 1.0;
            if (sub.size() != 3) c.link = -1;
            else if (sub[1] == "manipee")
            {
               size_t k = 0;
               for (; k < rb0.manips.size(); k++)
                  if (rb0.manips[k].name == sub[2]) break;
               if (k == rb0.manips.size()) throw module_error("con_tsr manip not found!");
               c.link = rb0.manips[k].link;
               memcpy(c.pose_link_ee, rb0.manips[k].tool, sizeof(c.pose_link_ee));
            }
            else if (sub[1] == "link")
            {
               size_t k = 0;
               for (; k < rb0.link_names.size(); k++)
                  if (rb0.link_names[k] == sub[2]) break;
               if (k == rb0.link_names.size()) throw module_error("con_tsr link not found!");
               c.link = (int) k;
            }
            else throw module_error("con_tsr first arg must be empty, manipee, or link!");
            Tsr t;
            if (!parse_tsr(argv[++i], t)) throw module_error("Cannot parse constraint TSR!");
            memcpy(c.T0w, t.T0w, sizeof(c.T0w)); memcpy(c.Twe, t.Twe, sizeof(c.Twe)); memcpy(c.Bw, t.Bw, sizeof(c.Bw));
            con_tsrs.push_back(c);
         }
         else if (batch_form && a == "n_runs" && has1) n_runs = atoi(argv[++i].c_str());
         else if (batch_form && (a == "adofgoals" || a == "adofstarts" || a == "seeds") && has1)
         {
            void *p = nullptr;
            sscanf(argv[i + 1].c_str(), "%p", &p);
            if (a == "adofgoals") goals_ptr = (const double *) p;
            else if (a == "adofstarts") starts_ptr = (const double *) p;
            else seeds_ptr = (const unsigned int *) p;
            ++i;
         }
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (unsupported)
         throw module_error(std::string("'") + unsupported + "' is not supported by the B200 engine (see orcdchomp_b200_module.h)");
      /* validity checks in the reference's order (mod.cpp:2091-2101) */
      if (robot_name.empty()) throw module_error("Did not pass a robot!");
      if (adofgoal.empty() && !goals_ptr && !have_starttraj) throw module_error("Did not pass either adofgoal or starttraj!");
      if (pr.floating_base && basegoal.empty() && !have_starttraj) throw module_error("Passed floating_base with no basegoal!");
      if (pr.floating_base && (have_starttraj || goals_ptr || starts_ptr))
         throw module_error("floating_base with starttraj or batched end points is not supported by the B200 engine");
      if (sdfs.empty()) throw module_error("No signed distance fields have yet been computed!");
      if (pr.lambda < 0.01) throw module_error("lambda must be >=0.01!");
      if (pr.n_points < 3) throw module_error("n_points must be >=3!");
      if (pr.floating_base && !start_tsrs.empty()) throw module_error("floating_base and start_tsr together is not yet implemented!");
      Robot &rb = *env->robots[robot_name];
      /* the constraint list in the order the reference registers it (mod.cpp:2571-2613) */
      std::vector<ocb_constraint> cons;
      if (!start_tsrs.empty()) cons.push_back(start_tsrs.back());
      if (!everyn_tsrs.empty()) cons.push_back(everyn_tsrs.back());
      cons.insert(cons.end(), con_tsrs.begin(), con_tsrs.end());
      for (ocb_constraint &c : cons)
         if (c.link < 0)
         {
            /* GetActiveManipulator()->GetEndEffectorTransform() (mod.cpp:1384-1385, 1545, 1701) */
            if (rb.manips.empty()) throw module_error("robot has no active manipulator for the TSR constraint!");
            const Robot::Manip &mp = rb.manips[rb.active_manip];
            c.link = mp.link;
            memcpy(c.pose_link_ee, mp.tool, sizeof(c.pose_link_ee));
         }
      pr.n_constraints = (int) cons.size();
      pr.constraints = cons.empty() ? nullptr : cons.data();
      const int n_adof = rb.desc.n_dof;
      const int n = n_adof + (pr.floating_base ? 7 : 0); /* mod.cpp:2128-2131 */
      if (!goals_ptr && !have_starttraj && (int) adofgoal.size() != n_adof) throw module_error("size of adofgoal does not match active dofs!");
      if (n_runs < 1) throw module_error("n_runs must be >=1!");
      std::vector<double> seed_traj; /* [n_points][n] when starttraj was passed */
      if (have_starttraj)
      {
         if (goals_ptr || starts_ptr) throw module_error("Cannot pass both adofgoals/adofstarts and starttraj!");
         sample_starttraj(starttraj, n, pr.n_points, seed_traj);
      }

      std::unique_ptr<Run> r(new Run());
      r->robot_name = robot_name;
      r->n_runs = n_runs;
      r->n_points = pr.n_points;
      r->n_dof = n;
      r->floating = pr.floating_base != 0;
      /* rooted fields: world pose of every grid at this instant (mod.cpp:2347-2369) */
      for (const Field &f : sdfs)
      {
         auto it = env->kinbodies.find(f.kinbody_name);
         if (it == env->kinbodies.end())
            throw module_error("KinBody " + f.kinbody_name + " referenced by active signed distance field does not exist!");
         double pose_world_gsdf[7];
         pose_compose(it->second.pose, f.pose, pose_world_gsdf);
         int id = -1;
         if (ocb_sdf_alias(engine, f.sdf_id, pose_world_gsdf, &id) != OCB_OK)
         {
            for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
            fail_engine("error creating chomp instance!");
         }
         r->sdf_ids.push_back(id);
      }
      std::vector<double> starts((size_t) n_runs * n), goals((size_t) n_runs * n);
      std::vector<unsigned int> seeds(n_runs, seed);
      for (int k = 0; k < n_runs; k++)
         for (int j = 0; j < n; j++)
         {
            if (have_starttraj)
            {
               /* the end rows of the sampled trajectory are the fixed end points (mod.cpp:2578-2580) */
               starts[(size_t) k * n + j] = seed_traj[j];
               goals[(size_t) k * n + j] = seed_traj[(size_t) (pr.n_points - 1) * n + j];
               continue;
            }
            if (pr.floating_base)
            {
               /* rows are [base pose, active dofs]: from the robot's transform to basegoal (mod.cpp:2424-2443) */
               starts[(size_t) k * n + j] = (j < 7) ? rb.desc.base_pose[j] : rb.q[j - 7];
               goals[(size_t) k * n + j] = (j < 7) ? basegoal[j] : adofgoal[j - 7];
               continue;
            }
            starts[(size_t) k * n + j] = starts_ptr ? starts_ptr[(size_t) k * n + j] : rb.q[j]; /* GetActiveDOFValues, mod.cpp:2447 */
            goals[(size_t) k * n + j] = goals_ptr ? goals_ptr[(size_t) k * n + j] : adofgoal[j];
         }
      if (seeds_ptr) seeds.assign(seeds_ptr, seeds_ptr + n_runs);
      const int rc = ocb_batch_create(engine, &rb.desc, &pr, (int) r->sdf_ids.size(), r->sdf_ids.data(), n_runs,
                                      starts.data(), goals.data(), seeds.data(), &r->batch);
      if (rc != OCB_OK)
      {
         for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
         /* the engine's argument messages are the reference's own texts where one exists */
         throw module_error(rc == OCB_ERR_ARG ? ocb_last_error() : std::string("Error initializing chomp instance. ") + ocb_last_error());
      }
      if (have_starttraj)
      {
         std::vector<double> all((size_t) n_runs * seed_traj.size());
         for (int k = 0; k < n_runs; k++) std::copy(seed_traj.begin(), seed_traj.end(), all.begin() + (size_t) k * seed_traj.size());
         if (ocb_batch_set_traj(r->batch, all.data()) != OCB_OK)
         {
            ocb_batch_destroy(r->batch);
            for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
            fail_engine("error creating chomp instance!");
         }
      }
      if (!dat_filename.empty())
      {
         r->fp_dat = fopen(dat_filename.c_str(), "w");
         if (!r->fp_dat)
         {
            ocb_batch_destroy(r->batch);
            for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
            throw module_error("could not open dat_filename file for writing!");
         }
         ocb_batch_enable_trace(r->batch, 1);
      }
      char buf[128];
      snprintf(buf, sizeof(buf), "%p", (void *) r.get()); /* mod.cpp:2670-2674 */
      sout << buf;
      runs.insert(r.release());
      return 0;
   }

   /* mod.cpp:2690-2852 */
   int iterate(const std::vector<std::string> &argv, std::ostream &sout)
   {
      Run *r = nullptr;
      int n_iter = 1;
      double max_time = HUGE_VAL;
      std::string trajs_fileformstr;
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "run" && i + 1 < argv.size())
         {
            if (r) throw module_error("Only one r can be passed!");
            r = run_from_handle(argv[++i].c_str());
         }
         else if (argv[i] == "n_iter" && i + 1 < argv.size()) n_iter = atoi(argv[++i].c_str());
         else if (argv[i] == "max_time" && i + 1 < argv.size()) max_time = atof(argv[++i].c_str());
         else if (argv[i] == "trajs_fileformstr" && i + 1 < argv.size()) trajs_fileformstr = argv[++i];
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!r) throw module_error("you must pass a created run!");
      if (n_iter < 0) throw module_error("n_iter must be >=0!");
      if (!trajs_fileformstr.empty())
      {
         /* the string is used as a printf format with the iteration number (mod.cpp:2779, sprintf): exactly
          * one integer conversion is accepted, anything else (%s, %n, a second conversion) is refused here
          * instead of being undefined behaviour in the planner process */
         int convs = 0;
         bool ok = true;
         for (size_t k = 0; k < trajs_fileformstr.size() && ok; k++)
         {
            if (trajs_fileformstr[k] != '%') continue;
            size_t e = k + 1;
            if (e < trajs_fileformstr.size() && trajs_fileformstr[e] == '%') { k = e; continue; }
            while (e < trajs_fileformstr.size() && strchr("0123456789-+ #", trajs_fileformstr[e])) e++;
            if (e < trajs_fileformstr.size() && strchr("diu", trajs_fileformstr[e])) convs++;
            else ok = false;
            k = e;
         }
         if (!ok || convs != 1) throw module_error("trajs_fileformstr must contain exactly one %d-style conversion!");
      }
      if (!trajs_fileformstr.empty() && r->floating)
         throw module_error("Error: trajs_fileformstr and floating_base combined is not yet implemented!"); /* mod.cpp:2772-2776 */
      std::vector<double> total(r->n_runs);
      std::vector<int> status(r->n_runs);
      struct timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      int done = 0;
      if (!trajs_fileformstr.empty())
      {
         /* the trajectory as it stands before every iteration goes to sprintf(fmt, iter)
          * (mod.cpp:2769-2795): one launch per iteration on this path */
         const int P = r->n_points, n = r->n_dof;
         std::vector<double> traj((size_t) r->n_runs * P * n);
         const Robot &rb = *env->robots[r->robot_name];
         for (; done < n_iter;)
         {
            if (ocb_batch_get_traj(r->batch, traj.data()) != OCB_OK) fail_engine("iterate");
            char fname[4096];
            snprintf(fname, sizeof(fname), trajs_fileformstr.c_str(), done);
            FILE *fp = fopen(fname, "w");
            if (fp)
            {
               fprintf(fp, "<trajectory>\n<configuration>\n<group name=\"joint_values %s", rb.name.c_str());
               for (int j = 0; j < n; j++) fprintf(fp, " %d", j);
               fprintf(fp, "\" offset=\"0\" dof=\"%d\" interpolation=\"linear\"/>\n</configuration>\n<data count=\"%d\">\n", n, P);
               for (int q = 0; q < P * n; q++) fprintf(fp, "%.17g ", traj[q]); /* run 0; full precision, mod.cpp:2790 */
               fprintf(fp, "\n</data>\n</trajectory>\n");
               fclose(fp);
            }
            if (ocb_batch_iterate_from(r->batch, done, 1, total.data(), nullptr, nullptr, status.data()) != OCB_OK) fail_engine("iterate");
            done++;
            if (status[0] == OCB_ERR_JLIMIT) break;
            if (max_time != HUGE_VAL)
            {
               clock_gettime(CLOCK_MONOTONIC, &t1);
               if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > max_time) break;
            }
         }
         if (done == 0 && ocb_batch_iterate(r->batch, 0, total.data(), nullptr, nullptr, status.data()) != OCB_OK) fail_engine("iterate");
      }
      else if (max_time == HUGE_VAL)
      {
         if (ocb_batch_iterate(r->batch, n_iter, total.data(), nullptr, nullptr, status.data()) != OCB_OK) fail_engine("iterate");
         done = n_iter;
      }
      else
      {
         /* the reference checks the clock after every iteration (mod.cpp:2820-2827); the engine
          * is driven one iteration per launch so the same check applies */
         for (; done < n_iter;)
         {
            if (ocb_batch_iterate_from(r->batch, done, 1, total.data(), nullptr, nullptr, status.data()) != OCB_OK) fail_engine("iterate");
            done++;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > max_time) break;
         }
         if (done == 0 && ocb_batch_iterate(r->batch, 0, total.data(), nullptr, nullptr, status.data()) != OCB_OK) fail_engine("iterate");
      }
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (r->fp_dat && max_time == HUGE_VAL && trajs_fileformstr.empty() && n_iter > 0)
      {
         /* rows "iter time total obs smooth" of run 0 (mod.cpp:2811-2818); time is apportioned */
         std::vector<double> trace((size_t) r->n_runs * n_iter * 3);
         if (ocb_batch_get_trace(r->batch, trace.data(), n_iter) == OCB_OK)
         {
            const double el = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
            for (int k = 0; k < n_iter; k++)
               fprintf(r->fp_dat, "%d %f %f %f %f\n", k, el * (k + 1) / n_iter, trace[3 * k], trace[3 * k + 1], trace[3 * k + 2]);
            fflush(r->fp_dat);
         }
      }
      for (int k = 0; k < r->n_runs; k++)
         if (status[k] == OCB_ERR_JLIMIT && r->n_runs == 1)
            throw module_error("Resulting trajectory is outside of joint limits!"); /* mod.cpp:2799-2803 */
      sout << total[0]; /* default stream precision, mod.cpp:2849 */
      for (int k = 1; k < r->n_runs; k++) sout << " " << total[k];
      return 0;
   }

   /* mod.cpp:2854-3011: waypoints in OpenRAVE's trajectory XML layout */
   int gettraj(const std::vector<std::string> &argv, std::ostream &sout)
   {
      Run *r = nullptr;
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "run" && i + 1 < argv.size())
         {
            if (r) throw module_error("Only one run can be passed!");
            r = run_from_handle(argv[++i].c_str());
         }
         else if (argv[i] == "no_collision_check" || argv[i] == "no_collision_exception" || argv[i] == "no_collision_details")
            ;
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!r) throw module_error("you must pass a created run!");
      const int P = r->n_points, n = r->n_dof;
      std::vector<double> traj((size_t) r->n_runs * P * n);
      if (ocb_batch_get_traj(r->batch, traj.data()) != OCB_OK) fail_engine("gettraj");
      const Robot &rb = *env->robots[r->robot_name];
      sout.precision(17);
      /* joint_values + deltatime, and for a floating base the affine_transform group the reference
       * merges in (mod.cpp:2912-2956; OpenRAVE order x y z qw qx qy qz) */
      const int na = r->floating ? n - 7 : n, j0 = r->floating ? 7 : 0;
      for (int k = 0; k < r->n_runs; k++)
      {
         sout << "<trajectory>\n<configuration>\n<group name=\"joint_values " << rb.name;
         for (int j = 0; j < na; j++) sout << " " << j;
         sout << "\" offset=\"0\" dof=\"" << na << "\" interpolation=\"linear\"/>\n"
              << "<group name=\"deltatime\" offset=\"" << na << "\" dof=\"1\" interpolation=\"\"/>\n";
         if (r->floating)
            sout << "<group name=\"affine_transform " << rb.name << " 127\" offset=\"" << na + 1
                 << "\" dof=\"7\" interpolation=\"linear\"/>\n";
         sout << "</configuration>\n<data count=\"" << P << "\">\n";
         for (int p = 0; p < P; p++)
         {
            const double *row = &traj[((size_t) k * P + p) * n];
            for (int j = 0; j < na; j++) sout << row[j0 + j] << " ";
            sout << (p == 0 ? 0.0 : 1.0 / (P - 1)) << " ";
            if (r->floating)
               sout << row[0] << " " << row[1] << " " << row[2] << " " << row[6] << " " << row[3] << " " << row[4] << " "
                    << row[5] << " ";
         }
         sout << "\n</data>\n</trajectory>\n";
      }
      return 0;
   }

   /* mod.cpp:3013-3066 */
   int destroy(const std::vector<std::string> &argv, std::ostream &)
   {
      Run *r = nullptr;
      size_t i = 1;
      for (; i < argv.size(); i++)
      {
         if (argv[i] == "run" && i + 1 < argv.size())
         {
            if (r) throw module_error("Only one run can be passed!");
            r = run_from_handle(argv[++i].c_str());
         }
         else break;
      }
      if (i < argv.size()) bad_args(argv, i);
      if (!r) throw module_error("you must pass a created run!");
      ocb_batch_destroy(r->batch);
      for (int s : r->sdf_ids) ocb_sdf_remove(engine, s);
      if (r->fp_dat) fclose(r->fp_dat);
      runs.erase(r);
      delete r;
      return 0;
   }

   /* orcwrap_call (orcwrap.cpp:37-69): argv[0] is the literal "openrave_command" */
   int dispatch(const std::string &cmdline, std::ostream &sout)
   {
      std::vector<std::string> words = shell_split(cmdline);
      if (words.empty()) throw module_error("empty command");
      const std::string name = words[0];
      words[0] = "openrave_command";
      if (name == "viewspheres") return viewspheres(words, sout);
      if (name == "computedistancefield") return computedistancefield(words, sout);
      if (name == "addfield_fromobsarray") return addfield_fromobsarray(words, sout);
      if (name == "viewfields") return viewfields(words, sout);
      if (name == "removefield") return removefield(words, sout);
      if (name == "create") return create(words, sout, false);
      if (name == "createbatch") return create(words, sout, true);
      if (name == "iterate") return iterate(words, sout);
      if (name == "gettraj") return gettraj(words, sout);
      if (name == "destroy") return destroy(words, sout);
      throw module_error("unknown command '" + name + "'");
   }
};

/* ------------------------------------------------------------------- C ABI */
static thread_local std::string g_module_err;

extern "C" int ocb_env_create(ocb_env **out)
{
   if (!out) return OCB_ERR_ARG;
   *out = new ocb_env();
   return OCB_OK;
}

extern "C" int ocb_env_destroy(ocb_env *env)
{
   delete env;
   return OCB_OK;
}

extern "C" int ocb_env_add_kinbody(ocb_env *env, const char *name, const double pose[7], const ocb_prim *prims, int n_prims)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!env || !name || !pose || env->kinbodies.count(name)) return OCB_ERR_ARG;
   Kinbody kb;
   kb.name = name;
   memcpy(kb.pose, pose, sizeof(kb.pose));
   if (prims) kb.prims.assign(prims, prims + n_prims);
   env->kinbodies[name] = kb;
   return OCB_OK;
}

extern "C" int ocb_env_set_kinbody_pose(ocb_env *env, const char *name, const double pose[7])
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!env || !name || !env->kinbodies.count(name)) return OCB_ERR_ARG;
   memcpy(env->kinbodies[name].pose, pose, 7 * sizeof(double));
   return OCB_OK;
}

extern "C" int ocb_env_enable_kinbody(ocb_env *env, const char *name, int enabled)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!env || !name || !env->kinbodies.count(name)) return OCB_ERR_ARG;
   env->kinbodies[name].enabled = enabled != 0;
   return OCB_OK;
}

extern "C" int ocb_env_add_robot(ocb_env *env, const char *name, const ocb_robot *robot, const double *values)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!env || !name || !robot || !values || env->robots.count(name) || env->kinbodies.count(name)) return OCB_ERR_ARG;
   std::unique_ptr<Robot> r(new Robot());
   r->name = name;
   r->assign(robot, values);
   env->robots[name] = std::move(r);
   Kinbody kb;
   kb.name = name;
   memcpy(kb.pose, robot->base_pose, sizeof(kb.pose));
   env->kinbodies[name] = kb;
   return OCB_OK;
}

extern "C" int ocb_tsr_parse(const char *text, double T0w[7], double Twe[7], double Bw[12])
{
   Tsr t;
   if (!text || !T0w || !Twe || !Bw || !parse_tsr(text, t)) return OCB_ERR_ARG;
   memcpy(T0w, t.T0w, sizeof(t.T0w));
   memcpy(Twe, t.Twe, sizeof(t.Twe));
   memcpy(Bw, t.Bw, sizeof(t.Bw));
   return OCB_OK;
}

extern "C" int ocb_env_set_link_names(ocb_env *env, const char *robot, const char *const *names, int n_names)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!robot || !names || !env->robots.count(robot)) return OCB_ERR_ARG;
   Robot &r = *env->robots[robot];
   if (n_names != r.desc.n_links) return OCB_ERR_ARG;
   r.link_names.assign(names, names + n_names);
   return OCB_OK;
}

extern "C" int ocb_env_add_manipulator(ocb_env *env, const char *robot, const char *name, int ee_link,
                                       const double local_tool[7])
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!robot || !name || !env->robots.count(robot)) return OCB_ERR_ARG;
   Robot &r = *env->robots[robot];
   if (ee_link < 0 || ee_link >= r.desc.n_links) return OCB_ERR_ARG;
   for (const Robot::Manip &m : r.manips)
      if (m.name == name) return OCB_ERR_ARG;
   Robot::Manip m;
   m.name = name;
   m.link = ee_link;
   const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
   memcpy(m.tool, local_tool ? local_tool : ident, sizeof(m.tool));
   r.manips.push_back(m);
   return OCB_OK;
}

extern "C" int ocb_env_set_active_manipulator(ocb_env *env, const char *robot, const char *name)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!robot || !name || !env->robots.count(robot)) return OCB_ERR_ARG;
   Robot &r = *env->robots[robot];
   for (size_t k = 0; k < r.manips.size(); k++)
      if (r.manips[k].name == name) { r.active_manip = (int) k; return OCB_OK; }
   return OCB_ERR_ARG;
}

extern "C" int ocb_env_set_active_dof_values(ocb_env *env, const char *name, const double *values)
{
   if (!env) return OCB_ERR_ARG;
   std::lock_guard<std::recursive_mutex> guard(env->mutex);
   if (!env || !name || !values || !env->robots.count(name)) return OCB_ERR_ARG;
   Robot &r = *env->robots[name];
   r.q.assign(values, values + r.desc.n_dof);
   return OCB_OK;
}

extern "C" int ocb_module_create(ocb_env *env, int device, ocb_module **out)
{
   if (!env || !out) return OCB_ERR_ARG;
   *out = nullptr;
   ocb_engine *e = nullptr;
   const int rc = ocb_engine_create(device, &e);
   if (rc != OCB_OK) return rc; /* no GPU: the module refuses to exist, there is no CPU path */
   ocb_module *m = new ocb_module();
   m->env = env;
   m->engine = e;
   *out = m;
   return OCB_OK;
}

extern "C" int ocb_module_destroy(ocb_module *m)
{
   delete m;
   return OCB_OK;
}

extern "C" const char *ocb_module_last_error(void) { return g_module_err.c_str(); }

extern "C" int ocb_module_send_command(ocb_module *m, const char *cmd, char *out, size_t out_cap, size_t *out_len)
{
   if (!m || !cmd) return OCB_ERR_ARG;
   std::ostringstream sout;
   int ret;
   try
   {
      std::lock_guard<std::recursive_mutex> guard(m->env->mutex);
      ret = m->dispatch(cmd, sout);
   }
   catch (const std::exception &ex)
   {
      g_module_err = ex.what();
      if (out_len) *out_len = 0;
      return 1;
   }
   const std::string s = sout.str();
   if (out_len) *out_len = s.size();
   if (out && out_cap)
   {
      const size_t k = s.size() < out_cap - 1 ? s.size() : out_cap - 1;
      memcpy(out, s.data(), k);
      out[k] = 0;
   }
   return ret;
}

extern "C" int ocb_module_run_batch(ocb_module *m, const char *handle, ocb_batch **batch)
{
   if (!m || !handle || !batch) return OCB_ERR_ARG;
   try
   {
      *batch = m->run_from_handle(handle)->batch;
   }
   catch (const std::exception &ex)
   {
      g_module_err = ex.what();
      return 1;
   }
   return OCB_OK;
}
