/* adapter_probe.cpp -- exercises include/orcdchomp_b200_openrave.h (the OpenRAVE-side chain extractor)
 * against a robot that implements the OpenRAVE accessors it uses (include/openrave_min/).  The robot is
 * a seeded random tree: revolute / prismatic / static joints, Left and Right hierarchy transforms with
 * rotation AND translation (so joint axes do not pass through the child links' origins), some joints
 * passive, some dofs inactive.  The probe prints, as text for tests/test_adapter.py:
 *   - the extracted ocb_robot arrays;
 *   - for a few random active-dof vectors, every link's world pose computed by OpenRAVE's own rule
 *       T_child = T_parent * Left * motion(axis, value) * Right
 *     so that the test can run forward kinematics through the extracted description (the CPU oracle's
 *     FK, i.e. what the kernels evaluate) and compare. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <openrave/openrave.h>
#include "orcdchomp_b200_openrave.h"

using namespace OpenRAVE;

static double urand() { return rand() / (double) RAND_MAX; }

/* quaternion helpers on libcd poses [x y z qx qy qz qw] */
struct Pose { double v[7]; };
static Pose pose_mul(const Pose &a, const Pose &b)
{
   const double ax = a.v[3], ay = a.v[4], az = a.v[5], aw = a.v[6];
   const double bx = b.v[3], by = b.v[4], bz = b.v[5], bw = b.v[6];
   Pose c;
   /* rotate b's translation by a */
   const double x = b.v[0], y = b.v[1], z = b.v[2];
   const double ix = aw * x + ay * z - az * y, iy = aw * y + az * x - ax * z, iz = aw * z + ax * y - ay * x,
                iw = -ax * x - ay * y - az * z;
   c.v[0] = ix * aw + iw * -ax + iy * -az - iz * -ay + a.v[0];
   c.v[1] = iy * aw + iw * -ay + iz * -ax - ix * -az + a.v[1];
   c.v[2] = iz * aw + iw * -az + ix * -ay - iy * -ax + a.v[2];
   c.v[3] = aw * bx + ax * bw + ay * bz - az * by;
   c.v[4] = aw * by - ax * bz + ay * bw + az * bx;
   c.v[5] = aw * bz + ax * by - ay * bx + az * bw;
   c.v[6] = aw * bw - ax * bx - ay * by - az * bz;
   return c;
}
static Pose random_pose(double span)
{
   Pose p;
   double q[4], n = 0;
   for (int k = 0; k < 3; k++) p.v[k] = span * (2 * urand() - 1);
   for (int k = 0; k < 4; k++) { q[k] = 2 * urand() - 1; n += q[k] * q[k]; }
   n = sqrt(n);
   for (int k = 0; k < 4; k++) p.v[3 + k] = q[k] / n;
   return p;
}
static Transform to_transform(const Pose &p)
{
   Transform t;
   t.trans = Vector(p.v[0], p.v[1], p.v[2]);
   t.rot = Vector(p.v[6], p.v[3], p.v[4], p.v[5]); /* (w, x, y, z) in the fields (x, y, z, w) */
   return t;
}

struct FakeLink : KinBody::Link
{
   int index;
   std::string name;
   Pose world;
   int GetIndex() const override { return index; }
   const std::string &GetName() const override { return name; }
   Transform GetTransform() const override { return to_transform(world); }
};

struct FakeJoint : KinBody::Joint
{
   KinBody::LinkPtr parent, child;
   Pose left, right;
   double axis[3];
   int type; /* 0 static, 1 revolute, 2 prismatic */
   int dofindex;
   double value;
   KinBody::LinkPtr GetHierarchyParentLink() const override { return parent; }
   KinBody::LinkPtr GetHierarchyChildLink() const override { return child; }
   Transform GetInternalHierarchyLeftTransform() const override { return to_transform(left); }
   Transform GetInternalHierarchyRightTransform() const override { return to_transform(right); }
   Vector GetInternalHierarchyAxis(int) const override { return Vector(axis[0], axis[1], axis[2]); }
   bool IsStatic() const override { return type == 0; }
   bool IsRevolute(int) const override { return type == 1; }
   bool IsPrismatic(int) const override { return type == 2; }
   bool IsMimic(int) const override { return false; }
   int GetDOF() const override { return type == 0 ? 0 : 1; }
   int GetDOFIndex() const override { return dofindex; }
   dReal GetValue(int) const override { return value; }
};

struct FakeManip : RobotBase::Manipulator
{
   std::string name;
   KinBody::LinkPtr ee;
   Pose tool;
   const std::string &GetName() const override { return name; }
   KinBody::LinkPtr GetEndEffector() const override { return ee; }
   Transform GetLocalToolTransform() const override { return to_transform(tool); }
};

struct FakeRobot : RobotBase
{
   std::vector<RobotBase::ManipulatorPtr> manips;
   const std::vector<RobotBase::ManipulatorPtr> &GetManipulators() const override { return manips; }
   RobotBase::ManipulatorPtr GetActiveManipulator() const override { return manips.empty() ? nullptr : manips[0]; }
   std::string name = "probe";
   std::vector<KinBody::LinkPtr> links;
   std::vector<KinBody::JointPtr> joints, passive;
   std::vector<int> adof;
   std::vector<double> lo, hi;
   Pose base;
   const std::string &GetName() const override { return name; }
   const std::vector<KinBody::LinkPtr> &GetLinks() const override { return links; }
   const std::vector<KinBody::JointPtr> &GetJoints() const override { return joints; }
   const std::vector<KinBody::JointPtr> &GetPassiveJoints() const override { return passive; }
   Transform GetTransform() const override { return to_transform(base); }
   void GetDOFLimits(std::vector<dReal> &l, std::vector<dReal> &u) const override { l = lo; u = hi; }
   const std::vector<int> &GetActiveDOFIndices() const override { return adof; }
};

int main(int argc, char **argv)
{
   srand(argc > 1 ? atoi(argv[1]) : 7);
   FakeRobot rb;
   const int nl = 9;
   rb.base = random_pose(0.5);
   std::vector<std::shared_ptr<FakeLink>> L;
   for (int i = 0; i < nl; i++)
   {
      std::shared_ptr<FakeLink> l(new FakeLink());
      l->index = i;
      l->name = "link" + std::to_string(i);
      L.push_back(l);
      rb.links.push_back(l);
   }
   /* link indices are deliberately not in tree order: link 0 is the root, parents drawn among
    * already connected links in a shuffled order */
   const int conn_order[nl] = {0, 4, 2, 7, 1, 8, 3, 6, 5};
   std::vector<std::shared_ptr<FakeJoint>> J;
   int ndof = 0;
   for (int k = 1; k < nl; k++)
   {
      std::shared_ptr<FakeJoint> j(new FakeJoint());
      j->child = L[conn_order[k]];
      j->parent = L[conn_order[rand() % k]];
      j->left = random_pose(0.3);
      j->right = random_pose(0.2);
      double n = 0;
      for (int c = 0; c < 3; c++) { j->axis[c] = 2 * urand() - 1; n += j->axis[c] * j->axis[c]; }
      for (int c = 0; c < 3; c++) j->axis[c] /= sqrt(n);
      j->type = (k == 3) ? 0 : ((k == 5) ? 2 : 1);
      const bool is_passive = (k == 6);
      j->dofindex = (j->type == 0 || is_passive) ? -1 : ndof++;
      j->value = (j->type == 0) ? 0.0 : (2 * urand() - 1);
      if (j->type == 0 || is_passive) rb.passive.push_back(j); else rb.joints.push_back(j);
      J.push_back(j);
   }
   for (int d = 0; d < ndof; d++) { rb.lo.push_back(-2.0 - d); rb.hi.push_back(1.5 + d); }
   /* active dofs: all but two, in a scrambled order */
   for (int d = ndof - 1; d >= 0; d--)
      if (d != 1 && d != 4) rb.adof.push_back(d);
   std::vector<ocb_or::SphereSpec> spheres;
   for (int s = 0; s < 6; s++)
   {
      ocb_or::SphereSpec sp;
      sp.linkname = L[(3 * s + 1) % nl]->name;
      for (int c = 0; c < 3; c++) sp.pos[c] = 0.1 * (2 * urand() - 1);
      sp.radius = 0.03 + 0.01 * s;
      spheres.push_back(sp);
   }
   ocb_or::RobotArrays A;
   try { ocb_or::extract_robot(rb, spheres, A); }
   catch (const std::exception &e) { printf("error %s\n", e.what()); return 1; }
   const ocb_robot &r = A.robot;
   printf("n_links %d n_dof %d n_spheres %d n_or_links %d\n", r.n_links, r.n_dof, r.n_spheres, nl);
   for (int i = 0; i < r.n_links; i++)
   {
      printf("link %d %d %d %.17g %.17g", r.parent[i], r.joint_type[i], r.dof_index[i], r.dof_coeff[2 * i], r.dof_coeff[2 * i + 1]);
      for (int c = 0; c < 7; c++) printf(" %.17g", r.pose_parent[7 * i + c]);
      for (int c = 0; c < 3; c++) printf(" %.17g", r.axis[3 * i + c]);
      printf("\n");
   }
   printf("base"); for (int c = 0; c < 7; c++) printf(" %.17g", r.base_pose[c]); printf("\n");
   printf("limits"); for (int d = 0; d < r.n_dof; d++) printf(" %.17g %.17g", r.limit_lower[d], r.limit_upper[d]); printf("\n");
   for (int s = 0; s < r.n_spheres; s++)
      printf("sphere %d %.17g %.17g %.17g %.17g\n", r.sphere_link[s], r.sphere_pos[3 * s], r.sphere_pos[3 * s + 1], r.sphere_pos[3 * s + 2], r.sphere_radius[s]);
   printf("map"); for (int i = 0; i < nl; i++) printf(" %d", A.link_map[i]); printf("\n");
   /* a manipulator on OpenRAVE link 5 with a tool offset: the constraint frame the adapter hands to the engine */
   {
      std::shared_ptr<FakeManip> mp(new FakeManip());
      mp->name = "arm";
      mp->ee = L[5];
      mp->tool = random_pose(0.1);
      rb.manips.push_back(mp);
      const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
      double Bw[6][2] = {{0, 0}, {-1, 1}, {0, 0}, {0, 0}, {-1, 1}, {-1, 1}};
      const ocb_constraint c1 = ocb_or::tsr_constraint(A, OCB_CON_ALL, rb.GetActiveManipulator().get(), nullptr, ident, ident, Bw);
      const ocb_constraint c2 = ocb_or::tsr_constraint(A, OCB_CON_START, nullptr, L[3].get(), ident, ident, Bw);
      printf("con %d %d", c1.where, c1.link); for (int k = 0; k < 7; k++) printf(" %.17g", c1.pose_link_ee[k]); printf("\n");
      printf("con %d %d", c2.where, c2.link); for (int k = 0; k < 7; k++) printf(" %.17g", c2.pose_link_ee[k]); printf("\n");
      printf("tool"); for (int k = 0; k < 7; k++) printf(" %.17g", mp->tool.v[k]); printf("\n");
   }
   /* reference link poses by OpenRAVE's rule, for three active-dof vectors */
   for (int trial = 0; trial < 3; trial++)
   {
      std::vector<double> q(rb.adof.size());
      for (double &x : q) x = 2 * urand() - 1;
      printf("q"); for (double x : q) printf(" %.17g", x); printf("\n");
      std::vector<int> done(nl, 0);
      L[0]->world = rb.base;
      done[0] = 1;
      for (int pass = 0; pass < nl; pass++)
         for (auto &j : J)
         {
            const int p = j->parent->GetIndex(), c = j->child->GetIndex();
            if (!done[p] || done[c]) continue;
            double val = j->value;
            for (size_t a = 0; a < rb.adof.size(); a++)
               if (j->dofindex >= 0 && rb.adof[a] == j->dofindex) val = q[a];
            Pose m = {{0, 0, 0, 0, 0, 0, 1}};
            if (j->type == 1) { const double s = sin(0.5 * val); m.v[3] = j->axis[0] * s; m.v[4] = j->axis[1] * s; m.v[5] = j->axis[2] * s; m.v[6] = cos(0.5 * val); }
            if (j->type == 2) for (int k = 0; k < 3; k++) m.v[k] = j->axis[k] * val;
            L[c]->world = pose_mul(pose_mul(pose_mul(L[p]->world, j->left), m), j->right);
            done[c] = 1;
         }
      for (int i = 0; i < nl; i++)
      {
         printf("world %d", i);
         for (int c = 0; c < 7; c++) printf(" %.17g", L[i]->world.v[c]);
         printf("\n");
      }
   }
   return 0;
}
