/* orcdchomp_b200_openrave.h -- the OpenRAVE side of the drop-in: what
 * mod::create pulls out of OpenRAVE once per run, written against OpenRAVE's public API.
 *
 *   ocb_or::extract_robot   the kinematic tree + dof bookkeeping + limits + sphere table as the
 *                           struct ocb_robot of orcdchomp_b200.h.  Replaces, for the device path,
 *                           the per-waypoint SetActiveDOFValues / Link::GetTransform /
 *                           CalculateJacobian calls of sphere_cost_pre (src/orcdchomp_mod.cpp:
 *                           1022-1049) and restates the bookkeeping of mod::create
 *                           (src/orcdchomp_mod.cpp:2104-2134 dofs, 2180-2210 spheres, 2639-2660 limits).
 *   ocb_or::tsr_constraint  one con_tsr / start_tsr / everyn_tsr argument (manipulator end effector with its
 *                           local tool transform, or a bare link) as the engine's ocb_constraint.
 *
 * Header only; include <openrave/openrave.h> (or include/openrave_min/, a declaration-only stand-in
 * used to build and test this file where OpenRAVE is absent) before it.
 *
 * Mapping.  OpenRAVE places a child link at
 *    T_child = T_parent * Left * motion(axis, value) * Right
 * (KinBody::Joint::GetInternalHierarchy{Left,Right}Transform, GetInternalHierarchyAxis), with the axis
 * through the origin of the JOINT frame (T_parent * Left), which in general is not the child link's
 * origin.  struct ocb_robot wants every moving joint's axis through the origin of the frame it
 * moves, so each OpenRAVE link with a parent joint becomes TWO ocb links: the joint frame
 * (pose_parent = Left, the joint's type / axis / dof) and the link frame (a FIXED child at Right).
 * The engine folds fixed frames away when it compiles the robot, so this costs nothing at run time.
 * ocb link of OpenRAVE link i: link_map[i]; spheres are attached to those.
 *
 * Not supported (extract_robot throws std::runtime_error): joints with more than one dof, mimic joints
 * that follow an ACTIVE dof (their value would have to move with it; a mimic joint of an inactive dof
 * is frozen at its current value like any other inactive joint), links with no parent joint other
 * than the root. */
#ifndef ORCDCHOMP_B200_OPENRAVE_H
#define ORCDCHOMP_B200_OPENRAVE_H

#include <stdexcept>
#include <string>
#include <vector>

#include "orcdchomp_b200.h"

namespace ocb_or
{

/* one entry of the <orcdchomp><spheres> block (src/orcdchomp_kdata.h: struct sphere) */
struct SphereSpec
{
   std::string linkname;
   double pos[3];
   double radius;
};

/* owns the arrays behind an ocb_robot */
struct RobotArrays
{
   std::vector<int> parent, joint_type, dof_index, sphere_link, link_map;
   std::vector<double> pose_parent, axis, dof_coeff, limit_lower, limit_upper, sphere_pos, sphere_radius;
   ocb_robot robot;
};

/* libcd pose [x y z qx qy qz qw] of an OpenRAVE transform (mod.cpp:483-489) */
inline void pose_from_transform(const OpenRAVE::Transform &t, double *pose)
{
   pose[0] = t.trans.x; pose[1] = t.trans.y; pose[2] = t.trans.z;
   pose[3] = t.rot.y; pose[4] = t.rot.z; pose[5] = t.rot.w; pose[6] = t.rot.x;
}

inline void extract_robot(const OpenRAVE::RobotBase &robot, const std::vector<SphereSpec> &spheres, RobotArrays &A)
{
   typedef OpenRAVE::KinBody::JointPtr JointPtr;
   const std::vector<OpenRAVE::KinBody::LinkPtr> &links = robot.GetLinks();
   const std::vector<int> &adof = robot.GetActiveDOFIndices(); /* r->adofindices, mod.cpp:2120-2134 */
   const int nl = (int) links.size();
   if (nl < 1) throw std::runtime_error("robot has no links");

   /* the joint whose hierarchy child is each link (active and passive joints alike) */
   std::vector<JointPtr> up(nl);
   std::vector<JointPtr> all(robot.GetJoints());
   all.insert(all.end(), robot.GetPassiveJoints().begin(), robot.GetPassiveJoints().end());
   for (const JointPtr &j : all)
   {
      const int child = j->GetHierarchyChildLink()->GetIndex();
      if (child < 0 || child >= nl) throw std::runtime_error("joint with a child link outside the robot");
      if (up[child]) throw std::runtime_error("link " + links[child]->GetName() + " has two parent joints (closed chain)");
      up[child] = j;
   }
   /* emit links parents first: ocb_robot wants parent[i] < i */
   std::vector<int> order, emitted(nl, 0);
   int root = -1;
   for (int i = 0; i < nl; i++)
      if (!up[i])
      {
         if (root >= 0) throw std::runtime_error("link " + links[i]->GetName() + " has no parent joint (second root)");
         root = i;
      }
   if (root < 0) throw std::runtime_error("no root link");
   order.push_back(root);
   emitted[root] = 1;
   for (size_t k = 0; k < order.size(); k++)
      for (int i = 0; i < nl; i++)
         if (!emitted[i] && up[i] && up[i]->GetHierarchyParentLink()->GetIndex() == order[k])
         {
            order.push_back(i);
            emitted[i] = 1;
         }
   if ((int) order.size() != nl) throw std::runtime_error("kinematic tree is not connected");

   A = RobotArrays();
   A.link_map.assign(nl, -1);
   auto add_link = [&](int parent, const double *pose, int type, const double *ax, int dof, double c0, double c1)
   {
      A.parent.push_back(parent);
      A.pose_parent.insert(A.pose_parent.end(), pose, pose + 7);
      A.joint_type.push_back(type);
      A.axis.insert(A.axis.end(), ax, ax + 3);
      A.dof_index.push_back(dof);
      A.dof_coeff.push_back(c0);
      A.dof_coeff.push_back(c1);
      return (int) A.parent.size() - 1;
   };
   const double ident[7] = {0, 0, 0, 0, 0, 0, 1}, zaxis[3] = {0, 0, 1};
   A.link_map[root] = add_link(-1, ident, OCB_JOINT_FIXED, zaxis, -1, 0.0, 0.0);
   for (size_t k = 1; k < order.size(); k++)
   {
      const int i = order[k];
      const JointPtr &j = up[i];
      const int parent = A.link_map[j->GetHierarchyParentLink()->GetIndex()];
      double left[7], right[7];
      pose_from_transform(j->GetInternalHierarchyLeftTransform(), left);
      pose_from_transform(j->GetInternalHierarchyRightTransform(), right);
      int type = OCB_JOINT_FIXED, dof = -1;
      double ax[3] = {0, 0, 1}, c0 = 0.0, c1 = 0.0;
      if (!j->IsStatic() && j->GetDOF() > 0)
      {
         if (j->GetDOF() != 1) throw std::runtime_error("joints with more than one dof are not supported");
         type = j->IsRevolute(0) ? OCB_JOINT_REVOLUTE : OCB_JOINT_PRISMATIC;
         const OpenRAVE::Vector a = j->GetInternalHierarchyAxis(0);
         ax[0] = a.x; ax[1] = a.y; ax[2] = a.z;
         const int di = j->GetDOFIndex();
         for (size_t q = 0; q < adof.size(); q++)
            if (di >= 0 && adof[q] == di) dof = (int) q;
         if (dof >= 0)
         {
            if (j->IsMimic(0)) throw std::runtime_error("an active dof drives a mimic joint: not supported");
            c0 = 1.0; /* value = q[dof] */
         }
         else
            c1 = j->GetValue(0); /* inactive dofs stay where the robot has them (SetActiveDOFValues leaves them) */
      }
      const int jf = add_link(parent, left, type, ax, dof, c0, c1);
      A.link_map[i] = add_link(jf, right, OCB_JOINT_FIXED, zaxis, -1, 0.0, 0.0);
   }

   /* limits of the active dofs (mod.cpp:2639-2660) */
   std::vector<OpenRAVE::dReal> lo, hi;
   robot.GetDOFLimits(lo, hi);
   for (size_t q = 0; q < adof.size(); q++)
   {
      if (adof[q] < 0 || adof[q] >= (int) lo.size()) throw std::runtime_error("active dof index out of range");
      A.limit_lower.push_back(lo[adof[q]]);
      A.limit_upper.push_back(hi[adof[q]]);
   }
   /* spheres, in <orcdchomp><spheres> order (the reference's two list reversals cancel) */
   for (const SphereSpec &s : spheres)
   {
      int li = -1;
      for (int i = 0; i < nl; i++)
         if (links[i]->GetName() == s.linkname) li = i;
      if (li < 0) throw std::runtime_error("link " + s.linkname + " in <orcdchomp> does not exist."); /* mod.cpp:2186 */
      A.sphere_link.push_back(A.link_map[li]);
      A.sphere_pos.insert(A.sphere_pos.end(), s.pos, s.pos + 3);
      A.sphere_radius.push_back(s.radius);
   }
   ocb_robot &r = A.robot;
   r.n_links = (int) A.parent.size();
   r.parent = A.parent.data();
   r.pose_parent = A.pose_parent.data();
   r.joint_type = A.joint_type.data();
   r.axis = A.axis.data();
   r.dof_index = A.dof_index.data();
   r.dof_coeff = A.dof_coeff.data();
   /* T_world_root: the robot's transform is its root link's (KinBody::GetTransform) */
   pose_from_transform(robot.GetTransform(), r.base_pose);
   r.n_dof = (int) adof.size();
   r.limit_lower = A.limit_lower.data();
   r.limit_upper = A.limit_upper.data();
   r.n_spheres = (int) A.sphere_radius.size();
   r.sphere_link = A.sphere_link.data();
   r.sphere_pos = A.sphere_pos.data();
   r.sphere_radius = A.sphere_radius.data();
}

/* One TSR argument of `create` as the engine's ocb_constraint.  The constrained frame is what the reference's
 * callbacks read on every evaluation: contsr->manip->GetEndEffectorTransform() = end-effector link transform *
 * GetLocalToolTransform() (src/orcdchomp_mod.cpp:1384-1385; the active manipulator for start_tsr / everyn_tsr,
 * 1545, 1701), or contsr->link->GetTransform() (1387).  Pass the manipulator, or a link with manip == nullptr.
 * where: OCB_CON_START / OCB_CON_END / OCB_CON_ALL for con_tsr 'start' / 'end' / 'all' (1947-1956),
 * OCB_CON_ALL for everyn_tsr, OCB_CON_START_TSR for start_tsr.  T0w, Twe, Bw: struct tsr as parsed by
 * tsr_create_parse (3068-3110). */
inline ocb_constraint tsr_constraint(const RobotArrays &A, int where, const OpenRAVE::RobotBase::Manipulator *manip,
                                     const OpenRAVE::KinBody::Link *link, const double T0w[7], const double Twe[7],
                                     const double Bw[6][2])
{
   ocb_constraint c;
   const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
   c.where = where;
   int or_link = -1;
   for (int k = 0; k < 7; k++) c.pose_link_ee[k] = ident[k];
   if (manip)
   {
      or_link = manip->GetEndEffector()->GetIndex();
      pose_from_transform(manip->GetLocalToolTransform(), c.pose_link_ee);
   }
   else if (link)
      or_link = link->GetIndex();
   if (or_link < 0 || or_link >= (int) A.link_map.size()) throw std::runtime_error("con_tsr link not found!"); /* mod.cpp:1974 */
   c.link = A.link_map[or_link];
   for (int k = 0; k < 7; k++) { c.T0w[k] = T0w[k]; c.Twe[k] = Twe[k]; }
   for (int i = 0; i < 6; i++) { c.Bw[i][0] = Bw[i][0]; c.Bw[i][1] = Bw[i][1]; }
   return c;
}

} /* namespace ocb_or */

#endif
