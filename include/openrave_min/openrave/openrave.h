/* openrave_min/openrave/openrave.h -- the handful of declarations of OpenRAVE's public API that
 * orcdchomp_b200_openrave.h uses, for building and testing that adapter on machines without OpenRAVE
 * (this image has no OpenRAVE, boost or ROS).  NOT OpenRAVE: no implementation, no behaviour -- only
 * the names, signatures and documented meaning of the accessors the reference module itself calls
 * (src/orcdchomp_mod.cpp:2104-2300, 2639-2660) plus the joint-hierarchy accessors of KinBody::Joint.
 * With a real installation put its include directory first: the adapter compiles against either.
 *
 * Kinematic convention these accessors describe (OpenRAVE KinBody::Joint):
 *    T_child = T_parent * GetInternalHierarchyLeftTransform() * motion(GetInternalHierarchyAxis(0), value)
 *                        * GetInternalHierarchyRightTransform()
 * Transform.rot is the quaternion (w, x, y, z) stored in the fields (x, y, z, w)
 * (hence rot.y, rot.z, rot.w, rot.x = qx, qy, qz, qw at mod.cpp:483-489). */
#ifndef OPENRAVE_MIN_OPENRAVE_H
#define OPENRAVE_MIN_OPENRAVE_H

#include <memory>
#include <string>
#include <vector>

#define OPENRAVE_MIN_STUB 1

namespace OpenRAVE
{
typedef double dReal;

struct Vector
{
   dReal x, y, z, w;
   Vector() : x(0), y(0), z(0), w(0) {}
   Vector(dReal x_, dReal y_, dReal z_, dReal w_ = 0) : x(x_), y(y_), z(z_), w(w_) {}
};

struct Transform
{
   Vector rot;   /* quaternion: rot.x = w, rot.y = qx, rot.z = qy, rot.w = qz */
   Vector trans;
   Transform() : rot(1, 0, 0, 0) {}
};

class KinBody
{
public:
   class Link
   {
   public:
      virtual ~Link() {}
      virtual int GetIndex() const = 0;
      virtual const std::string &GetName() const = 0;
      virtual Transform GetTransform() const = 0;
   };
   typedef std::shared_ptr<Link> LinkPtr;

   class Joint
   {
   public:
      virtual ~Joint() {}
      virtual LinkPtr GetHierarchyParentLink() const = 0;
      virtual LinkPtr GetHierarchyChildLink() const = 0;
      virtual Transform GetInternalHierarchyLeftTransform() const = 0;
      virtual Transform GetInternalHierarchyRightTransform() const = 0;
      virtual Vector GetInternalHierarchyAxis(int iaxis = 0) const = 0;
      virtual bool IsStatic() const = 0;
      virtual bool IsRevolute(int iaxis) const = 0;
      virtual bool IsPrismatic(int iaxis) const = 0;
      virtual bool IsMimic(int iaxis = -1) const = 0;
      virtual int GetDOF() const = 0;
      virtual int GetDOFIndex() const = 0; /* -1 for passive joints */
      virtual dReal GetValue(int iaxis) const = 0;
   };
   typedef std::shared_ptr<Joint> JointPtr;

   virtual ~KinBody() {}
   virtual const std::string &GetName() const = 0;
   virtual const std::vector<LinkPtr> &GetLinks() const = 0;
   virtual const std::vector<JointPtr> &GetJoints() const = 0;
   virtual const std::vector<JointPtr> &GetPassiveJoints() const = 0;
   virtual Transform GetTransform() const = 0;
   virtual void GetDOFLimits(std::vector<dReal> &lower, std::vector<dReal> &upper) const = 0;
};

class RobotBase : public KinBody
{
public:
   class Manipulator
   {
   public:
      virtual ~Manipulator() {}
      virtual const std::string &GetName() const = 0;
      virtual LinkPtr GetEndEffector() const = 0;
      virtual Transform GetLocalToolTransform() const = 0;
   };
   typedef std::shared_ptr<Manipulator> ManipulatorPtr;

   virtual const std::vector<int> &GetActiveDOFIndices() const = 0;
   virtual const std::vector<ManipulatorPtr> &GetManipulators() const = 0;
   virtual ManipulatorPtr GetActiveManipulator() const = 0;
};
} /* namespace OpenRAVE */

#endif
