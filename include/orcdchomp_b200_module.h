/* orcdchomp_b200_module.h -- the module-command boundary of or_cdchomp, as a C ABI.
 *
 * The reference registers nine string commands on an OpenRAVE module called
 * "orcdchomp" (src/orcdchomp_mod.h:55-71; entry point src/orcdchomp.cpp:50-74;
 * stream -> argv adapter src/orcwrap.cpp:37-82, tokeniser
 * src/libcd/util_shparse.c:37-128).  This header exposes the same commands, with
 * the same grammar, return strings and error texts, on top of the B200 engine
 * (include/orcdchomp_b200.h):
 *
 *   viewspheres            mod.cpp:175-289   (viewer only: lists the spheres as text)
 *   computedistancefield   mod.cpp:297-589
 *   addfield_fromobsarray  mod.cpp:592-722
 *   viewfields             mod.cpp:724-797   (viewer only: lists the fields as text)
 *   removefield            mod.cpp:799-847
 *   create                 mod.cpp:1800-2688
 *   iterate                mod.cpp:2690-2852
 *   gettraj                mod.cpp:2854-3011
 *   destroy                mod.cpp:3013-3037
 *
 * plus one extension, `createbatch`, which makes R runs behind one handle (the
 * handle then works with iterate / gettraj / destroy exactly like a single run).
 *
 * OpenRAVE itself is not part of this build.  What the commands need from the
 * environment -- named kinbodies with a pose and collision geometry, named robots
 * with a kinematic tree, sphere table and current active-DOF values -- is held
 * by a small stand-in (`ocb_env`).  In a live OpenRAVE plugin the same data is
 * read from EnvironmentBase / KinBody / RobotBase.
 *
 * Differences from the reference that a caller can observe:
 *   - occupancy uses analytic box / sphere primitives instead of
 *     EnvironmentBase::CheckCollision(cube) (third party, mod.cpp:520);
 *   - gettraj returns the waypoints in OpenRAVE's trajectory XML layout with a
 *     uniform deltatime; the LinearTrajectoryRetimer and the collision sweep
 *     (mod.cpp:2906-3006) are OpenRAVE's and are not reproduced -- the
 *     no_collision_* flags are accepted and have no effect;
 *   - starttraj is read in the XML layout gettraj emits (joint_values + deltatime
 *     groups) and sampled as mod.cpp:2375-2415 does;
 *   - floating_base + basegoal are supported for single runs with adofgoal (gettraj then
 *     carries an affine_transform group, x y z qw qx qy qz, as mod.cpp:2912-2956);
 *   - trajs_fileformstr writes the waypoints before every iteration in the same XML layout
 *     (joint_values group only, full precision, mod.cpp:2769-2795), one launch per iteration;
 *   - con_tsr 'start|end|all [manipee M | link L]' TSR, everyn_tsr TSR and start_tsr TSR are read with the
 *     reference's grammar, TSR text format and error texts (mod.cpp:1930-1997, 3068-3110) and run on the
 *     device (ocb_params.constraints); the stand-in environment learns link names and manipulators through
 *     ocb_env_set_link_names / ocb_env_add_manipulator / ocb_env_set_active_manipulator;
 *   - start_cost (a host callback evaluated every iteration, mod.cpp:1787-1792) is recognised and rejected
 *     with an error; ee_force, ee_force_at and ee_torque_weights are accepted and ignored, as in the
 *     reference (mod.cpp:1323).
 */
#ifndef ORCDCHOMP_B200_MODULE_H
#define ORCDCHOMP_B200_MODULE_H

#include <stddef.h>
#include "orcdchomp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ocb_env ocb_env;       /* stand-in for OpenRAVE::EnvironmentBase */
typedef struct ocb_module ocb_module; /* stand-in for the orcdchomp ModuleBase  */

int ocb_env_create(ocb_env **out);
int ocb_env_destroy(ocb_env *env);
/* a kinbody: name, world pose [x y z qx qy qz qw], geometry in the kinbody frame */
int ocb_env_add_kinbody(ocb_env *env, const char *name, const double pose[7],
                        const ocb_prim *prims, int n_prims);
int ocb_env_set_kinbody_pose(ocb_env *env, const char *name, const double pose[7]);
/* KinBody::Enable(): disabled bodies are ignored by the occupancy test */
int ocb_env_enable_kinbody(ocb_env *env, const char *name, int enabled);
/* a robot: name, description (deep-copied), current active-DOF values [n_dof].
 * A robot is also a kinbody (no geometry) so fields can be attached to it. */
int ocb_env_add_robot(ocb_env *env, const char *name, const ocb_robot *robot,
                      const double *active_dof_values);
int ocb_env_set_active_dof_values(ocb_env *env, const char *name, const double *values);
/* what the TSR constraint arguments of `create` look up on the robot (mod.cpp:1957-1977, 1384-1387):
 * KinBody::GetLink(name), RobotBase::GetManipulators() / GetActiveManipulator() with
 * GetEndEffector() and GetLocalToolTransform().  names: one per link of the description.  The first
 * manipulator added is the active one until ocb_env_set_active_manipulator. */
int ocb_env_set_link_names(ocb_env *env, const char *robot, const char *const *names, int n_names);
int ocb_env_add_manipulator(ocb_env *env, const char *robot, const char *name, int ee_link,
                            const double local_tool[7]);
int ocb_env_set_active_manipulator(ocb_env *env, const char *robot, const char *name);

/* RaveCreateModule(env, "orcdchomp") on GPU `device` */
int ocb_module_create(ocb_env *env, int device, ocb_module **out);
int ocb_module_destroy(ocb_module *m);
/* ModuleBase::SendCommand: returns 0 and the command's output text on success;
 * non-zero when the command threw (text of the exception in ocb_module_last_error()),
 * like the RuntimeError openravepy raises.  out may be NULL; *out_len receives
 * the full length even when it exceeds out_cap. */
int ocb_module_send_command(ocb_module *m, const char *cmd, char *out, size_t out_cap, size_t *out_len);
/* text of the exception thrown by the last failed command on this thread */
const char *ocb_module_last_error(void);
/* The <orcdchomp><spheres> block of a robot / kinbody XML text, read as the reference's XML reader
 * reads it (src/orcdchomp_kdata.cpp:65-98): every <sphere link=".." pos="x y z" radius="r"/> in
 * document order.  link_names [cap][64], pos [cap][3], radius [cap] (each may be NULL); *n_out = number
 * of spheres found (only the first cap are stored).  Returns 0, or -2 with the reason in err. */
int ocb_kdata_parse_spheres(const char *xml, int cap, char *link_names, double *pos, double *radius,
                            int *n_out, char *err, size_t err_cap);
/* A TSR in the text form of the create command, read as tsr_create_parse reads it
 * (src/orcdchomp_mod.cpp:3068-3110): "manipindex bodyandlink", T0_w as nine rotation entries by column and
 * three translation entries, the same for Tw_e, then Bw (6 x 2, rows x y z roll pitch yaw).  The two
 * transforms are returned as poses [x y z qx qy qz qw] (cd_kin_pose_from_dR).  0, or OCB_ERR_ARG when the
 * text does not hold exactly those 38 fields. */
int ocb_tsr_parse(const char *text, double T0w[7], double Twe[7], double Bw[12]);
/* numeric access to a run handle returned by create / createbatch ("%p" text) */
int ocb_module_run_batch(ocb_module *m, const char *handle, ocb_batch **batch);

#ifdef __cplusplus
}
#endif
#endif
