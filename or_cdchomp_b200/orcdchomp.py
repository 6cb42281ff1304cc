"""Python face of the module commands -- the counterpart of the reference's
pythonsrc/orcdchomp/orcdchomp.py (bind, viewspheres, computedistancefield,
addfield_fromobsarray, viewfields, removefield, create, iterate, gettraj, destroy,
runchomp; reference lines 28-220).  Every function only builds the command string
the reference builds (same keywords, same number formatting: %f, lambda %0.04f)
and hands it to `mod.SendCommand(cmd, releasegil)`.

`Environment` / `Module` wrap the C ABI of include/orcdchomp_b200_module.h and play
the roles of openravepy.Environment / RaveCreateModule(env, 'orcdchomp').
gettraj returns the waypoints as a numpy array [n_points, n_dof] (the reference
returns an openravepy trajectory object built from the same XML).
"""
import ctypes as C
import re
import types

import numpy as np

from . import capi
from .capi import OcbPrim, OcbRobot, as_f64, c_double_p, dptr

_MODULE_EXPORTS = {
    "ocb_env_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ocb_env_destroy": (C.c_int, [C.c_void_p]),
    "ocb_env_add_kinbody": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p, C.POINTER(OcbPrim), C.c_int]),
    "ocb_env_set_kinbody_pose": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p]),
    "ocb_env_enable_kinbody": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ocb_env_add_robot": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(OcbRobot), c_double_p]),
    "ocb_env_set_active_dof_values": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p]),
    "ocb_env_set_link_names": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]),
    "ocb_env_add_manipulator": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, c_double_p]),
    "ocb_env_set_active_manipulator": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "ocb_module_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ocb_module_destroy": (C.c_int, [C.c_void_p]),
    "ocb_module_send_command": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ocb_module_last_error": (C.c_char_p, []),
    "ocb_module_run_batch": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "ocb_tsr_parse": (C.c_int, [C.c_char_p, c_double_p, c_double_p, c_double_p]),
    "ocb_kdata_parse_spheres": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, c_double_p, c_double_p, capi.c_int_p,
                                          C.c_char_p, C.c_size_t]),
}


def _lib():
    lib = capi.load_library()
    if not getattr(lib, "_ocb_module_declared", False):
        for name, (res, args) in _MODULE_EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib._ocb_module_declared = True
    return lib


class _Named:
    def __init__(self, name):
        self._name = name

    def GetName(self):
        return self._name


class Environment:
    """Stand-in for the parts of openravepy.Environment the module commands read."""

    def __init__(self):
        self.lib = _lib()
        h = C.c_void_p()
        if self.lib.ocb_env_create(C.byref(h)):
            raise RuntimeError("ocb_env_create failed")
        self.h = h
        self._keep = []

    def AddKinBody(self, name, pose, prims):
        """prims: list of ('box', pose7, half_extents) / ('sphere', centre3, radius) in the body frame."""
        arr = capi.make_prims(prims)
        pose = as_f64(pose)
        if self.lib.ocb_env_add_kinbody(self.h, name.encode(), dptr(pose), arr, len(prims)):
            raise RuntimeError("could not add kinbody %r" % name)
        return _Named(name)

    def AddRobot(self, name, robot_desc, active_dof_values):
        q = as_f64(active_dof_values)
        self._keep.append(robot_desc)
        if self.lib.ocb_env_add_robot(self.h, name.encode(), C.byref(robot_desc.struct), dptr(q)):
            raise RuntimeError("could not add robot %r" % name)
        names = [s.encode() for s in getattr(robot_desc, "names", [])]
        if len(names) == robot_desc.struct.n_links:
            arr = (C.c_char_p * len(names))(*names)
            self.lib.ocb_env_set_link_names(self.h, name.encode(), arr, len(names))
        return _Named(name)

    def AddManipulator(self, robot, name, ee_link, local_tool=None):
        """RobotBase manipulator: end-effector link (name or index) and GetLocalToolTransform();
        the first one added is the active manipulator"""
        rname = _name(robot)
        if isinstance(ee_link, str):
            desc = [d for d in self._keep if ee_link in getattr(d, "names", [])]
            ee_link = desc[-1].names.index(ee_link)
        tool = as_f64(local_tool if local_tool is not None else [0, 0, 0, 0, 0, 0, 1])
        if self.lib.ocb_env_add_manipulator(self.h, rname.encode(), name.encode(), int(ee_link), dptr(tool)):
            raise RuntimeError("could not add manipulator %r" % name)

    def SetActiveManipulator(self, robot, name):
        if self.lib.ocb_env_set_active_manipulator(self.h, _name(robot).encode(), name.encode()):
            raise RuntimeError("no such manipulator")

    def SetTransform(self, body, pose):
        pose = as_f64(pose)
        if self.lib.ocb_env_set_kinbody_pose(self.h, _name(body).encode(), dptr(pose)):
            raise RuntimeError("no such kinbody")

    def Enable(self, body, enabled):
        if self.lib.ocb_env_enable_kinbody(self.h, _name(body).encode(), int(bool(enabled))):
            raise RuntimeError("no such kinbody")

    def SetActiveDOFValues(self, robot, values):
        q = as_f64(values)
        if self.lib.ocb_env_set_active_dof_values(self.h, _name(robot).encode(), dptr(q)):
            raise RuntimeError("no such robot")

    def close(self):
        if self.h:
            self.lib.ocb_env_destroy(self.h)
            self.h = None


class Module:
    """RaveCreateModule(env, 'orcdchomp'): SendCommand(cmd) -> output text, RuntimeError on failure."""

    def __init__(self, env, device=0):
        self.lib = _lib()
        self.env = env
        h = C.c_void_p()
        rc = self.lib.ocb_module_create(env.h, int(device), C.byref(h))
        if rc:
            raise RuntimeError("no orcdchomp module: %s" % self.lib.ocb_last_error().decode())
        self.h = h
        bind(self)

    def GetEnv(self):
        return self.env

    def SendCommand(self, cmd, releasegil=False):
        n = C.c_size_t()
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            rc = self.lib.ocb_module_send_command(self.h, cmd.encode(), buf, cap, C.byref(n))
            if rc != 0:
                raise RuntimeError(self.lib.ocb_module_last_error().decode())
            if n.value < cap:
                return buf.value.decode()
            if not cmd.startswith(("gettraj", "view")):  # only read-only commands may be replayed
                raise RuntimeError("command output of %d bytes was truncated" % n.value)
            cap = n.value + 1

    def batch_handle(self, run):
        b = C.c_void_p()
        if self.lib.ocb_module_run_batch(self.h, str(run).encode(), C.byref(b)):
            raise RuntimeError(self.lib.ocb_module_last_error().decode())
        return b

    def close(self):
        if self.h:
            self.lib.ocb_module_destroy(self.h)
            self.h = None


class TSR:
    """A task space region in the text form `create` reads (tsr_create_parse, src/orcdchomp_mod.cpp:3068-3110):
    manipulator index, body-and-link name, T0_w and Tw_e each as rotation matrix column by column followed by the
    translation, then the bounds Bw (6 x 2, rows x y z roll pitch yaw).  T0_w / Tw_e: 4 x 4 matrices or poses
    [x y z qx qy qz qw].  A row of Bw that is all zero makes that entry a hard constraint."""

    def __init__(self, T0_w=None, Tw_e=None, Bw=None, manipindex=0, bodyandlink="NULL"):
        self.T0_w, self.Tw_e = self._matrix(T0_w), self._matrix(Tw_e)
        self.Bw = np.zeros((6, 2)) if Bw is None else np.asarray(Bw, dtype=float).reshape(6, 2)
        self.manipindex, self.bodyandlink = int(manipindex), bodyandlink

    @staticmethod
    def _matrix(T):
        if T is None:
            return np.eye(4)
        T = np.asarray(T, dtype=float)
        if T.shape == (4, 4):
            return T
        x, y, z, w = T[3:7]
        M = np.eye(4)
        M[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        M[:3, 3] = T[:3]
        return M

    def serialize(self, *unused):
        out = ["%d" % self.manipindex, self.bodyandlink]
        for M in (self.T0_w, self.Tw_e):
            out += [repr(float(M[r, c])) for c in range(3) for r in range(3)]
            out += [repr(float(M[r, 3])) for r in range(3)]
        out += [repr(float(v)) for v in self.Bw.reshape(-1)]
        return " ".join(out)


def parse_tsr(text):
    """(T0w pose, Twe pose, Bw 6x2) of a TSR text, as the create command reads it (ocb_tsr_parse)"""
    T0w, Twe, Bw = np.zeros(7), np.zeros(7), np.zeros(12)
    if _lib().ocb_tsr_parse(text.encode(), dptr(T0w), dptr(Twe), dptr(Bw)):
        raise ValueError("Cannot parse TSR!")
    return T0w, Twe, Bw.reshape(6, 2)


def _name(obj):
    return obj.GetName() if hasattr(obj, "GetName") else obj


def shquot(s):
    return "'" + s.replace("'", "'\\''") + "'"


def _vec(values):
    return shquot(" ".join(str(v) for v in values))


def bind(mod):
    for fn in (viewspheres, computedistancefield, addfield_fromobsarray, viewfields, removefield, create,
               createbatch, iterate, gettraj, destroy, runchomp):
        setattr(mod, fn.__name__, types.MethodType(fn, mod))


def viewspheres(mod, robot=None, releasegil=False):
    cmd = "viewspheres"
    if robot is not None:
        cmd += " robot %s" % shquot(_name(robot))
    return mod.SendCommand(cmd, releasegil)


def computedistancefield(mod, kinbody=None, cube_extent=None, aabb_padding=None, cache_filename=None,
                         require_cache=None, releasegil=False):
    parts = ["computedistancefield"]
    if kinbody is not None:
        parts.append("kinbody %s" % shquot(_name(kinbody)))
    if cube_extent is not None:
        parts.append("cube_extent %f" % cube_extent)
    if aabb_padding is not None:
        parts.append("aabb_padding %f" % aabb_padding)
    if cache_filename is not None:
        parts.append("cache_filename %s" % shquot(cache_filename))
    if require_cache:
        parts.append("require_cache")
    return mod.SendCommand(" ".join(parts), releasegil)


def addfield_fromobsarray(mod, kinbody=None, obsarray=None, sizes=None, lengths=None, pose=None, releasegil=False):
    parts = ["addfield_fromobsarray"]
    if kinbody is not None:
        parts.append("kinbody %s" % shquot(_name(kinbody)))
    if obsarray is not None:
        parts.append("obsarray %s" % obsarray)
    if sizes is not None:
        parts.append("sizes %s" % _vec(sizes))
    if lengths is not None:
        parts.append("lengths %s" % _vec(lengths))
    if pose is not None:
        parts.append("pose %s" % _vec(pose))
    return mod.SendCommand(" ".join(parts), releasegil)


def viewfields(mod, releasegil=False):
    return mod.SendCommand("viewfields", releasegil)


def removefield(mod, kinbody=None, releasegil=False):
    cmd = "removefield"
    if kinbody is not None:
        cmd += " kinbody %s" % shquot(_name(kinbody))
    return mod.SendCommand(cmd, releasegil)


# (keyword, kind) in the order the reference emits them (orcdchomp.py:106-168)
_CREATE_SCALARS = [
    ("n_points", "n_points %d"), ("derivative", "derivative %d"),
]
_CREATE_TAIL = [
    ("hmc_resample_lambda", "hmc_resample_lambda %f"), ("seed", "seed %d"), ("epsilon", "epsilon %f"),
    ("epsilon_self", "epsilon_self %f"), ("obs_factor", "obs_factor %f"), ("obs_factor_self", "obs_factor_self %f"),
]


def _create_parts(robot, adofgoal, basegoal, floating_base, lambda_, n_points, derivative, use_momentum, use_hmc,
                  dat_filename, opts):
    parts = []
    if robot is not None:
        parts.append("robot %s" % shquot(_name(robot)))
    if adofgoal is not None:
        parts.append("adofgoal %s" % _vec(adofgoal))
    if basegoal is not None:
        parts.append("basegoal %s" % _vec(basegoal))
    if floating_base:
        parts.append("floating_base")
    if lambda_ is not None:
        parts.append("lambda %0.04f" % lambda_)
    if n_points is not None:
        parts.append("n_points %d" % n_points)
    if derivative is not None:
        parts.append("derivative %d" % derivative)
    if use_momentum:
        parts.append("use_momentum")
    if use_hmc:
        parts.append("use_hmc")
    for key, fmt in _CREATE_TAIL:
        if opts.get(key) is not None:
            parts.append(fmt % opts[key])
    if dat_filename is not None:
        parts.append("dat_filename %s" % shquot(dat_filename))
    return parts


def create(mod, robot=None, adofgoal=None, basegoal=None, floating_base=None, lambda_=None, starttraj=None,
           n_points=None, con_tsr=None, con_tsrs=None, start_tsr=None, start_cost=None, everyn_tsr=None,
           use_momentum=None, use_hmc=None, hmc_resample_lambda=None, seed=None, epsilon=None, epsilon_self=None,
           obs_factor=None, obs_factor_self=None, no_report_cost=None, dat_filename=None, releasegil=False,
           derivative=None, **kwargs):
    opts = dict(hmc_resample_lambda=hmc_resample_lambda, seed=seed, epsilon=epsilon, epsilon_self=epsilon_self,
                obs_factor=obs_factor, obs_factor_self=obs_factor_self)
    parts = ["create"] + _create_parts(robot, adofgoal, basegoal, floating_base, lambda_, n_points, derivative,
                                       use_momentum, use_hmc, dat_filename, opts)
    for key, val in (("starttraj", starttraj), ("start_tsr", start_tsr), ("everyn_tsr", everyn_tsr)):
        if val is not None:
            parts.append("%s %s" % (key, shquot(val.serialize(0) if hasattr(val, "serialize") else str(val))))
    if con_tsr is not None:
        con_tsrs = [con_tsr] + list(con_tsrs or [])
    for sub in con_tsrs or []:
        parts.append("con_tsr '%s' '%s'" % (sub[0], sub[1].serialize() if hasattr(sub[1], "serialize") else sub[1]))
    if start_cost is not None:
        parts.append("start_cost '%s'" % (start_cost if isinstance(start_cost, str) else "%s %s" % tuple(start_cost)))
    if no_report_cost:
        parts.append("no_report_cost")  # rejected by the C++ side, as in the reference
    return mod.SendCommand(" ".join(parts), releasegil)


def createbatch(mod, robot=None, adofgoals=None, adofstarts=None, seeds=None, lambda_=None, n_points=None,
                derivative=None, use_momentum=None, use_hmc=None, hmc_resample_lambda=None, epsilon=None,
                epsilon_self=None, obs_factor=None, obs_factor_self=None, releasegil=False):
    """Extension: R runs behind one handle.  adofgoals / adofstarts: [R, n_dof] float64, seeds: [R] uint32;
    the arrays are passed by address, like obsarray in addfield_fromobsarray."""
    goals = as_f64(adofgoals)
    opts = dict(hmc_resample_lambda=hmc_resample_lambda, seed=None, epsilon=epsilon, epsilon_self=epsilon_self,
                obs_factor=obs_factor, obs_factor_self=obs_factor_self)
    parts = ["createbatch"] + _create_parts(robot, None, None, None, lambda_, n_points, derivative, use_momentum,
                                            use_hmc, None, opts)
    parts.append("n_runs %d" % goals.shape[0])
    parts.append("adofgoals 0x%x" % goals.ctypes.data)
    keep = [goals]
    if adofstarts is not None:
        starts = as_f64(adofstarts)
        keep.append(starts)
        parts.append("adofstarts 0x%x" % starts.ctypes.data)
    if seeds is not None:
        sd = np.ascontiguousarray(seeds, dtype=np.uint32)
        keep.append(sd)
        parts.append("seeds 0x%x" % sd.ctypes.data)
    out = mod.SendCommand(" ".join(parts), releasegil)
    del keep
    return out


def iterate(mod, run=None, n_iter=None, max_time=None, trajs_fileformstr=None, cost=None, releasegil=False):
    parts = ["iterate"]
    if run is not None:
        parts.append("run %s" % run)
    if n_iter is not None:
        parts.append("n_iter %d" % n_iter)
    if max_time is not None:
        parts.append("max_time %f" % max_time)
    if trajs_fileformstr is not None:
        parts.append("trajs_fileformstr %s" % shquot(trajs_fileformstr))
    cost_data = mod.SendCommand(" ".join(parts), releasegil)
    if cost is not None:
        cost[0] = float(cost_data.split()[0])
    return cost_data


_DATA_RE = re.compile(r"<data count=\"(\d+)\">\s*(.*?)\s*</data>", re.S)
_DOF_RE = re.compile(r"joint_values [^\"]*\" offset=\"0\" dof=\"(\d+)\"")
_AFFINE_RE = re.compile(r"affine_transform [^\"]*\" offset=\"(\d+)\" dof=\"7\"")


def gettraj(mod, run=None, no_collision_check=None, no_collision_exception=None, no_collision_details=None,
            releasegil=False):
    parts = ["gettraj"]
    if run is not None:
        parts.append("run %s" % run)
    if no_collision_check:
        parts.append("no_collision_check")
    if no_collision_exception:
        parts.append("no_collision_exception")
    if no_collision_details:
        parts.append("no_collision_details")
    xml = mod.SendCommand(" ".join(parts), releasegil)
    n = int(_DOF_RE.search(xml).group(1))
    aff = _AFFINE_RE.search(xml)
    trajs = []
    for count, body in _DATA_RE.findall(xml):
        vals = np.array(body.split(), dtype=np.float64).reshape(int(count), n + 1 + (7 if aff else 0))
        if aff:
            # floating base: rows as the engine holds them, [x y z qx qy qz qw, active dofs]
            o = int(aff.group(1))
            pose = vals[:, [o, o + 1, o + 2, o + 4, o + 5, o + 6, o + 3]]
            trajs.append(np.concatenate([pose, vals[:, :n]], axis=1))
        else:
            trajs.append(vals[:, :n])
    return trajs[0] if len(trajs) == 1 else np.array(trajs)


def destroy(mod, run=None, releasegil=False):
    cmd = "destroy"
    if run is not None:
        cmd += " run %s" % run
    return mod.SendCommand(cmd, releasegil)


def runchomp(mod, n_iter=None, max_time=None, trajs_fileformstr=None, cost=None, no_collision_check=None,
             no_collision_exception=None, no_collision_details=None, releasegil=False, **kwargs):
    """create -> iterate -> gettraj -> destroy (orcdchomp.py:205-220 in the reference)."""
    run = create(mod, releasegil=releasegil, **kwargs)
    try:
        iterate(mod, run=run, n_iter=n_iter, max_time=max_time, trajs_fileformstr=trajs_fileformstr, cost=cost,
                releasegil=releasegil)
        traj = gettraj(mod, run=run, no_collision_check=no_collision_check,
                       no_collision_exception=no_collision_exception, no_collision_details=no_collision_details,
                       releasegil=releasegil)
    finally:
        destroy(mod, run=run, releasegil=releasegil)
    return traj
