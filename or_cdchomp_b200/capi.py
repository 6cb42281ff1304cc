"""ctypes view of the C ABI declared in include/orcdchomp_b200.h.

This module only describes the boundary (struct layouts, argument types) and
loads the in-tree shared library built from or_cdchomp_b200/csrc.  It contains no
numerics and no fallback: if the library is missing, loading raises.

Reference interfaces mirrored by the structs (paths relative to the reference
root): struct run / run_sphere / run_rsdf (src/orcdchomp_mod.cpp:850-966),
struct cd_grid (src/libcd/grid.h:29-41), `create` parameters
(src/orcdchomp_mod.cpp:1818-1848).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "liborcdchomp_b200.so")

OCB_OK = 0
OCB_ERR_ALLOC = -1
OCB_ERR_ARG = -2
OCB_ERR_CUDA = -3
OCB_ERR_NODEVICE = -4
OCB_ERR_JLIMIT = -5

JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2
PRIM_BOX, PRIM_SPHERE, PRIM_TRIANGLE = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_uint_p = C.POINTER(C.c_uint)


class OcbRobot(C.Structure):
    _fields_ = [
        ("n_links", C.c_int),
        ("parent", c_int_p),
        ("pose_parent", c_double_p),
        ("joint_type", c_int_p),
        ("axis", c_double_p),
        ("dof_index", c_int_p),
        ("dof_coeff", c_double_p),
        ("base_pose", C.c_double * 7),
        ("n_dof", C.c_int),
        ("limit_lower", c_double_p),
        ("limit_upper", c_double_p),
        ("n_spheres", C.c_int),
        ("sphere_link", c_int_p),
        ("sphere_pos", c_double_p),
        ("sphere_radius", c_double_p),
    ]


class OcbSdf(C.Structure):
    _fields_ = [
        ("sizes", C.c_int * 3),
        ("lengths", C.c_double * 3),
        ("pose_world_gsdf", C.c_double * 7),
        ("data", c_double_p),
    ]


class OcbConstraint(C.Structure):
    """ocb_constraint (include/orcdchomp_b200.h): one TSR hard constraint."""
    _fields_ = [
        ("where", C.c_int),
        ("link", C.c_int),
        ("pose_link_ee", C.c_double * 7),
        ("T0w", C.c_double * 7),
        ("Twe", C.c_double * 7),
        ("Bw", (C.c_double * 2) * 6),
    ]


CON_START, CON_END, CON_ALL, CON_START_TSR = 0, 1, 2, 3


class OcbParams(C.Structure):
    _fields_ = [
        ("n_points", C.c_int),
        ("derivative", C.c_int),
        ("lambda_", C.c_double),
        ("use_momentum", C.c_int),
        ("use_hmc", C.c_int),
        ("hmc_resample_lambda", C.c_double),
        ("epsilon", C.c_double),
        ("epsilon_self", C.c_double),
        ("obs_factor", C.c_double),
        ("obs_factor_self", C.c_double),
        ("floating_base", C.c_int),
        ("n_constraints", C.c_int),
        ("constraints", C.POINTER(OcbConstraint)),
    ]


class OcbPrim(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("pose", C.c_double * 7),
        ("extents", C.c_double * 3),
    ]


def default_params(**kw):
    """`create` defaults (src/orcdchomp_mod.cpp:1818-1848, 1875)."""
    p = OcbParams(
        n_points=101, derivative=1, lambda_=10.0, use_momentum=0, use_hmc=0,
        hmc_resample_lambda=0.02, epsilon=0.1, epsilon_self=0.04,
        obs_factor=200.0, obs_factor_self=10.0)
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if k == "constraints":
            set_constraints(p, v)
            continue
        if not hasattr(p, k):
            raise TypeError("unknown parameter %r" % k)
        setattr(p, k, v)
    return p


_IDENTITY_POSE = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)


def make_constraint(where, link, Bw, T0w=_IDENTITY_POSE, Twe=_IDENTITY_POSE, pose_link_ee=_IDENTITY_POSE):
    """One ocb_constraint.  `where`: CON_START / CON_END / CON_ALL / CON_START_TSR (or their names);
    `Bw`: 6 x 2 bounds, rows x y z roll pitch yaw -- a row of zeros holds that entry at zero."""
    if isinstance(where, str):
        where = {"start": CON_START, "end": CON_END, "all": CON_ALL, "start_tsr": CON_START_TSR}[where]
    c = OcbConstraint()
    c.where, c.link = int(where), int(link)
    for name, val in (("pose_link_ee", pose_link_ee), ("T0w", T0w), ("Twe", Twe)):
        v = np.asarray(val, dtype=np.float64).reshape(7)
        for i in range(7):
            getattr(c, name)[i] = v[i]
    b = np.asarray(Bw, dtype=np.float64).reshape(6, 2)
    for i in range(6):
        c.Bw[i][0], c.Bw[i][1] = b[i, 0], b[i, 1]
    return c


def set_constraints(params, constraints):
    """Attach a list of OcbConstraint to an OcbParams (the array is kept alive on the object)."""
    constraints = list(constraints or [])
    arr = (OcbConstraint * max(len(constraints), 1))(*constraints)
    params._constraints_keepalive = arr
    params.n_constraints = len(constraints)
    params.constraints = C.cast(arr, C.POINTER(OcbConstraint)) if constraints else None
    return params


def count_start_tsr(params):
    return sum(1 for i in range(params.n_constraints) if params.constraints[i].where == CON_START_TSR)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


class RobotDesc:
    """Owns the numpy arrays behind an OcbRobot struct (keeps them alive)."""

    def __init__(self, names, parent, pose_parent, joint_type, axis, dof_index,
                 dof_coeff, base_pose, limit_lower, limit_upper, sphere_link,
                 sphere_pos, sphere_radius):
        self.names = list(names)
        self.parent = as_i32(parent)
        self.pose_parent = as_f64(pose_parent).reshape(-1, 7)
        self.joint_type = as_i32(joint_type)
        self.axis = as_f64(axis).reshape(-1, 3)
        self.dof_index = as_i32(dof_index)
        self.dof_coeff = as_f64(dof_coeff).reshape(-1, 2)
        self.base_pose = as_f64(base_pose)
        self.limit_lower = as_f64(limit_lower)
        self.limit_upper = as_f64(limit_upper)
        self.sphere_link = as_i32(sphere_link)
        self.sphere_pos = as_f64(sphere_pos).reshape(-1, 3)
        self.sphere_radius = as_f64(sphere_radius)
        self.n_links = len(self.parent)
        self.n_dof = len(self.limit_lower)
        self.n_spheres = len(self.sphere_radius)
        s = OcbRobot()
        s.n_links = self.n_links
        s.parent = iptr(self.parent)
        s.pose_parent = dptr(self.pose_parent)
        s.joint_type = iptr(self.joint_type)
        s.axis = dptr(self.axis)
        s.dof_index = iptr(self.dof_index)
        s.dof_coeff = dptr(self.dof_coeff)
        for i in range(7):
            s.base_pose[i] = float(self.base_pose[i])
        s.n_dof = self.n_dof
        s.limit_lower = dptr(self.limit_lower)
        s.limit_upper = dptr(self.limit_upper)
        s.n_spheres = self.n_spheres
        s.sphere_link = iptr(self.sphere_link)
        s.sphere_pos = dptr(self.sphere_pos)
        s.sphere_radius = dptr(self.sphere_radius)
        self.struct = s

    def link_is_active(self, link):
        a = link
        while a > 0:
            if self.dof_index[a] >= 0 and self.joint_type[a] != JOINT_FIXED:
                return True
            a = int(self.parent[a])
        return False

    @property
    def n_spheres_active(self):
        return sum(1 for l in self.sphere_link if self.link_is_active(int(l)))


class SdfDesc:
    """Owns the grid behind an OcbSdf struct."""

    def __init__(self, data, lengths, pose_world_gsdf):
        self.data = as_f64(data)
        assert self.data.ndim == 3
        self.lengths = as_f64(lengths)
        self.pose = as_f64(pose_world_gsdf)
        s = OcbSdf()
        for i in range(3):
            s.sizes[i] = self.data.shape[i]
            s.lengths[i] = float(self.lengths[i])
        for i in range(7):
            s.pose_world_gsdf[i] = float(self.pose[i])
        s.data = dptr(self.data)
        self.struct = s


def make_prims(prims):
    """prims: list of ('box', pose7, half_extents3) / ('sphere', centre3, radius) / ('tri', v0, v1, v2)."""
    arr = (OcbPrim * max(1, len(prims)))()
    for i, p in enumerate(prims):
        if p[0] == "box":
            arr[i].type = PRIM_BOX
            for k in range(7):
                arr[i].pose[k] = float(p[1][k])
            for k in range(3):
                arr[i].extents[k] = float(p[2][k])
        elif p[0] == "sphere":
            arr[i].type = PRIM_SPHERE
            for k in range(3):
                arr[i].pose[k] = float(p[1][k])
            arr[i].pose[6] = 1.0
            arr[i].extents[0] = float(p[2])
        elif p[0] == "tri":
            arr[i].type = PRIM_TRIANGLE
            v = [float(x) for vert in p[1:4] for x in vert]
            for k in range(7):
                arr[i].pose[k] = v[k]
            arr[i].extents[0], arr[i].extents[1] = v[7], v[8]
        else:
            raise ValueError(p[0])
    return arr


_lib = None


def load_library(path=None):
    """Load liborcdchomp_b200.so (built in-tree by __graft_entry__.build()).

    Raises OSError if it has not been built: there is no Python or CPU fallback
    for the hot path.
    """
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise OSError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    _declare(lib)
    if path == LIB_PATH:
        _lib = lib
    return lib


EXPORTS = {
    # name: (restype, argtypes)
    "ocb_last_error": (C.c_char_p, []),
    "ocb_version": (C.c_char_p, []),
    "ocb_params_default": (None, [C.POINTER(OcbParams)]),
    "ocb_engine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ocb_engine_destroy": (C.c_int, [C.c_void_p]),
    "ocb_engine_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ocb_engine_trim": (C.c_int, [C.c_void_p]),
    "ocb_engine_enable_jit": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_batch_uses_jit": (C.c_int, [C.c_void_p]),
    "ocb_batch_tile_width": (C.c_int, [C.c_void_p]),
    "ocb_engine_sync": (C.c_int, [C.c_void_p]),
    "ocb_sdf_upload": (C.c_int, [C.c_void_p, C.POINTER(OcbSdf), c_int_p]),
    "ocb_sdf_adopt_device": (C.c_int, [C.c_void_p, c_int_p, c_double_p, c_double_p, C.c_void_p, c_int_p]),
    "ocb_sdf_remove": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_sdf_build_host": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, c_double_p]),
    "ocb_sdf_build_device": (C.c_int, [C.c_void_p, C.c_void_p, c_int_p, c_double_p, C.c_void_p]),
    "ocb_engine_force_general_sdf": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_dt_sqeuc_device": (C.c_int, [C.c_void_p, C.c_void_p, c_int_p, c_double_p, C.c_void_p]),
    "ocb_dt_sqeuc_host": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, c_double_p]),
    "ocb_occupancy_device": (C.c_int, [C.c_void_p, C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p, C.c_double, C.c_void_p]),
    "ocb_flood_relabel_device": (C.c_int, [C.c_void_p, C.c_void_p, c_int_p, C.c_size_t]),
    "ocb_flood_relabel_host": (C.c_int, [C.c_void_p, c_double_p, c_int_p, C.c_size_t]),
    "ocb_computedistancefield_host": (C.c_int, [C.c_void_p, C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p, C.c_double, c_double_p, c_double_p]),
    "ocb_computedistancefield_resident": (C.c_int, [C.c_void_p, C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p,
                                                    C.c_double, c_double_p, c_int_p]),
    "ocb_sdf_build_resident": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, c_double_p, c_int_p]),
    "ocb_sdf_download": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "ocb_sdf_alias": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p]),
    "ocb_batch_create": (C.c_int, [C.c_void_p, C.POINTER(OcbRobot), C.POINTER(OcbParams), C.c_int, c_int_p, C.c_int, c_double_p, c_double_p, c_uint_p, C.POINTER(C.c_void_p)]),
    "ocb_batch_reset": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_uint_p]),
    "ocb_batch_set_traj": (C.c_int, [C.c_void_p, c_double_p]),
    "ocb_sdf_sample_host": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, c_int_p]),
    "ocb_batch_iterate": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]),
    "ocb_batch_iterate_async": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_batch_iterate_from": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]),
    "ocb_batch_iterate_from_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ocb_batch_get_costs": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    "ocb_batch_set_momentum": (C.c_int, [C.c_void_p, c_double_p, c_int_p]),
    "ocb_batch_get_momentum": (C.c_int, [C.c_void_p, c_double_p, c_int_p]),
    "ocb_batch_set_lambda": (C.c_int, [C.c_void_p, C.c_double]),
    "ocb_batch_get_iterations": (C.c_int, [C.c_void_p, c_int_p]),
    "ocb_batch_get_limit_rounds": (C.c_int, [C.c_void_p, c_int_p]),
    "ocb_batch_get_constraint_skips": (C.c_int, [C.c_void_p, c_int_p]),
    "ocb_batch_enable_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_batch_get_trace": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "ocb_batch_get_traj": (C.c_int, [C.c_void_p, c_double_p]),
    "ocb_batch_capture_gradient": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_batch_get_gradient": (C.c_int, [C.c_void_p, c_double_p]),
    "ocb_batch_best": (C.c_int, [C.c_void_p, c_int_p, c_double_p]),
    "ocb_batch_destroy": (C.c_int, [C.c_void_p]),
    "ocb_batch_dims": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p]),
    "ocb_batch_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "ocb_batch_copy_run_traj_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ocb_engine_launch_count": (C.c_long, [C.c_void_p]),
    "ocb_debug_jit_robot_header": (C.c_long, [C.POINTER(OcbRobot), C.POINTER(OcbParams), C.c_char_p, C.c_size_t]),
    "ocb_debug_metric": (C.c_int, [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p,
                                   c_double_p]),
    # several GPUs in one process (csrc/ocb_multi.cpp)
    "ocb_multi_last_error": (C.c_char_p, []),
    "ocb_multi_create": (C.c_int, [C.c_int, c_int_p, C.POINTER(C.c_void_p)]),
    "ocb_multi_destroy": (C.c_int, [C.c_void_p]),
    "ocb_multi_device_count": (C.c_int, [C.c_void_p]),
    "ocb_multi_uses_nccl": (C.c_int, [C.c_void_p]),
    "ocb_multi_engine": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ocb_multi_enable_jit": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_multi_sdf_upload": (C.c_int, [C.c_void_p, C.POINTER(OcbSdf), c_int_p]),
    "ocb_multi_computedistancefield_resident": (C.c_int, [C.c_void_p, C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p,
                                                          C.c_double, c_double_p, c_int_p]),
    "ocb_multi_sdf_remove": (C.c_int, [C.c_void_p, C.c_int]),
    "ocb_multi_batch_create": (C.c_int, [C.c_void_p, C.POINTER(OcbRobot), C.POINTER(OcbParams), C.c_int, c_int_p, C.c_int,
                                         c_double_p, c_double_p, c_uint_p, C.POINTER(C.c_void_p)]),
    "ocb_multi_batch_dims": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p]),
    "ocb_multi_batch_iterate": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]),
    "ocb_multi_batch_get_traj": (C.c_int, [C.c_void_p, c_double_p]),
    "ocb_multi_batch_best": (C.c_int, [C.c_void_p, c_int_p, c_double_p, c_double_p]),
    "ocb_multi_best_traj_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ocb_multi_batch_destroy": (C.c_int, [C.c_void_p]),
}


def _declare(lib):
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args


class OcbError(RuntimeError):
    pass


def check(lib, code, what):
    if code != OCB_OK:
        msg = lib.ocb_last_error()
        raise OcbError("%s failed with code %d: %s" % (what, code, (msg or b"").decode()))
