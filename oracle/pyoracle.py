"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Loads  oracle/build/liboracle_port.so  ("port": oracle/libcd_port.c +
oracle/orcdchomp_port.c) or  oracle/_ref/liboracle_ref.so  ("reference": the
unmodified /root/reference/src/libcd sources + oracle/orcdchomp_port.c).
May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs only; nothing under or_cdchomp_b200/ imports it.
"""
import ctypes as C
import glob
import os

import numpy as np

from or_cdchomp_b200.capi import (OcbParams, OcbPrim, OcbRobot, OcbSdf, as_f64, c_double_p, count_start_tsr,
                                  c_int_p, dptr)

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {
    "port": os.path.join(HERE, "build", "liboracle_port.so"),
    "reference": os.path.join(HERE, "_ref", "liboracle_ref.so"),
}
_libs = {}


def available(flavour):
    return os.path.exists(PATHS[flavour])


def _preload_openblas():
    """liboracle_ref.so carries an rpath to scipy.libs; if the wheel hash in the
    file name differs on another box, preload whatever is there."""
    try:
        import scipy
        pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so")
        for p in sorted(glob.glob(pat)):
            C.CDLL(os.path.realpath(p), mode=C.RTLD_GLOBAL)
            return
    except Exception:
        pass


def load(flavour="port"):
    if flavour in _libs:
        return _libs[flavour]
    path = PATHS[flavour]
    if not os.path.exists(path):
        raise OSError("%s missing: run `make -C oracle %s`" % (path, "port" if flavour == "port" else "ref"))
    if flavour == "reference":
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        _preload_openblas()
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp = C.c_void_p
    lib.orc_flavour.restype = C.c_char_p
    lib.orc_run_create.restype = C.c_int
    lib.orc_run_create.argtypes = [C.POINTER(OcbRobot), C.POINTER(OcbParams), C.c_int, C.POINTER(OcbSdf),
                                   c_double_p, c_double_p, C.c_uint, C.POINTER(vp)]
    lib.orc_run_set_traj.restype = None
    lib.orc_run_set_traj.argtypes = [vp, c_double_p]
    lib.orc_run_iterate.restype = C.c_int
    lib.orc_run_iterate.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p]
    lib.orc_run_get_traj.restype = None
    lib.orc_run_get_traj.argtypes = [vp, c_double_p]
    lib.orc_run_get_momentum.restype = None
    lib.orc_run_get_momentum.argtypes = [vp, c_double_p]
    lib.orc_run_hmc_next.restype = C.c_int
    lib.orc_run_hmc_next.argtypes = [vp]
    lib.orc_run_obstacle_gradient.restype = C.c_int
    lib.orc_run_obstacle_gradient.argtypes = [vp, c_double_p, c_double_p]
    lib.orc_run_sphere_positions.restype = C.c_int
    lib.orc_run_sphere_positions.argtypes = [vp, c_double_p, c_int_p]
    lib.orc_run_constraint_eval.restype = C.c_int
    lib.orc_run_constraint_eval.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p]
    lib.orc_run_destroy.restype = None
    lib.orc_run_destroy.argtypes = [vp]
    lib.orc_sdf_from_obsarray.restype = C.c_int
    lib.orc_sdf_from_obsarray.argtypes = [c_double_p, c_int_p, c_double_p, c_double_p]
    lib.orc_dt_sqeuc.restype = C.c_int
    lib.orc_dt_sqeuc.argtypes = [c_double_p, c_int_p, c_double_p, c_double_p]
    lib.orc_computedistancefield.restype = C.c_int
    lib.orc_computedistancefield.argtypes = [C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p, C.c_double,
                                             c_double_p, c_double_p]
    lib.orc_occupancy.restype = C.c_int
    lib.orc_occupancy.argtypes = [C.POINTER(OcbPrim), C.c_int, c_int_p, c_double_p, C.c_double, c_double_p]
    lib.orc_sdf_sample.restype = C.c_int
    lib.orc_sdf_sample.argtypes = [c_double_p, c_int_p, c_double_p, c_double_p, C.c_int, c_double_p,
                                   c_double_p, c_int_p]
    lib.orc_fk.restype = None
    lib.orc_fk.argtypes = [C.POINTER(OcbRobot), c_double_p, c_double_p]
    lib.orc_pose_block.restype = None
    lib.orc_pose_block.argtypes = [c_double_p, c_double_p, c_double_p]
    lib.orc_mt_seed.restype = None
    lib.orc_mt_seed.argtypes = [vp, C.c_uint]
    lib.orc_mt_next.restype = C.c_uint
    lib.orc_mt_next.argtypes = [vp]
    lib.orc_mt_uniform.restype = C.c_double
    lib.orc_mt_uniform.argtypes = [vp]
    lib.orc_mt_gaussian.restype = C.c_double
    lib.orc_mt_gaussian.argtypes = [vp, C.c_double]
    assert lib.orc_flavour().decode() == flavour
    _libs[flavour] = lib
    return lib


def debug_limit_rounds(reset=False):
    """port flavour only: the largest number of joint-limit rounds (chomp.c:608-655) any
    iteration has needed since the last reset."""
    lib = load("port")
    cnt = C.c_int.in_dll(lib, "orc_debug_max_limit_rounds")
    v = cnt.value
    if reset:
        cnt.value = 0
    return v


def best_flavour():
    """'reference' when the compiled reference is present, else 'port'."""
    return "reference" if available("reference") else "port"


def _i3(a):
    return (C.c_int * 3)(*[int(x) for x in a])


def _d3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


class Run:
    """One reference CHOMP run (struct run, src/orcdchomp_mod.cpp:887-966)."""

    def __init__(self, robot, params, sdfs, q_start, q_goal, seed=0, flavour="port"):
        self.lib = load(flavour)
        self.robot, self.params, self.sdfs = robot, params, list(sdfs)
        self.n = robot.n_dof + (7 if params.floating_base else 0)
        self.P = params.n_points
        self.m = self.P - 2 + count_start_tsr(params)  # start_tsr: mod.cpp:2316
        arr = (OcbSdf * len(self.sdfs))(*[s.struct for s in self.sdfs])
        self._arr = arr
        self.q_start, self.q_goal = as_f64(q_start), as_f64(q_goal)
        h = C.c_void_p()
        err = self.lib.orc_run_create(C.byref(robot.struct), C.byref(params), len(self.sdfs), arr,
                                      dptr(self.q_start), dptr(self.q_goal), int(seed), C.byref(h))
        if err:
            raise RuntimeError("orc_run_create failed: %d" % err)
        self.h = h

    def set_traj(self, traj):
        traj = as_f64(traj)
        assert traj.shape == (self.P, self.n)
        self.lib.orc_run_set_traj(self.h, dptr(traj))

    def iterate(self, n_iter, want_trace=False, want_grads=False):
        costs = np.zeros(3)
        trace = np.zeros((n_iter, 3)) if want_trace else None
        grads = np.zeros((n_iter, self.m, self.n)) if want_grads else None
        ret = self.lib.orc_run_iterate(self.h, n_iter, dptr(costs),
                                       dptr(trace) if want_trace and n_iter else None,
                                       dptr(grads) if want_grads and n_iter else None)
        return ret, costs, trace, grads

    def traj(self):
        out = np.zeros((self.P, self.n))
        self.lib.orc_run_get_traj(self.h, dptr(out))
        return out

    def momentum(self):
        out = np.zeros((self.m, self.n))
        self.lib.orc_run_get_momentum(self.h, dptr(out))
        return out

    def hmc_next(self):
        return self.lib.orc_run_hmc_next(self.h)

    def obstacle_gradient(self):
        g = np.zeros((self.m, self.n))
        c = np.zeros(self.m)
        self.lib.orc_run_obstacle_gradient(self.h, dptr(g), dptr(c))
        return g, c

    def constraint_eval(self, index, point):
        """value (k) and Jacobian (k x n) of constraint `index` at configuration `point`
        (con_tsr, src/orcdchomp_mod.cpp:1330-1497)"""
        point = as_f64(point).copy()
        val = np.zeros(6)
        jac = np.zeros((6, self.n))
        k = self.lib.orc_run_constraint_eval(self.h, int(index), dptr(point), dptr(val), dptr(jac))
        if k < 0:
            raise RuntimeError("orc_run_constraint_eval: %d" % k)
        return val[:k].copy(), jac.reshape(-1)[:k * self.n].reshape(k, self.n).copy()

    def sphere_positions(self):
        na = self.robot.n_spheres_active
        out = np.zeros((self.P, na, 3))
        k = C.c_int()
        self.lib.orc_run_sphere_positions(self.h, dptr(out), C.byref(k))
        assert k.value == na
        return out

    def close(self):
        if self.h:
            self.lib.orc_run_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sdf_from_obsarray(obs, lengths, flavour="port"):
    lib = load(flavour)
    obs = as_f64(obs)
    out = np.empty_like(obs)
    err = lib.orc_sdf_from_obsarray(dptr(obs), _i3(obs.shape), _d3(lengths), dptr(out))
    if err:
        raise RuntimeError("orc_sdf_from_obsarray: %d" % err)
    return out


def dt_sqeuc(func, lengths, flavour="port"):
    lib = load(flavour)
    func = as_f64(func)
    out = np.empty_like(func)
    err = lib.orc_dt_sqeuc(dptr(func), _i3(func.shape), _d3(lengths), dptr(out))
    if err:
        raise RuntimeError("orc_dt_sqeuc: %d" % err)
    return out


def computedistancefield(prim_array, n_prims, sizes, lengths, cube_extent, flavour="port", want_sdf=True):
    lib = load(flavour)
    obs = np.empty(tuple(int(s) for s in sizes))
    sdf = np.empty_like(obs) if want_sdf else None
    err = lib.orc_computedistancefield(prim_array, n_prims, _i3(sizes), _d3(lengths), float(cube_extent),
                                       dptr(obs), dptr(sdf) if want_sdf else None)
    if err:
        raise RuntimeError("orc_computedistancefield: %d" % err)
    return obs, sdf


def occupancy(prim_array, n_prims, sizes, lengths, cube_extent, flavour="port"):
    lib = load(flavour)
    out = np.empty(tuple(int(s) for s in sizes))
    lib.orc_occupancy(prim_array, n_prims, _i3(sizes), _d3(lengths), float(cube_extent), dptr(out))
    return out


def sdf_sample(data, lengths, points, flavour="port"):
    lib = load(flavour)
    data, points = as_f64(data), as_f64(points).reshape(-1, 3)
    k = len(points)
    vals, grads = np.zeros(k), np.zeros((k, 3))
    errs = np.zeros(k, dtype=np.int32)
    lib.orc_sdf_sample(dptr(data), _i3(data.shape), _d3(lengths), dptr(points), k, dptr(vals), dptr(grads),
                       errs.ctypes.data_as(c_int_p))
    return vals, grads, errs


def fk(robot, q, flavour="port"):
    lib = load(flavour)
    q = as_f64(q)
    out = np.zeros((robot.n_links, 7))
    lib.orc_fk(C.byref(robot.struct), dptr(q), dptr(out))
    return out


def pose_block(pose, v, flavour="port"):
    """3 x 7 pose block of a sphere Jacobian in floating-base mode (mod.cpp:1050-1080), 0.01 scaling included"""
    lib = load(flavour)
    pose, v = as_f64(pose), as_f64(v)
    out = np.zeros((3, 7))
    lib.orc_pose_block(dptr(pose), dptr(v), dptr(out))
    return out


class MT:
    """gsl_rng_mt19937 restatement (oracle/orcdchomp_port.c)."""

    def __init__(self, seed, flavour="port"):
        self.lib = load(flavour)
        self.buf = C.create_string_buffer(625 * 4)
        self.lib.orc_mt_seed(self.buf, int(seed))

    def next(self):
        return self.lib.orc_mt_next(self.buf)

    def uniform(self):
        return self.lib.orc_mt_uniform(self.buf)

    def gaussian(self, sigma):
        return self.lib.orc_mt_gaussian(self.buf, float(sigma))
