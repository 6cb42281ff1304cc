"""ctypes view of libcd_b200.so: the reference's own grid entry points
(`cd_grid_double_bin_sdf`, `cd_grid_double_dt_sqeuc`, `cd_grid_double_sedt`; src/libcd/grid.h:84-93)
with libcd's `struct cd_grid` layout (grid.h:29-41), backed by the GPU engine.

This is how a C caller holding `struct cd_grid *` uses the engine without knowing about it;
the Python helpers here exist for the parity tests and mirror what such a caller does
(build a grid, call, read the result, release it with free() block by block as
cd_grid_destroy does, grid.c:134-143).
"""
import ctypes as C
import os

import numpy as np

from . import capi

LIBCD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcd_b200.so")

EXPORTS = ["cd_grid_double_bin_sdf", "cd_grid_double_dt_sqeuc", "cd_grid_double_sedt",
           "cd_grid_b200_flood_relabel", "cd_grid_b200_set_device", "cd_grid_b200_last_error",
           "cd_chomp_create", "cd_chomp_free", "cd_chomp_init", "cd_chomp_iterate",
           "cd_chomp_b200_set_sphere_cost"]


class CdGrid(C.Structure):
    """struct cd_grid, src/libcd/grid.h:29-41"""
    _fields_ = [("n", C.c_int), ("sizes", C.POINTER(C.c_int)), ("ncells", C.c_size_t),
                ("cell_size", C.c_int), ("data", C.c_void_p), ("lengths", C.POINTER(C.c_double))]


class Timespec(C.Structure):
    _fields_ = [("tv_sec", C.c_long), ("tv_nsec", C.c_long)]


_dp, _dpp = C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_double))


class CdChomp(C.Structure):
    """struct cd_chomp, src/libcd/chomp.h:38-101 (field for field)"""
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("lambda_", C.c_double), ("dt", C.c_double),
                ("T", _dp), ("ldt", C.c_int), ("T_points", _dpp), ("G", _dp), ("G_points", _dpp),
                ("AG", _dp), ("AG_points", _dpp), ("D", C.c_int), ("wds", _dp), ("inits", _dpp),
                ("finals", _dpp), ("initsfinals", _dp), ("A", _dp), ("Ainv", _dp), ("B", _dp),
                ("trC", C.c_double), ("jlimit_lower", _dp), ("jlimit_upper", _dp), ("Kvels", _dp),
                ("Evels", _dp), ("vels", _dp), ("cost_nxn", _dp), ("cost_mxn", _dp), ("Gjlimit", _dp),
                ("GjlimitAinv", _dp), ("cptr", C.c_void_p), ("cost_pre", C.c_void_p), ("cost", C.c_void_p),
                ("cost_extra", C.c_void_p), ("use_momentum", C.c_int), ("leapfrog_first", C.c_int),
                ("cons", C.c_void_p), ("cons_k", C.c_int), ("cons_h", _dp), ("cons_Jcol", _dp),
                ("cons_JAJT", _dp), ("cons_ipiv", C.POINTER(C.c_int)), ("cons_delta", _dp),
                ("ticks_vels", Timespec), ("ticks_callback_pre", Timespec), ("ticks_callbacks", Timespec),
                ("ticks_smoothgrad", Timespec), ("ticks_smoothcost", Timespec)]


class LibcdError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, "libcd_b200: %d (%s)" % (code, text))
        self.code = code


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBCD_PATH):
        raise OSError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIBCD_PATH)
    capi.load_library()  # the dependency, from the same directory
    lib = C.CDLL(LIBCD_PATH, mode=C.RTLD_LOCAL)
    gpp, gp = C.POINTER(C.POINTER(CdGrid)), C.POINTER(CdGrid)
    for name in ("cd_grid_double_bin_sdf", "cd_grid_double_dt_sqeuc", "cd_grid_double_sedt"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = C.c_int, [gpp, gp]
    lib.cd_grid_b200_flood_relabel.restype = C.c_int
    lib.cd_grid_b200_flood_relabel.argtypes = [gp, C.c_size_t]
    lib.cd_grid_b200_set_device.restype = C.c_int
    lib.cd_grid_b200_set_device.argtypes = [C.c_int]
    lib.cd_grid_b200_last_error.restype = C.c_char_p
    lib.cd_grid_b200_last_error.argtypes = []
    cp = C.POINTER(CdChomp)
    lib.cd_chomp_create.restype = C.c_int
    lib.cd_chomp_create.argtypes = [C.POINTER(cp), C.c_int, C.c_int, C.c_int, _dp, C.c_int]
    lib.cd_chomp_free.restype = None
    lib.cd_chomp_free.argtypes = [cp]
    lib.cd_chomp_init.restype = C.c_int
    lib.cd_chomp_init.argtypes = [cp]
    lib.cd_chomp_iterate.restype = C.c_int
    lib.cd_chomp_iterate.argtypes = [cp, C.c_int, _dp, _dp, _dp]
    lib.cd_chomp_b200_set_sphere_cost.restype = C.c_int
    lib.cd_chomp_b200_set_sphere_cost.argtypes = [cp, C.POINTER(capi.OcbRobot), C.POINTER(capi.OcbParams), C.c_int,
                                                  C.POINTER(capi.OcbSdf)]
    _lib = lib
    return lib


class HostGrid:
    """A caller-side `struct cd_grid` over numpy storage (C order, doubles)."""

    def __init__(self, array, lengths, cell_size=8):
        self.array = np.ascontiguousarray(array, dtype=np.float64)
        self.sizes = (C.c_int * self.array.ndim)(*self.array.shape)
        self.lengths = (C.c_double * self.array.ndim)(*[float(x) for x in lengths])
        self.c = CdGrid(self.array.ndim, C.cast(self.sizes, C.POINTER(C.c_int)), self.array.size, cell_size,
                        self.array.ctypes.data, C.cast(self.lengths, C.POINTER(C.c_double)))


def _take(gp):
    """Copy a grid returned by the library into numpy and release it the way cd_grid_destroy does."""
    g = gp.contents
    shape = [g.sizes[i] for i in range(g.n)]
    lengths = [g.lengths[i] for i in range(g.n)]
    assert g.cell_size == 8 and g.ncells == int(np.prod(shape))
    out = np.ctypeslib.as_array(C.cast(g.data, C.POINTER(C.c_double)), shape=(g.ncells,)).reshape(shape).copy()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(g.data)
    libc.free(C.cast(g.sizes, C.c_void_p))
    libc.free(C.cast(g.lengths, C.c_void_p))
    libc.free(C.cast(gp, C.c_void_p))
    return out, lengths


def _call(name, array, lengths):
    lib = load()
    src = HostGrid(array, lengths)
    out = C.POINTER(CdGrid)()
    rc = getattr(lib, name)(C.byref(out), C.byref(src.c))
    if rc:
        raise LibcdError(rc, lib.cd_grid_b200_last_error().decode())
    return _take(out)


def bin_sdf(obs, lengths):
    """cd_grid_double_bin_sdf on a host array -> (sdf array, lengths of the returned grid)."""
    return _call("cd_grid_double_bin_sdf", obs, lengths)


def dt_sqeuc(func, lengths, legacy_name=False):
    return _call("cd_grid_double_sedt" if legacy_name else "cd_grid_double_dt_sqeuc", func, lengths)


def flood_relabel(grid, lengths, index_start=0):
    """In place on a copy of `grid`; returns the relabelled array."""
    lib = load()
    g = HostGrid(np.array(grid, dtype=np.float64, copy=True), lengths)
    rc = lib.cd_grid_b200_flood_relabel(C.byref(g.c), index_start)
    if rc:
        raise LibcdError(rc, lib.cd_grid_b200_last_error().decode())
    return g.array


class ChompRun:
    """Drives the cd_chomp facade the way mod::create / mod::iterate drive libcd
    (mod.cpp:2521-2664, 2752-2831): the caller owns the trajectory, points inits[0] / finals[0]
    at its end rows, sets lambda / use_momentum / joint limits on the struct, then iterates."""

    def __init__(self, robot, params, sdfs, q_start, q_goal):
        self.lib = load()
        n, P = robot.n_dof, params.n_points
        self.n, self.P, self.m = n, P, P - 2
        self.traj = np.zeros((P, n))
        for i in range(P):  # mod.cpp:2456-2458
            self.traj[i] = q_start + ((np.asarray(q_goal) - q_start) * i) / (P - 1)
        self.c = C.POINTER(CdChomp)()
        rc = self.lib.cd_chomp_create(C.byref(self.c), self.m, n, params.derivative,
                                      self.traj[1:].ctypes.data_as(_dp), n)
        self._check(rc)
        c = self.c.contents
        c.dt = 1.0 / (P - 1)                                         # mod.cpp:2567
        c.inits[0] = self.traj[0:].ctypes.data_as(_dp)               # mod.cpp:2578
        c.finals[0] = self.traj[P - 1:].ctypes.data_as(_dp)          # mod.cpp:2580
        arr = (capi.OcbSdf * len(sdfs))(*[s.struct for s in sdfs])
        self._keep = (arr, sdfs, robot)
        self._check(self.lib.cd_chomp_b200_set_sphere_cost(self.c, C.byref(robot.struct), C.byref(params),
                                                           len(sdfs), arr))   # in place of mod.cpp:2616-2618
        c.lambda_ = params.lambda_                                   # mod.cpp:2625
        c.use_momentum = 1 if params.use_momentum else 0             # mod.cpp:2631-2632
        for j in range(n):                                           # mod.cpp:2639-2660
            c.jlimit_lower[j] = robot.struct.limit_lower[j]
            c.jlimit_upper[j] = robot.struct.limit_upper[j]
        self._check(self.lib.cd_chomp_init(self.c))

    def _check(self, rc):
        if rc:
            raise LibcdError(rc, self.lib.cd_grid_b200_last_error().decode())

    def iterate(self, do_iteration=1):
        t, o, s = C.c_double(), C.c_double(), C.c_double()
        rc = self.lib.cd_chomp_iterate(self.c, do_iteration, C.byref(t), C.byref(o), C.byref(s))
        return rc, np.array([t.value, o.value, s.value])

    def momentum(self):
        return np.ctypeslib.as_array(self.c.contents.AG, shape=(self.m, self.n))

    def gradient(self):
        return np.ctypeslib.as_array(self.c.contents.G, shape=(self.m, self.n))

    def close(self):
        if self.c:
            self.lib.cd_chomp_free(self.c)
            self.c = None
