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
           "cd_grid_b200_flood_relabel", "cd_grid_b200_set_device", "cd_grid_b200_last_error"]


class CdGrid(C.Structure):
    """struct cd_grid, src/libcd/grid.h:29-41"""
    _fields_ = [("n", C.c_int), ("sizes", C.POINTER(C.c_int)), ("ncells", C.c_size_t),
                ("cell_size", C.c_int), ("data", C.c_void_p), ("lengths", C.POINTER(C.c_double))]


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
