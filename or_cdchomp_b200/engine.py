"""Thin Python host layer over the C ABI (include/orcdchomp_b200.h).

Every numeric operation happens inside liborcdchomp_b200.so on the GPU; this file
only marshals numpy arrays.  It mirrors the life cycle of the reference module's
run handle: create -> iterate -> gettraj -> destroy
(src/orcdchomp_mod.cpp:1800, 2690, 2854, 3013 in the reference).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import as_f64, c_double_p, c_int_p, c_uint_p, check, dptr


def _i3(a):
    return (C.c_int * 3)(*[int(x) for x in a])


def _d3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


class Engine:
    """One engine per GPU: owns the resident SDFs, a stream and scratch."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or capi.load_library()
        h = C.c_void_p()
        check(self.lib, self.lib.ocb_engine_create(int(device), C.byref(h)), "ocb_engine_create")
        self.h = h
        self.device = device

    def trim(self):
        """hand the pool's unused device memory back to the driver"""
        check(self.lib, self.lib.ocb_engine_trim(self.h), "ocb_engine_trim")

    def enable_jit(self, on=True):
        """compile the persistent kernel per batch configuration (ocb_engine_enable_jit)"""
        check(self.lib, self.lib.ocb_engine_enable_jit(self.h, 1 if on else 0), "ocb_engine_enable_jit")

    # -- SDF residency -----------------------------------------------------
    def upload_sdf(self, sdf_desc):
        sid = C.c_int()
        check(self.lib, self.lib.ocb_sdf_upload(self.h, C.byref(sdf_desc.struct), C.byref(sid)), "ocb_sdf_upload")
        return sid.value

    def adopt_sdf(self, sizes, lengths, pose_world_gsdf, device_ptr):
        sid = C.c_int()
        pose = as_f64(pose_world_gsdf)
        check(self.lib, self.lib.ocb_sdf_adopt_device(self.h, _i3(sizes), _d3(lengths), dptr(pose),
                                                     C.c_void_p(int(device_ptr)), C.byref(sid)),
              "ocb_sdf_adopt_device")
        return sid.value

    def remove_sdf(self, sid):
        check(self.lib, self.lib.ocb_sdf_remove(self.h, int(sid)), "ocb_sdf_remove")

    def sdf_sample(self, sid, points):
        """value, gradient and range flag of resident field `sid` at points of its grid frame
        (cd_grid_double_interp / grad, grid.c:331-454), evaluated by the kernels' own device function"""
        pts = as_f64(points).reshape(-1, 3)
        k = len(pts)
        vals, grads = np.zeros(k), np.zeros((k, 3))
        errs = np.zeros(k, dtype=np.int32)
        check(self.lib, self.lib.ocb_sdf_sample_host(self.h, int(sid), dptr(pts), k, dptr(vals), dptr(grads),
                                                     errs.ctypes.data_as(c_int_p)), "ocb_sdf_sample_host")
        return vals, grads, errs

    # -- SDF build ---------------------------------------------------------
    def sdf_build(self, obs, lengths):
        """cd_grid_double_bin_sdf on host arrays (copies in and out)."""
        obs = as_f64(obs)
        out = np.empty_like(obs)
        check(self.lib, self.lib.ocb_sdf_build_host(self.h, dptr(obs), _i3(obs.shape), _d3(lengths), dptr(out)),
              "ocb_sdf_build_host")
        return out

    def force_general_sdf(self, on=True):
        check(self.lib, self.lib.ocb_engine_force_general_sdf(self.h, int(bool(on))), "ocb_engine_force_general_sdf")

    def sdf_build_device(self, d_obs, sizes, lengths, d_sdf):
        check(self.lib, self.lib.ocb_sdf_build_device(self.h, C.c_void_p(int(d_obs)), _i3(sizes), _d3(lengths),
                                                     C.c_void_p(int(d_sdf))), "ocb_sdf_build_device")

    def dt_sqeuc_device(self, d_func, sizes, lengths, d_out):
        check(self.lib, self.lib.ocb_dt_sqeuc_device(self.h, C.c_void_p(int(d_func)), _i3(sizes), _d3(lengths),
                                                    C.c_void_p(int(d_out))), "ocb_dt_sqeuc_device")

    def occupancy_device(self, prims, sizes, lengths, cube_extent, d_grid):
        arr = capi.make_prims(prims)
        check(self.lib, self.lib.ocb_occupancy_device(self.h, arr, len(prims), _i3(sizes), _d3(lengths),
                                                     float(cube_extent), C.c_void_p(int(d_grid))),
              "ocb_occupancy_device")

    def flood_relabel_device(self, d_grid, sizes, index_start=0):
        check(self.lib, self.lib.ocb_flood_relabel_device(self.h, C.c_void_p(int(d_grid)), _i3(sizes),
                                                         int(index_start)), "ocb_flood_relabel_device")

    def flood_relabel(self, grid, index_start=0):
        """cd_grid_flood_fill(1.0 -> 0.0 from index_start) + relabel of the remaining 1.0 on a host
        array (returns a new array)"""
        g = np.array(grid, dtype=np.float64, order="C", copy=True)
        check(self.lib, self.lib.ocb_flood_relabel_host(self.h, dptr(g), _i3(g.shape), int(index_start)),
              "ocb_flood_relabel_host")
        return g

    def computedistancefield(self, prims, sizes, lengths, cube_extent, want_sdf=True, out=None):
        """occupancy -> flood fill + relabel -> SDF with host outputs
        (src/orcdchomp_mod.cpp:498-560 in the reference).  out: optional (obs, sdf) host arrays to
        fill (e.g. page-locked ones)."""
        arr = capi.make_prims(prims)
        shape = tuple(int(s) for s in sizes)
        if out is not None:
            obs, sdf = out
            assert obs.shape == shape and obs.dtype == np.float64 and obs.flags.c_contiguous
            assert sdf is None or (sdf.shape == shape and sdf.dtype == np.float64 and sdf.flags.c_contiguous)
            want_sdf = sdf is not None
        else:
            obs = np.empty(shape)
            sdf = np.empty(shape) if want_sdf else None
        check(self.lib, self.lib.ocb_computedistancefield_host(
            self.h, arr, len(prims), _i3(sizes), _d3(lengths), float(cube_extent), dptr(obs),
            dptr(sdf) if want_sdf else None), "ocb_computedistancefield_host")
        return obs, sdf

    def computedistancefield_resident(self, prims, sizes, lengths, cube_extent, pose_world_gsdf):
        """the same pipeline, but the field stays in HBM: returns its SDF id"""
        arr = capi.make_prims(prims)
        sid = C.c_int()
        pose = as_f64(pose_world_gsdf)
        check(self.lib, self.lib.ocb_computedistancefield_resident(
            self.h, arr, len(prims), _i3(sizes), _d3(lengths), float(cube_extent), dptr(pose), C.byref(sid)),
            "ocb_computedistancefield_resident")
        return sid.value

    def sdf_build_resident(self, obs, lengths, pose_world_gsdf):
        obs = as_f64(obs)
        sid = C.c_int()
        pose = as_f64(pose_world_gsdf)
        check(self.lib, self.lib.ocb_sdf_build_resident(self.h, dptr(obs), _i3(obs.shape), _d3(lengths), dptr(pose),
                                                       C.byref(sid)), "ocb_sdf_build_resident")
        return sid.value

    def download_sdf(self, sid, shape):
        out = np.empty(tuple(int(x) for x in shape))
        check(self.lib, self.lib.ocb_sdf_download(self.h, int(sid), dptr(out)), "ocb_sdf_download")
        return out

    def alias_sdf(self, sid, pose_world_gsdf):
        out = C.c_int()
        pose = as_f64(pose_world_gsdf)
        check(self.lib, self.lib.ocb_sdf_alias(self.h, int(sid), dptr(pose), C.byref(out)), "ocb_sdf_alias")
        return out.value

    # -- batches -----------------------------------------------------------
    def create_batch(self, robot, params, sdf_ids, q_start, q_goal, seeds=None):
        return Batch(self, robot, params, sdf_ids, q_start, q_goal, seeds)

    def set_stream(self, cuda_stream):
        check(self.lib, self.lib.ocb_engine_set_stream(self.h, C.c_void_p(int(cuda_stream or 0))),
              "ocb_engine_set_stream")

    def sync(self):
        check(self.lib, self.lib.ocb_engine_sync(self.h), "ocb_engine_sync")

    def launch_count(self):
        return int(self.lib.ocb_engine_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.ocb_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """R independent CHOMP runs on one GPU (struct run x R)."""

    def __init__(self, engine, robot, params, sdf_ids, q_start, q_goal, seeds=None):
        self.engine, self.lib = engine, engine.lib
        self.robot, self.params = robot, params
        q_start, q_goal = as_f64(q_start), as_f64(q_goal)
        if q_start.ndim == 1:
            q_start, q_goal = q_start[None, :], q_goal[None, :]
        # floating base: rows are [x y z qx qy qz qw, active dofs] (mod.cpp:2424-2443)
        assert q_start.shape == q_goal.shape and q_start.shape[1] == robot.n_dof + (7 if params.floating_base else 0)
        self.R, self.n = q_start.shape
        self.P = params.n_points
        # start_tsr: the start point is optimised too (mod.cpp:2316)
        self.m = self.P - 2 + capi.count_start_tsr(params)
        ids = np.ascontiguousarray(sdf_ids, dtype=np.int32)
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
            assert len(self._seeds) == self.R
            sp = self._seeds.ctypes.data_as(c_uint_p)
        h = C.c_void_p()
        check(self.lib, self.lib.ocb_batch_create(engine.h, C.byref(robot.struct), C.byref(params), len(ids),
                                                 ids.ctypes.data_as(c_int_p), self.R, dptr(q_start), dptr(q_goal),
                                                 sp, C.byref(h)), "ocb_batch_create")
        self.h = h

    def uses_jit(self):
        return bool(self.lib.ocb_batch_uses_jit(self.h))

    def tile_width(self):
        """waypoints per tile on the tiled large-robot path, 0 on the persistent kernel"""
        return int(self.lib.ocb_batch_tile_width(self.h))

    def reset(self, q_start=None, q_goal=None, seeds=None):
        """Re-arm the batch (straight lines, zero momentum, fresh rng); async."""
        qs = qg = sp = None
        if q_start is not None:
            self._qs, self._qg = as_f64(q_start), as_f64(q_goal)
            assert self._qs.shape == (self.R, self.n) and self._qg.shape == (self.R, self.n)
            qs, qg = dptr(self._qs), dptr(self._qg)
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
            sp = self._seeds.ctypes.data_as(c_uint_p)
        check(self.lib, self.lib.ocb_batch_reset(self.h, qs, qg, sp), "ocb_batch_reset")

    def set_traj(self, traj):
        traj = as_f64(traj)
        assert traj.shape == (self.R, self.P, self.n)
        check(self.lib, self.lib.ocb_batch_set_traj(self.h, dptr(traj)), "ocb_batch_set_traj")

    def enable_trace(self, on=True):
        check(self.lib, self.lib.ocb_batch_enable_trace(self.h, int(bool(on))), "ocb_batch_enable_trace")

    def capture_gradient(self, mode):
        check(self.lib, self.lib.ocb_batch_capture_gradient(self.h, int(mode)), "ocb_batch_capture_gradient")

    def iterate(self, n_iter, first_iter=0):
        """Returns (costs[R,3] = total/obs/smooth of the final cost-only pass, status[R]).
        first_iter: number of this call's first iteration within one `iterate` command that is
        split over several calls (HMC schedule, ocb_batch_iterate_from)."""
        ct, co, cs = np.empty(self.R), np.empty(self.R), np.empty(self.R)
        st = np.empty(self.R, dtype=np.int32)
        check(self.lib, self.lib.ocb_batch_iterate_from(self.h, int(first_iter), int(n_iter), dptr(ct), dptr(co),
                                                       dptr(cs), st.ctypes.data_as(c_int_p)), "ocb_batch_iterate_from")
        return np.stack([ct, co, cs], axis=1), st

    def iterate_async(self, n_iter):
        check(self.lib, self.lib.ocb_batch_iterate_async(self.h, int(n_iter)), "ocb_batch_iterate_async")

    def get_costs(self):
        ct, co, cs = np.empty(self.R), np.empty(self.R), np.empty(self.R)
        st = np.empty(self.R, dtype=np.int32)
        check(self.lib, self.lib.ocb_batch_get_costs(self.h, dptr(ct), dptr(co), dptr(cs),
                                                    st.ctypes.data_as(c_int_p)), "ocb_batch_get_costs")
        return np.stack([ct, co, cs], axis=1), st

    def get_trace(self, n_iter):
        out = np.empty((self.R, n_iter, 3))
        check(self.lib, self.lib.ocb_batch_get_trace(self.h, dptr(out), int(n_iter)), "ocb_batch_get_trace")
        return out

    def get_iterations(self):
        """iterations each run completed in the last iterate call"""
        out = np.zeros(self.R, dtype=np.int32)
        check(self.lib, self.lib.ocb_batch_get_iterations(self.h, out.ctypes.data_as(c_int_p)), "ocb_batch_get_iterations")
        return out

    def get_limit_rounds(self):
        """most joint-limit projection steps one iteration of the last iterate call needed, per run"""
        out = np.zeros(self.R, dtype=np.int32)
        check(self.lib, self.lib.ocb_batch_get_limit_rounds(self.h, out.ctypes.data_as(c_int_p)), "ocb_batch_get_limit_rounds")
        return out

    def get_constraint_skips(self):
        """per run: constraint systems that were singular (dependent rows) and skipped"""
        out = np.zeros(self.R, dtype=np.int32)
        check(self.lib, self.lib.ocb_batch_get_constraint_skips(self.h, out.ctypes.data_as(c_int_p)),
              "ocb_batch_get_constraint_skips")
        return out

    def get_traj(self, out=None):
        if out is None:
            out = np.empty((self.R, self.P, self.n))
        check(self.lib, self.lib.ocb_batch_get_traj(self.h, dptr(out)), "ocb_batch_get_traj")
        return out

    def get_gradient(self):
        out = np.empty((self.R, self.m, self.n))
        check(self.lib, self.lib.ocb_batch_get_gradient(self.h, dptr(out)), "ocb_batch_get_gradient")
        return out

    def best(self):
        idx, cost = C.c_int(), C.c_double()
        check(self.lib, self.lib.ocb_batch_best(self.h, C.byref(idx), C.byref(cost)), "ocb_batch_best")
        return idx.value, cost.value

    def copy_run_traj_device(self, run, d_dst):
        check(self.lib, self.lib.ocb_batch_copy_run_traj_device(self.h, int(run), C.c_void_p(int(d_dst))),
              "ocb_batch_copy_run_traj_device")

    def device_ptrs(self):
        t, c = C.c_void_p(), C.c_void_p()
        check(self.lib, self.lib.ocb_batch_device_ptrs(self.h, C.byref(t), C.byref(c)), "ocb_batch_device_ptrs")
        return t.value, c.value

    def close(self):
        if self.h:
            self.lib.ocb_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiEngine:
    """G engines in one process, one host thread each, behind the C ABI (csrc/ocb_multi.cpp):
    runs dealt round-robin, fields replicated, best-cost gather over NCCL / peer copies."""

    def __init__(self, devices, lib=None):
        self.lib = lib or capi.load_library()
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        self._check(self.lib.ocb_multi_create(len(devs), devs.ctypes.data_as(c_int_p), C.byref(h)), "ocb_multi_create")
        self.h = h
        self.G = len(devs)

    def _check(self, code, what):
        if code != capi.OCB_OK:
            raise capi.OcbError("%s failed with code %d: %s" % (what, code, (self.lib.ocb_multi_last_error() or b"").decode()))

    def uses_nccl(self):
        return bool(self.lib.ocb_multi_uses_nccl(self.h))

    def enable_jit(self, on=True):
        self._check(self.lib.ocb_multi_enable_jit(self.h, int(bool(on))), "ocb_multi_enable_jit")

    def upload_sdf(self, sdf_desc):
        sid = C.c_int()
        self._check(self.lib.ocb_multi_sdf_upload(self.h, C.byref(sdf_desc.struct), C.byref(sid)), "ocb_multi_sdf_upload")
        return sid.value

    def remove_sdf(self, sid):
        self._check(self.lib.ocb_multi_sdf_remove(self.h, int(sid)), "ocb_multi_sdf_remove")

    def create_batch(self, robot, params, sdf_ids, q_start, q_goal, seeds=None):
        return MultiBatch(self, robot, params, sdf_ids, q_start, q_goal, seeds)

    def close(self):
        if self.h:
            self.lib.ocb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBatch:
    def __init__(self, multi, robot, params, sdf_ids, q_start, q_goal, seeds=None):
        self.multi, self.lib = multi, multi.lib
        q_start, q_goal = as_f64(q_start), as_f64(q_goal)
        self.R, self.n = q_start.shape
        self.P = params.n_points
        ids = np.ascontiguousarray(sdf_ids, dtype=np.int32)
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
            sp = self._seeds.ctypes.data_as(c_uint_p)
        h = C.c_void_p()
        multi._check(self.lib.ocb_multi_batch_create(multi.h, C.byref(robot.struct), C.byref(params), len(ids),
                                                     ids.ctypes.data_as(c_int_p), self.R, dptr(q_start), dptr(q_goal),
                                                     sp, C.byref(h)), "ocb_multi_batch_create")
        self.h = h

    def iterate(self, n_iter):
        ct, co, cs = np.empty(self.R), np.empty(self.R), np.empty(self.R)
        st = np.empty(self.R, dtype=np.int32)
        self.multi._check(self.lib.ocb_multi_batch_iterate(self.h, int(n_iter), dptr(ct), dptr(co), dptr(cs),
                                                           st.ctypes.data_as(c_int_p)), "ocb_multi_batch_iterate")
        return np.stack([ct, co, cs], axis=1), st

    def get_traj(self):
        out = np.empty((self.R, self.P, self.n))
        self.multi._check(self.lib.ocb_multi_batch_get_traj(self.h, dptr(out)), "ocb_multi_batch_get_traj")
        return out

    def best(self):
        idx, cost = C.c_int(), C.c_double()
        traj = np.zeros((self.P, self.n))
        self.multi._check(self.lib.ocb_multi_batch_best(self.h, C.byref(idx), C.byref(cost), dptr(traj)), "ocb_multi_batch_best")
        return idx.value, cost.value, traj

    def close(self):
        if self.h and self.multi.h:
            self.lib.ocb_multi_batch_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
