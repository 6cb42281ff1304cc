#!/usr/bin/env python
"""bench.py -- batched CHOMP run-iterations/s (WAM7), BASELINE.json configs[1], plus every other
BASELINE configuration as a sub-record of the same JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--scaling weak|strong] [--no-extras] [--no-cpu-baseline] [--no-jit]
  (N > 1: launched by torchrun, one rank per GPU)

Headline (the line's own keys).  One "step" = the hot path over one batch: every GPU re-arms its
shard of independent WAM7 runs (random start/goal, n_points=100, table SDF) with straight-line
trajectories and runs 100 CHOMP iterations + the final cost pass (`iterate run ... n_iter 100`).
  --scaling weak    4096 runs PER GPU (default)
  --scaling strong  4096 runs IN TOTAL, dealt round-robin over the GPUs (BASELINE configs[1]:
                    "4096 runs sharded at 1/2/4/8"); at N > 1 the weak line also carries this
                    measurement as the sub-record "strong_scaling"

  value         run-iterations/s with the end points already in HBM (CUDA events on the launching
                stream, max over ranks, summed over the K steps)
  e2e           the same metric through the public C ABI with HOST buffers: ocb_batch_create (H2D
                end points) -> ocb_batch_iterate -> ocb_batch_get_traj (D2H) -> ocb_batch_destroy
  roofline      algorithmic bytes (58 128 B per run-iteration) / CUDA-event time
                of the iterate kernel, against MEASURED_PEAKS.json hbm_gbs
  roofline_fp64 the second ceiling: fp64 FLOP per run-iteration counted by ncu
                (profiles/r2_fp64_ops.json) x achieved rate, against the measured fp64 peak
  cpu_baseline  the reference CPU path (oracle/_ref: the reference's own libcd + restated
                callbacks) on ONE host core, bounded sample, rank 0 only

Sub-records under "configs" (rank 0, single GPU), each with roofline, cpu_baseline and e2e:
  cfg3_sdf_build   computedistancefield on the cluttered kinbody, 400^3 (Mvoxels/s)
  cfg4_hmc         8192 seeds, use_momentum + use_hmc, n_points=256
  cfg5_dense       200 spheres, n_points=1024, 4 rotated 128^3 fields (tiled path)
  cfg2_hbm_field   the cfg2 batch against the 400^3 (512 MB, HBM-resident) field of cfg3

--impl reference times the reference CPU path on all host cores (one run per thread).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RUNS_PER_GPU = 4096
N_POINTS = 100
N_ITER = 100
LAMBDA = 100.0
OBS_FACTOR = 500.0
METRIC = "chomp_run_iterations_per_s"
UNIT = "run-iterations/s"
FP64_PEAK_TFLOPS = 37.1   # measured on this part: profiles/r1_microbench.txt (148 x 8 blocks x 256 threads, ILP 8)


def algorithmic_bytes_per_run_iter(P, n, n_active, n_sdf, momentum):
    """8*[P*n + m*n + 2*m*n*momentum + 4*m*S_a*K]."""
    m = P - 2
    return 8 * (P * n + m * n + (2 * m * n if momentum else 0) + 4 * m * n_active * n_sdf)


def load_json(*path):
    try:
        with open(os.path.join(ROOT, *path)) as f:
            return json.load(f)
    except Exception:
        return {}


def hbm_peak():
    peaks = load_json("MEASURED_PEAKS.json")
    return float(peaks.get("hbm_gbs", 6650.0)), ("MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else
                                                 "fallback 6650 (of fallback)")


def build_scene():
    from or_cdchomp_b200 import capi, models
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gprims = models.prims_to_grid_frame(prims, gpose)
    pose_world = models.pose_compose(kin_pose, gpose)
    params = capi.default_params(n_points=N_POINTS, lambda_=LAMBDA, obs_factor=OBS_FACTOR)
    return robot, params, gprims, sizes, lengths, pose_world


def config_dict(n_gpus, scaling="weak", runs_per_gpu=RUNS_PER_GPU):
    if scaling == "weak":
        what = "WAM7 batch of %d independent runs per GPU" % runs_per_gpu
    else:
        what = "WAM7 batch of %d independent runs in total, dealt round-robin over the GPUs" % RUNS_PER_GPU
    return {
        "workload": "BASELINE configs[1]: %s (random start/goal, n_points=100, 100 iterations, one table SDF "
                    "31x40x11)" % what,
        "runs_per_gpu": runs_per_gpu, "n_points": N_POINTS, "n_iter_per_step": N_ITER,
        "n_dof": 7, "spheres_active": 15, "n_sdfs": 1, "lambda": LAMBDA, "obs_factor": OBS_FACTOR,
        "parallelism": "runs sharded, %d GPU(s), SDF replicated, best-cost gather only" % n_gpus,
        "l2": "flushed between timed steps (256 MiB write)",
    }


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


class _StdoutToStderr:
    """libcd printf()s diagnostics ("ran too many joint limit fixes!", chomp.c:653) on the C
    stdout; keep them off this script's stdout, which must carry exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


CPU_SAMPLE_RUNS = 384  # ~12 s of single-core reference work at ~3.2 k run-iterations/s


class _PinOneCore:
    """one core, pinned (`taskset -c`), affinity restored afterwards"""

    def __enter__(self):
        self.saved = None
        try:
            self.saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, {sorted(self.saved)[0]})
        except (AttributeError, OSError):
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def cpu_baseline(flavour, robot, params, sds, starts, goals, n_iter, threads=1, seeds=None):
    """run-iterations/s of the CPU oracle over the given runs on `threads` host threads (one run each at a
    time).  Returns (rate, seconds)."""
    with _StdoutToStderr():
        return _cpu_baseline(flavour, robot, params, sds, starts, goals, n_iter, threads, seeds)


def _cpu_baseline(flavour, robot, params, sds, starts, goals, n_iter, threads, seeds):
    from oracle import pyoracle as po
    po.load(flavour)
    n_runs = len(starts)
    done = [0] * n_runs

    def work(r):
        run = po.Run(robot, params, sds, starts[r], goals[r], seed=0 if seeds is None else int(seeds[r]),
                     flavour=flavour)
        ret, _, tr, _ = run.iterate(n_iter, want_trace=True)
        if ret == 0:
            t_iter = n_iter
        else:  # the reference aborts the run when it leaves the joint limits (mod.cpp:2799-2803)
            t_iter = int(np.count_nonzero(tr[:, 0])) or 1
        run.close()
        done[r] = t_iter

    t0 = time.perf_counter()
    if threads <= 1:
        with _PinOneCore():
            for r in range(n_runs):
                work(r)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(n_runs)))
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from or_cdchomp_b200 import capi, models
    from oracle import pyoracle as po
    robot, params, gprims, sizes, lengths, pose_world = build_scene()
    flavour = po.best_flavour()
    pa = capi.make_prims(gprims)
    _, sdf = po.computedistancefield(pa, len(gprims), sizes, lengths, 0.02, flavour=flavour)
    sd = capi.SdfDesc(sdf, lengths, pose_world)
    cores = os.cpu_count() or 1
    runs_per_step = cores * 32
    starts, goals = models.random_endpoints(robot, runs_per_step)
    for _ in range(args.warmup):
        cpu_baseline(flavour, robot, params, [sd], starts[:min(cores, 4)], goals[:min(cores, 4)], 10, threads=cores)
    t_total, it_total = 0.0, 0.0
    for _ in range(args.steps):
        v, dt = cpu_baseline(flavour, robot, params, [sd], starts, goals, N_ITER, threads=cores)
        t_total += dt
        it_total += v * dt
    value = it_total / t_total
    sample = "%d runs x %d iterations per step on %d threads (one run per thread)" % (runs_per_step, N_ITER, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": "reference" if flavour == "reference" else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# one batch configuration, timed: used by the headline and by the cfg4 / cfg5 / HBM-field sub-records


class BatchBench:
    def __init__(self, torch, eng, stream, device, robot, params, sids, starts, goals, seeds=None):
        self.torch, self.eng, self.stream, self.device = torch, eng, stream, device
        self.robot, self.params, self.sids = robot, params, sids
        self.starts, self.goals, self.seeds = starts, goals, seeds
        self.batch = eng.create_batch(robot, params, sids, starts, goals, seeds=seeds)
        self.R = len(starts)

    def step(self, n_iter, events=None):
        if events is not None:
            events[0].record(self.stream)
        self.batch.reset()
        if events is not None:
            events[1].record(self.stream)
        self.batch.iterate_async(n_iter)
        if events is not None:
            events[2].record(self.stream)

    def timed(self, n_iter, steps, warmup, flush=None):
        """(kernel ms per step, run-iterations done per step, runs failed)"""
        torch = self.torch
        for _ in range(warmup):
            self.step(n_iter)
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for k in range(steps):
            if flush is not None:
                flush.fill_(k & 0xFF)
            self.step(n_iter, evs[k])
        torch.cuda.synchronize()
        kern_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
        costs, status = self.batch.get_costs()
        done = int(self.batch.get_iterations().sum())
        return kern_ms, done, int((status != 0).sum())

    def e2e(self, n_iter, steps, out_traj, starts_h, goals_h):
        """create + iterate + gettraj + destroy through the C ABI with page-locked host buffers, wall clock"""
        def one():
            b2 = self.eng.create_batch(self.robot, self.params, self.sids, starts_h, goals_h, seeds=self.seeds)
            b2.iterate(n_iter)
            b2.get_traj(out_traj)
            b2.close()
        one()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        self.torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    def close(self):
        self.batch.close()


def pinned(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().numpy()


def roofline_record(abytes, run_iters, kern_ms, traffic=None, kernel="chomp_iterate_kernel", note=None):
    peak, src = hbm_peak()
    achieved = abytes * run_iters / (kern_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": traffic, "kernel": kernel, "algorithmic_bytes_per_run_iteration": abytes, "peak_source": src}
    if note:
        r["note"] = note
    return r


def measure_sdf_build(torch, eng, stream, device, with_cpu=True, n=400):
    """BASELINE configs[2]: computedistancefield on the cluttered kinbody, cube_extent 0.005 -> 400^3.
    value = cd_grid_double_bin_sdf equivalent, occupancy grid in HBM -> SDF in HBM (SURVEY.md section 8d:
    16 algorithmic bytes per voxel); voxelisation and flood fill reported separately; e2e = the whole
    ocb_computedistancefield_host call (primitives in, occupancy + SDF out to page-locked host buffers);
    cpu_baseline = the reference's cd_grid_double_bin_sdf on the same occupancy grid, one core."""
    from or_cdchomp_b200 import capi, models
    prims, apos, aext = models.clutter_scene()
    ce = 0.005 * 400 / n
    sizes, lengths, gpose = models.field_geometry(apos, aext, ce, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    ncell = int(np.prod(sizes))
    d_obs = torch.empty(ncell, dtype=torch.float64, device=device)
    d_sdf = torch.empty(ncell, dtype=torch.float64, device=device)

    def timed(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    d_occ = torch.empty(ncell, dtype=torch.float64, device=device)
    occ_ms = timed(lambda: eng.occupancy_device(gp, sizes, lengths, ce, d_occ.data_ptr()))

    def flood():
        d_obs.copy_(d_occ)
        eng.flood_relabel_device(d_obs.data_ptr(), sizes, 0)
    copy_ms = timed(lambda: d_obs.copy_(d_occ))
    flood_ms = max(timed(flood) - copy_ms, 0.0)
    del d_occ
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.sdf_build_device(d_obs.data_ptr(), sizes, lengths, d_sdf.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times[1:]))
    peak, src = hbm_peak()
    gbs = 16.0 * ncell / (ms * 1e-3) / 1e9
    obs_host = d_obs.cpu().numpy().reshape(sizes) if with_cpu else None
    del d_obs, d_sdf
    # end to end with host buffers
    shape = tuple(int(s) for s in sizes)
    out_obs = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    out_sdf = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    eng.computedistancefield(gp, sizes, lengths, ce, out=(out_obs, out_sdf))
    e2e_s = 1e30
    for _ in range(3):  # best of three: the 1 GB device-to-host copy shares the host's PCIe / memory with other tenants
        t0 = time.perf_counter()
        eng.computedistancefield(gp, sizes, lengths, ce, out=(out_obs, out_sdf))
        e2e_s = min(e2e_s, time.perf_counter() - t0)
    traffic = (load_json("profiles", "traffic.json").get("sdf_build_bytes_400cubed_r2") or {}).get("total") if n == 400 else None
    rec = {"metric": "sdf_build_mvoxels_per_s", "value": ncell / (ms * 1e-3) / 1e6, "unit": "Mvoxels/s",
           "workload": "BASELINE configs[2]: computedistancefield, 64 boxes + 32 spheres, cube_extent %g -> %s voxels"
                       % (ce, "x".join(str(int(s)) for s in sizes)),
           "sizes": [int(x) for x in sizes], "ms": ms,
           "occupancy_ms": occ_ms, "flood_relabel_ms": flood_ms, "n_primitives": len(gp),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "traffic": traffic, "algorithmic_bytes_per_voxel": 16, "peak_source": src,
                        "kernel": "pack_rows + edt_zy + edt_x (the whole build)"},
           "e2e": {"value": ncell / e2e_s / 1e6, "unit": "Mvoxels/s", "h2d_bytes_per_step": len(gp) * 88,
                   "d2h_bytes_per_step": 2 * ncell * 8, "ms": e2e_s * 1e3,
                   "what": "ocb_computedistancefield_host: occupancy + flood fill + SDF, both grids copied to "
                           "page-locked host buffers, wall clock"}}
    if with_cpu:
        from oracle import pyoracle as po
        flavour = po.best_flavour()
        with _PinOneCore():
            t0 = time.perf_counter()
            ref = po.sdf_from_obsarray(obs_host, lengths, flavour=flavour)
            dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": ncell / dt / 1e6, "unit": "Mvoxels/s", "cores": 1,
                               "kind": "reference" if flavour == "reference" else "port",
                               "sample": "cd_grid_double_bin_sdf on the same %s occupancy grid, one thread (%.1f s)"
                                         % ("x".join(str(int(s)) for s in sizes), dt)}
        rec["max_abs_diff_vs_cpu"] = float(np.max(np.abs(ref - out_sdf)))
    return rec


def measure_cfg4(torch, eng, stream, device, sid, sd, with_cpu, steps=2):
    """BASELINE configs[3]: one start/goal, 8192 seeds, use_momentum + use_hmc, n_points=256."""
    from or_cdchomp_b200 import capi, models
    robot = models.wam7_robot()
    R, P = 8192, 256
    params = capi.default_params(n_points=P, lambda_=LAMBDA, obs_factor=OBS_FACTOR, use_momentum=1, use_hmc=1,
                                 hmc_resample_lambda=0.02)
    starts, goals = models.random_endpoints(robot, 1, shrink=0.3)
    st, go = np.repeat(starts, R, 0), np.repeat(goals, R, 0)
    seeds = np.arange(1, R + 1, dtype=np.uint32)
    bb = BatchBench(torch, eng, stream, device, robot, params, [sid], st, go, seeds)
    kern_ms, done, failed = bb.timed(N_ITER, steps, 1)
    best = bb.batch.best()
    out_traj = torch.empty((R, P, 7), dtype=torch.float64, pin_memory=True).numpy()
    e2e_s = bb.e2e(N_ITER, 1, out_traj, pinned(torch, st), pinned(torch, go))
    kernel = "run-time specialised (NVRTC)" if bb.batch.uses_jit() else "library instantiation"
    bb.close()
    abytes = algorithmic_bytes_per_run_iter(P, 7, robot.n_spheres_active, 1, True)
    rec = {"metric": METRIC, "value": done / (kern_ms * 1e-3), "unit": UNIT,
           "workload": "BASELINE configs[3]: WAM7 use_hmc + use_momentum, 8192 seeds, n_points=256, 100 iterations, "
                       "best-cost arg-min",
           "runs": R, "kernel_ms_per_step": kern_ms, "runs_failed_joint_limits": failed, "kernel": kernel,
           "best_run": best[0], "best_cost": best[1],
           "roofline": roofline_record(abytes, done, kern_ms, kernel="chomp_iterate_kernel<256>"),
           "e2e": {"value": done / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * R * 7 * 8 + R * 4 + R * 625 * 4,
                   "d2h_bytes_per_step": R * P * 7 * 8 + R * 28,
                   "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"}}
    if with_cpu:
        from oracle import pyoracle as po
        flavour = po.best_flavour()
        k = 24
        v, dt = cpu_baseline(flavour, robot, params, [sd], st[:k], go[:k], N_ITER, threads=1, seeds=seeds[:k])
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if flavour == "reference" else "port",
                               "sample": "seeds 1..%d x %d iterations, one thread (%.1f s)" % (k, N_ITER, dt)}
    return rec


def dense_fields(n=128, k=4, seed=9):
    from or_cdchomp_b200 import capi, models
    rng = np.random.default_rng(seed)
    x = (np.arange(n) + 0.5) / n
    sds = []
    for _ in range(k):
        f = (0.15 + 0.5 * np.abs(x[:, None, None] - rng.uniform(0.3, 0.7)) + 0.4 * np.abs(x[None, :, None] - 0.5)
             + 0.3 * np.abs(x[None, None, :] - rng.uniform(0.3, 0.7))
             + 0.02 * np.sin(9.0 * x[:, None, None]) * np.cos(7.0 * x[None, :, None] + 5.0 * x[None, None, :]))
        pose = models.pose_make(rng.uniform(-1.2, -0.6, size=3),
                                models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 1.0)))
        sds.append(capi.SdfDesc(f, [2.0, 2.0, 2.0], pose))
    return sds


def measure_cfg5(torch, eng, stream, device, with_cpu, steps=2):
    """BASELINE configs[4]: 7-dof arm with 200 spheres, n_points=1024, 4 rotated 128^3 fields, R = 256."""
    from or_cdchomp_b200 import capi, models
    robot = models.dense_sphere_arm(200, seed=5)
    sds = dense_fields()
    ids = [eng.upload_sdf(s) for s in sds]
    R, P, n_iter = 256, 1024, 10
    params = capi.default_params(n_points=P, lambda_=200.0, obs_factor=100.0, epsilon=0.2)
    starts, goals = models.random_endpoints(robot, R, seed0=77, shrink=0.4)
    bb = BatchBench(torch, eng, stream, device, robot, params, ids, starts, goals)
    launches0 = eng.launch_count()
    kern_ms, done, failed = bb.timed(n_iter, steps, 1)
    launches = (eng.launch_count() - launches0) // (steps + 1)
    tile_w = bb.batch.tile_width()
    out_traj = torch.empty((R, P, 7), dtype=torch.float64, pin_memory=True).numpy()
    e2e_s = bb.e2e(n_iter, 1, out_traj, pinned(torch, starts), pinned(torch, goals))
    bb.close()
    abytes = algorithmic_bytes_per_run_iter(P, 7, robot.n_spheres_active, 4, False)
    rec = {"metric": METRIC, "value": done / (kern_ms * 1e-3), "unit": UNIT,
           "workload": "BASELINE configs[4]: 7-dof arm, 200 spheres, n_points=1024, 4 rotated 128^3 SDFs, %d runs x %d "
                       "iterations" % (R, n_iter),
           "runs": R, "n_iter_per_step": n_iter, "kernel_ms_per_step": kern_ms, "runs_failed_joint_limits": failed,
           "tile_width": tile_w, "launches_per_step": launches,
           "roofline": roofline_record(abytes, done, kern_ms,
                                       traffic=load_json("profiles", "traffic.json").get(
                                           "chomp_tile_cost_kernel_bytes_per_launch_16_runs_cfg5"),
                                       kernel="chomp_tile_cost_kernel + chomp_run_update_kernel (%d launches per step)" % launches),
           "e2e": {"value": done / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * R * 7 * 8,
                   "d2h_bytes_per_step": R * P * 7 * 8 + R * 28,
                   "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"}}
    if with_cpu:
        from oracle import pyoracle as po
        flavour = po.best_flavour()
        v, dt = cpu_baseline(flavour, robot, params, sds, starts[:1], goals[:1], 3, threads=1)
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if flavour == "reference" else "port",
                               "sample": "run 0 x 3 iterations incl. cd_chomp_init (1022 x 1022 LU), one thread (%.1f s)" % dt}
    for i in ids:
        eng.remove_sdf(i)
    return rec


def measure_hbm_field(torch, eng, stream, device, with_cpu, steps=3):
    """The cfg2 batch against an HBM-resident field: the 400^3 (512 MB) field of BASELINE configs[2],
    built on this GPU and kept resident; the robot stands in the middle of the cluttered kinbody.
    Here the byte roofline is real: every lookup reads 4 cells of a field 4x the size of L2."""
    from or_cdchomp_b200 import capi, models
    robot = models.wam7_robot()
    prims, apos, aext = models.clutter_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.005, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    sid = eng.computedistancefield_resident(gp, sizes, lengths, 0.005, gpose)
    # lambda 1000 instead of 100: inside the clutter the obstacle gradients are ~10x the table's and steps of
    # the config-2 size throw most runs out of their joint limits (38 of 64 in a CPU probe; 0 of 64 here)
    params = capi.default_params(n_points=N_POINTS, lambda_=10.0 * LAMBDA, obs_factor=OBS_FACTOR)
    R = RUNS_PER_GPU
    starts, goals = models.random_endpoints(robot, R)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    bb = BatchBench(torch, eng, stream, device, robot, params, [sid], starts, goals)
    kern_ms, done, failed = bb.timed(N_ITER, steps, 2, flush)
    out_traj = torch.empty((R, N_POINTS, 7), dtype=torch.float64, pin_memory=True).numpy()
    e2e_s = bb.e2e(N_ITER, 2, out_traj, pinned(torch, starts), pinned(torch, goals))
    kernel = "run-time specialised (NVRTC)" if bb.batch.uses_jit() else "library instantiation"
    bb.close()
    abytes = algorithmic_bytes_per_run_iter(N_POINTS, 7, robot.n_spheres_active, 1, False)
    tr = load_json("profiles", "traffic.json")
    rec = {"metric": METRIC, "value": done / (kern_ms * 1e-3), "unit": UNIT,
           "workload": "cfg2 batch (4096 WAM7 runs x 100 iterations, lambda 1000) against the HBM-resident 400^3 / 512 MB "
                       "field of configs[2]",
           "runs": R, "kernel_ms_per_step": kern_ms, "runs_failed_joint_limits": failed, "kernel": kernel,
           "field_bytes": int(np.prod(sizes)) * 8,
           "roofline": roofline_record(abytes, done, kern_ms, traffic=tr.get("chomp_hbm_field_bytes_per_launch"),
                                       note="sector-granular variant: 96 B per lookup instead of 32 B -> "
                                            "%d B per run-iteration" % (abytes + 8 * 8 * 98 * 15)),
           "e2e": {"value": done / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * R * 7 * 8,
                   "d2h_bytes_per_step": R * N_POINTS * 7 * 8 + R * 28,
                   "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"}}
    if with_cpu:
        from oracle import pyoracle as po
        flavour = po.best_flavour()
        sdf_host = eng.download_sdf(sid, sizes)
        sd = capi.SdfDesc(sdf_host, lengths, gpose)
        k = 40
        v, dt = cpu_baseline(flavour, robot, params, [sd], starts[:k], goals[:k], N_ITER, threads=1)
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if flavour == "reference" else "port",
                               "sample": "first %d runs x %d iterations on the same 512 MB field, one thread (%.1f s)" % (k, N_ITER, dt)}
    eng.remove_sdf(sid)
    del flush
    return rec


def measure_tsr(torch, eng, stream, device, sid, sd, with_cpu, steps=3):
    """the cfg2 scene with a hard task-space-region constraint on EVERY waypoint (everyn_tsr:
    tool height, roll and pitch held -- 'carry the cup upright'), 4096 runs x 100 iterations.  Starts share one arm
    posture, goals differ mostly in the first joint so that every run can satisfy its constraint."""
    from or_cdchomp_b200 import capi, models
    from oracle import pyoracle as po
    robot = models.wam7_robot()
    ee = robot.names.index("wam7")
    R = RUNS_PER_GPU
    rng = np.random.default_rng(20260217)
    base = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    starts = np.repeat(base[None], R, 0)
    goals = starts.copy()
    goals[:, 0] += rng.uniform(0.6, 1.3, R)
    goals[:, 1:] += rng.uniform(-0.05, 0.05, (R, 6))
    flavour = po.best_flavour()
    pe = po.fk(robot, base, flavour=flavour)[ee]  # set-up only: where the tool is at the start
    Bw = np.tile(np.array([-10.0, 10.0]), (6, 1))
    Bw[[2, 3, 4]] = 0.0
    cons = [capi.make_constraint("all", ee, Bw, T0w=models.pose_make((0, 0, pe[2])), Twe=models.pose_make((0, 0, 0), pe[3:7]))]
    params = capi.default_params(n_points=N_POINTS, lambda_=LAMBDA, obs_factor=OBS_FACTOR, constraints=cons)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    bb = BatchBench(torch, eng, stream, device, robot, params, [sid], starts, goals)
    kern_ms, done, failed = bb.timed(N_ITER, steps, 2, flush)
    skips = int(bb.batch.get_constraint_skips().sum())
    out_traj = torch.empty((R, N_POINTS, 7), dtype=torch.float64, pin_memory=True).numpy()
    e2e_s = bb.e2e(N_ITER, 2, out_traj, pinned(torch, starts), pinned(torch, goals))
    bb.close()
    del flush
    rec = {"metric": METRIC, "value": done / (kern_ms * 1e-3), "unit": UNIT,
           "workload": "cfg2 scene, 4096 WAM7 runs x 100 iterations, every waypoint under a 3-row TSR constraint "
                       "(294 constraint rows per run; block-tridiagonal projection, library kernel)",
           "runs": R, "kernel_ms_per_step": kern_ms, "runs_failed_joint_limits": failed, "constraint_rows": 3 * (N_POINTS - 2),
           "singular_waypoint_systems": skips,
           "roofline": roofline_record(algorithmic_bytes_per_run_iter(N_POINTS, 7, robot.n_spheres_active, 1, False), done, kern_ms,
                                       traffic=load_json("profiles", "traffic.json").get("chomp_tsr_bytes_per_launch"),
                                       kernel="chomp_iterate_kernel<128,0,0,0,1> (library kernel with the constraint projection)",
                                       note="same compulsory bytes as the unconstrained iteration (the constraint's operands never leave "
                                            "the SM); the projection sweep is latency-bound: profiles/r2_tsr_ncu.csv"),
           "e2e": {"value": done / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * R * 7 * 8,
                   "d2h_bytes_per_step": R * N_POINTS * 7 * 8 + R * 28,
                   "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"}}
    if with_cpu:
        k = 8
        v, dt = cpu_baseline(flavour, robot, params, [sd], starts[:k], goals[:k], N_ITER, threads=1)
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if flavour == "reference" else "port",
                               "sample": "first %d runs x %d iterations, one thread (%.1f s); the reference solves a dense "
                                         "294 x 294 system per iteration (LAPACKE_dgesv, chomp.c:579)" % (k, N_ITER, dt)}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--runs", type=int, default=RUNS_PER_GPU, help="runs per GPU (weak) or in total (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", "--no-sdf", dest="no_extras", action="store_true",
                    help="headline only: skip the cfg3 / cfg4 / cfg5 / HBM-field / TSR sub-records")
    ap.add_argument("--only", default="", help="comma list of sub-records to run (cfg3,cfg4,cfg5,hbm,tsr); default all")
    ap.add_argument("--no-jit", action="store_true", help="use the library's own kernel instead of the run-time specialised one")
    ap.add_argument("--shrink", type=float, default=0.05,
                    help="end points drawn in the joint limits shrunk by this fraction (BASELINE: 0.05; experiments only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from or_cdchomp_b200 import capi, models, sharding
    from or_cdchomp_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    robot, params, gprims, sizes, lengths, pose_world = build_scene()
    P, n = params.n_points, robot.n_dof

    eng = Engine(local_rank)
    if not args.no_jit:
        # the engine's run-time specialisation (ocb_engine_enable_jit): the kernel is compiled once for
        # this batch shape during set-up, outside every timed region, and cached
        eng.enable_jit(True)
    # a real (non-default) stream shared by torch and the engine, so torch.cuda.Event sees the kernels
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    # SDF built on this GPU (replicated per GPU, never communicated)
    obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
    sd = capi.SdfDesc(sdf, lengths, pose_world)
    sid = eng.upload_sdf(sd)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_headline(scaling, steps, warmup):
        """runs are dealt round-robin: global run id = rank + world * i (the run's seed is its global id, so
        results do not depend on N; joint-limit-heavy runs spread evenly over the ranks)"""
        total = args.runs * world if scaling == "weak" else args.runs
        ids = np.arange(rank, total, world)
        s_all, g_all = models.random_endpoints(robot, total, shrink=args.shrink)
        starts, goals = np.ascontiguousarray(s_all[ids]), np.ascontiguousarray(g_all[ids])
        bb = BatchBench(torch, eng, stream, device, robot, params, [sid], starts, goals)
        batch = bb.batch

        def step(ev=None):
            if ev is not None:
                ev[0].record(stream)
            batch.reset()
            if ev is not None:
                ev[1].record(stream)
            batch.iterate_async(N_ITER)
            if ev is not None:
                ev[2].record(stream)
            if world > 1:
                # the one collective of the path: best cost over all GPUs + winner's trajectory
                idx, cost = batch.best()
                tr = torch.empty((P, n), dtype=torch.float64, device=device)
                if idx >= 0:
                    batch.copy_run_traj_device(idx, tr.data_ptr())
                sharding.gather_best(cost if idx >= 0 else float("inf"), int(ids[idx]) if idx >= 0 else -1, tr, P, n, device)
            if ev is not None:
                ev[3].record(stream)

        for _ in range(warmup):
            step()
        barrier()
        launches0 = eng.launch_count()
        sampler = ClockSampler(local_rank)
        sampler.start()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        barrier()
        t_wall0 = time.perf_counter()
        for k in range(steps):
            flush.fill_(k & 0xFF)  # evict L2 between timed steps (not inside the event brackets)
            step(evs[k])
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
        launches = eng.launch_count() - launches0
        step_ms = sum(e[0].elapsed_time(e[3]) for e in evs)
        kern_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
        t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, kern_ms = float(t[0]), float(t[1])
        costs, status = batch.get_costs()
        # run-iterations actually performed per step: a run that leaves the joint limits stops there, as in
        # the reference (mod.cpp:2799-2803); the CPU arms count the same way
        local_done = float(batch.get_iterations().sum())
        d = torch.tensor([local_done, float((status != 0).sum())], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
        done, n_failed = float(d[0]), int(d[1])
        # ---- end to end through the C ABI with host buffers ----
        e2e_steps = max(2, min(steps, 3))
        R = len(ids)
        out_traj = torch.empty((R, P, n), dtype=torch.float64, pin_memory=True).numpy()
        starts_h, goals_h = pinned(torch, starts), pinned(torch, goals)
        barrier()
        t_e2e = bb.e2e(N_ITER, e2e_steps, out_traj, starts_h, goals_h)
        te = torch.tensor([t_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e2e = float(te[0])
        kernel_kind = "run-time specialised (NVRTC)" if batch.uses_jit() else "library instantiation"
        bb.close()
        return dict(step_ms=step_ms / steps, kern_ms=kern_ms / steps, done=done, local_done=local_done, failed=n_failed,
                    clocks=clocks, launches=launches, t_wall=t_wall, e2e_value=done / t_e2e, R=R,
                    h2d=2 * R * n * 8, d2h=R * P * n * 8 + R * 3 * 8 + R * 4, kernel_kind=kernel_kind,
                    starts=starts, goals=goals)

    warm = max(args.warmup, 3)
    h = run_headline(args.scaling, args.steps, warm)
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_extras:
        s = run_headline("strong", max(args.steps, 5), warm)
        strong = {"metric": METRIC, "value": s["done"] / (s["step_ms"] * 1e-3), "unit": UNIT, "scaling": "strong",
                  "runs_total": args.runs, "runs_per_gpu": s["R"], "ms_per_step": s["step_ms"],
                  "kernel_ms_per_step": s["kern_ms"], "e2e": {"value": s["e2e_value"], "unit": UNIT},
                  "runs_failed_joint_limits": s["failed"]}

    extras = {}
    want = [w for w in args.only.split(",") if w] or ["cfg3", "cfg4", "cfg5", "hbm", "tsr"]
    with_cpu = not args.no_cpu_baseline
    if rank == 0 and world == 1 and not args.no_extras:   # single-GPU measurements: not repeated under torchrun
        if "cfg3" in want:
            extras["cfg3_sdf_build"] = measure_sdf_build(torch, eng, stream, device, with_cpu)
            eng.trim()
        if "cfg4" in want:
            extras["cfg4_hmc"] = measure_cfg4(torch, eng, stream, device, sid, sd, with_cpu)
        if "cfg5" in want:
            extras["cfg5_dense"] = measure_cfg5(torch, eng, stream, device, with_cpu)
        if "hbm" in want:
            extras["cfg2_hbm_field"] = measure_hbm_field(torch, eng, stream, device, with_cpu)
            eng.trim()
        if "tsr" in want:
            extras["cfg2_tsr_constraint"] = measure_tsr(torch, eng, stream, device, sid, sd, with_cpu)
            eng.trim()

    if rank == 0:
        value = h["done"] / (h["step_ms"] * 1e-3)
        abytes = algorithmic_bytes_per_run_iter(P, n, robot.n_spheres_active, 1, False)
        tr = load_json("profiles", "traffic.json")
        roofline = roofline_record(abytes, h["local_done"], h["kern_ms"], traffic=tr.get("chomp_iterate_kernel_bytes_per_launch"),
                                   kernel="chomp_iterate_jit" if h["kernel_kind"].startswith("run-time") else "chomp_iterate_kernel",
                                   note="fp64-issue/latency bound, not HBM bound: see roofline_fp64 for the second ceiling")
        ops = load_json("profiles", "r2_fp64_ops.json")
        fp64 = None
        if ops.get("run_iterations"):
            flop = (2.0 * ops["dfma"] + ops["dmul"] + ops["dadd"]) / ops["run_iterations"]
            tfl = flop * h["local_done"] / (h["kern_ms"] * 1e-3) / 1e12
            fp64 = {"bound": "fp64", "flop_per_run_iteration": flop, "achieved": tfl, "achieved_tflops": tfl,
                    "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tfl / FP64_PEAK_TFLOPS,
                    "fp64_instructions_per_run_iteration": (ops["dfma"] + ops["dmul"] + ops["dadd"]) / ops["run_iterations"],
                    "source": "profiles/r2_fp64_ops.json (ncu smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on of "
                              "this launch; FMA = 2 FLOP); peak: profiles/r1_microbench.txt"}
        cpu = None
        if with_cpu:
            from oracle import pyoracle as po
            flavour = po.best_flavour()
            k = CPU_SAMPLE_RUNS
            s_all, g_all = models.random_endpoints(robot, k)
            v, dt = cpu_baseline(flavour, robot, params, [sd], s_all, g_all, N_ITER, threads=1)
            cpu = {"value": v, "unit": UNIT, "cores": 1,
                   "kind": "reference" if flavour == "reference" else "port",
                   "sample": "first %d runs of the batch x %d iterations, one thread (%.1f s)" % (k, N_ITER, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": h["step_ms"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(world, args.scaling, h["R"]), kernel=h["kernel_kind"]), "clocks": h["clocks"],
            "e2e": {"value": h["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": h["h2d"], "d2h_bytes_per_step": h["d2h"],
                    "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"},
            "gpu_launches": h["launches"], "roofline": roofline, "roofline_fp64": fp64, "cpu_baseline": cpu,
            "runs_failed_joint_limits": h["failed"], "run_iterations_per_step": h["done"],
            "wall_s_timed_region": h["t_wall"], "kernel_ms_per_step": h["kern_ms"],
            "strong_scaling": strong, "configs": extras,
            "sdf_build": extras.get("cfg3_sdf_build"),
        }
        print(json.dumps(line))
    eng.remove_sdf(sid)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
