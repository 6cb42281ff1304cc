#!/usr/bin/env python
"""bench.py -- batched CHOMP run-iterations/s (WAM7), BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU)

One "step" = the hot path over one batch: every GPU re-arms its shard of
4096 independent WAM7 runs (random start/goal, n_points=100, table SDF) with
straight-line trajectories and runs 100 CHOMP iterations + the final cost pass
(`iterate run ... n_iter 100`).  Weak scaling: 4096 runs PER GPU.

  value  run-iterations/s with the end points already in HBM (CUDA events on the
         launching stream, max over ranks, summed over the K steps)
  e2e    the same metric through the public C ABI with HOST buffers:
         ocb_batch_create (H2D end points) -> ocb_batch_iterate -> ocb_batch_get_traj
         (D2H trajectories + costs) -> ocb_batch_destroy, wall clock
  roofline  algorithmic bytes (58 128 B per run-iteration) / CUDA-event
         time of the iterate kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference CPU path (oracle/_ref: the reference's own libcd +
         restated callbacks) on ONE host core, bounded sample, rank 0 only

--impl reference times that CPU path on all host cores (one run per thread).
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


def algorithmic_bytes_per_run_iter(P, n, n_active, n_sdf, momentum):
    """8*[P*n + m*n + 2*m*n*momentum + 4*m*S_a*K]."""
    m = P - 2
    return 8 * (P * n + m * n + (2 * m * n if momentum else 0) + 4 * m * n_active * n_sdf)


def build_scene():
    from or_cdchomp_b200 import capi, models
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gprims = models.prims_to_grid_frame(prims, gpose)
    pose_world = models.pose_compose(kin_pose, gpose)
    params = capi.default_params(n_points=N_POINTS, lambda_=LAMBDA, obs_factor=OBS_FACTOR)
    return robot, params, gprims, sizes, lengths, pose_world


def config_dict(n_gpus):
    return {
        "workload": "BASELINE configs[1]: WAM7 batch of 4096 independent runs per GPU "
                    "(random start/goal, n_points=100, 100 iterations, one table SDF 31x40x11)",
        "runs_per_gpu": RUNS_PER_GPU, "n_points": N_POINTS, "n_iter_per_step": N_ITER,
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


def cpu_baseline(flavour, robot, params, sd, n_runs, n_iter, threads=1):
    """run-iterations/s of the CPU oracle on `threads` host threads (one run each at a time)."""
    with _StdoutToStderr():
        return _cpu_baseline(flavour, robot, params, sd, n_runs, n_iter, threads)


def _cpu_baseline(flavour, robot, params, sd, n_runs, n_iter, threads=1):
    from oracle import pyoracle as po
    from or_cdchomp_b200 import models
    starts, goals = models.random_endpoints(robot, n_runs)
    po.load(flavour)
    done = [0] * n_runs

    def work(r):
        run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=flavour)
        t_iter = 0
        ret, _, tr, _ = run.iterate(n_iter, want_trace=True)
        if ret == 0:
            t_iter = n_iter
        else:  # the reference aborts the run when it leaves the joint limits (mod.cpp:2799-2803)
            t_iter = int(np.count_nonzero(tr[:, 0])) or 1
        run.close()
        done[r] = t_iter

    t0 = time.perf_counter()
    if threads <= 1:
        # one core, pinned (`taskset -c`), affinity restored afterwards
        saved = None
        try:
            saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, {sorted(saved)[0]})
        except (AttributeError, OSError):
            saved = None
        try:
            for r in range(n_runs):
                work(r)
        finally:
            if saved is not None:
                os.sched_setaffinity(0, saved)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(n_runs)))
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt


def measure_sdf_build(eng, stream, device, n=400):
    """cd_grid_double_bin_sdf equivalent on a synthetic cluttered 400^3 grid: occupancy grid in HBM ->
    SDF in HBM (16 algorithmic bytes per voxel)."""
    import torch
    from or_cdchomp_b200 import models
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
    # voxelisation and flood fill, reported separately
    d_occ = torch.empty(ncell, dtype=torch.float64, device=device)
    occ_ms = timed(lambda: eng.occupancy_device(gp, sizes, lengths, ce, d_occ.data_ptr()))

    def flood():
        d_obs.copy_(d_occ)
        eng.flood_relabel_device(d_obs.data_ptr(), sizes, 0)
    copy_ms = timed(lambda: d_obs.copy_(d_occ))
    flood_ms = max(timed(flood) - copy_ms, 0.0)
    del d_occ
    times = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.sdf_build_device(d_obs.data_ptr(), sizes, lengths, d_sdf.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = 16.0 * ncell / (ms * 1e-3) / 1e9
    del d_obs, d_sdf
    return {"metric": "sdf_build_mvoxels_per_s", "value": ncell / (ms * 1e-3) / 1e6, "unit": "Mvoxels/s",
            "sizes": [int(x) for x in sizes], "ms": ms,
            "occupancy_ms": occ_ms, "flood_relabel_ms": flood_ms, "n_primitives": len(gp),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "algorithmic_bytes_per_voxel": 16}}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from or_cdchomp_b200 import capi
    from oracle import pyoracle as po
    robot, params, gprims, sizes, lengths, pose_world = build_scene()
    flavour = po.best_flavour()
    pa = capi.make_prims(gprims)
    _, sdf = po.computedistancefield(pa, len(gprims), sizes, lengths, 0.02, flavour=flavour)
    sd = capi.SdfDesc(sdf, lengths, pose_world)
    cores = os.cpu_count() or 1
    runs_per_step = cores * 32
    for _ in range(args.warmup):
        cpu_baseline(flavour, robot, params, sd, min(cores, 4), 10, threads=cores)
    t_total, it_total = 0.0, 0.0
    for _ in range(args.steps):
        v, dt = cpu_baseline(flavour, robot, params, sd, runs_per_step, N_ITER, threads=cores)
        t_total += dt
        it_total += v * dt
    value = it_total / t_total
    sample = "%d runs x %d iterations per step on %d threads (one run per thread)" % (runs_per_step, N_ITER, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": "reference" if flavour == "reference" else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--runs", type=int, default=RUNS_PER_GPU, help="runs per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sdf", action="store_true", help="skip the secondary SDF-build measurement")
    ap.add_argument("--no-jit", action="store_true", help="use the library's own kernel instead of the run-time specialised one")
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
    R = args.runs
    total_runs = R * world
    lo, hi = sharding.shard_bounds(total_runs, rank, world)
    # every rank draws only its own shard (run ids are global, so results do not depend on N)
    starts = np.empty((R, robot.n_dof))
    goals = np.empty((R, robot.n_dof))
    s_all, g_all = models.random_endpoints(robot, hi)  # cheap; keeps seeds global
    starts[:], goals[:] = s_all[lo:hi], g_all[lo:hi]

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
    batch = eng.create_batch(robot, params, [sid], starts, goals)
    kernel_kind = "run-time specialised (NVRTC)" if batch.uses_jit() else "library instantiation"
    P, n = params.n_points, robot.n_dof
    d_traj, d_costs = batch.device_ptrs()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(timed_events=None):
        if timed_events is not None:
            timed_events[0].record(stream)
        batch.reset()
        if timed_events is not None:
            timed_events[1].record(stream)
        batch.iterate_async(N_ITER)
        if timed_events is not None:
            timed_events[2].record(stream)
        if world > 1:
            # the one collective of the path: best cost over all GPUs + winner's trajectory
            idx, cost = batch.best()
            tr = torch.empty((P, n), dtype=torch.float64, device=device)
            if idx >= 0:
                batch.copy_run_traj_device(idx, tr.data_ptr())
            sharding.gather_best(cost if idx >= 0 else float("inf"), lo + idx if idx >= 0 else -1, tr, P, n, device)
        if timed_events is not None:
            timed_events[3].record(stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
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
    n_failed = int((status != 0).sum())
    # run-iterations actually performed per step: a run that leaves the joint limits stops there, as in
    # the reference (mod.cpp:2799-2803); the CPU arms count the same way
    done = torch.tensor([int(batch.get_iterations().sum())], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(done, op=dist.ReduceOp.SUM)
    run_iters_done = float(done[0])

    # ---- end to end through the C ABI with host buffers ----
    e2e_steps = max(2, min(args.steps, 3))
    # caller-side host buffers are page-locked, as a planner that cares about latency would hold them
    out_traj = torch.empty((R, P, n), dtype=torch.float64, pin_memory=True).numpy()
    starts_h = torch.from_numpy(np.ascontiguousarray(starts)).pin_memory().numpy()
    goals_h = torch.from_numpy(np.ascontiguousarray(goals)).pin_memory().numpy()
    def e2e_step():
        b2 = eng.create_batch(robot, params, [sid], starts_h, goals_h)
        b2.iterate(N_ITER)
        b2.get_traj(out_traj)
        b2.close()
    e2e_step()  # warm-up (untimed), as for the device-timed steps
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te[0])
    e2e_value = run_iters_done * e2e_steps / t_e2e
    h2d = 2 * R * n * 8
    d2h = R * P * n * 8 + R * 3 * 8 + R * 4

    # ---- secondary metric: SDF build Mvoxels/s, BASELINE configs[2] (400^3, device resident) ----
    sdf_build = None
    if rank == 0 and not args.no_sdf:
        sdf_build = measure_sdf_build(eng, stream, device)

    run_iters_per_step = run_iters_done
    value = run_iters_per_step * args.steps / (step_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes_per_run_iter(P, n, robot.n_spheres_active, 1, False)
        launch_s = (kern_ms * 1e-3) / args.steps
        achieved = abytes * (run_iters_done / world) / launch_s / 1e9  # units one launch processed (this GPU's share)
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("chomp_iterate_kernel_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "chomp_iterate_kernel",
                    "algorithmic_bytes_per_run_iteration": abytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "note": "fp64-issue/latency bound, not HBM bound: see DESIGN.md for the second ceiling"}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import pyoracle as po
            flavour = po.best_flavour()
            v, dt = cpu_baseline(flavour, robot, params, sd, CPU_SAMPLE_RUNS, N_ITER, threads=1)
            cpu = {"value": v, "unit": UNIT, "cores": 1,
                   "kind": "reference" if flavour == "reference" else "port",
                   "sample": "first %d runs of the batch x %d iterations, one thread (%.1f s)" % (CPU_SAMPLE_RUNS, N_ITER, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(world), kernel=kernel_kind), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "create+iterate+gettraj+destroy through the C ABI, pinned host buffers, wall clock"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "runs_failed_joint_limits": n_failed, "run_iterations_per_step": run_iters_done, "wall_s_timed_region": t_wall,
            "kernel_ms_per_step": kern_ms / args.steps, "sdf_build": sdf_build,
        }
        print(json.dumps(line))
    batch.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
