"""Developer check run on the GPU box: parity vs oracle + first timings."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
from oracle import pyoracle as po

fl = po.best_flavour()
print("oracle flavour", fl)
robot = models.wam7_robot()
kin_pose, prims, apos, aext = models.table_scene()
sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
gprims = models.prims_to_grid_frame(prims, gpose)
eng = Engine(0)
t = time.time()
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
print("gpu cdf", time.time() - t)
pa = capi.make_prims(gprims)
obs_ref, sdf_ref = po.computedistancefield(pa, len(gprims), sizes, lengths, 0.02, flavour=fl)
print("occupancy equal:", np.array_equal(obs, obs_ref), " sdf maxdiff:", np.nanmax(np.abs(sdf - sdf_ref)))

# random sdf build parity incl. anisotropic + non-binary
rng = np.random.default_rng(1)
for shape, lens in (((17, 9, 23), (1.7, 0.9, 2.3)), ((20, 31, 12), (1.0, 2.0, 0.7)), ((8, 8, 8), (1, 1, 1))):
    o = np.where(rng.uniform(size=shape) < 0.1, np.inf, 0.0)
    s_gpu = eng.sdf_build(o, lens)
    s_ref = po.sdf_from_obsarray(o, lens, flavour=fl)
    print("sdf", shape, "maxdiff", np.max(np.abs(s_gpu - s_ref)), "bit-equal", np.array_equal(s_gpu, s_ref))

sd = capi.SdfDesc(sdf_ref, lengths, models.pose_compose(kin_pose, gpose))
params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
R = 8
starts, goals = models.random_endpoints(robot, R)
starts[0], goals[0] = models.WAM7_DEMO_START, models.WAM7_DEMO_GOAL
sid = eng.upload_sdf(sd)
batch = eng.create_batch(robot, params, [sid], starts, goals)
batch.enable_trace(True)
batch.capture_gradient(1)
# iteration-by-iteration gradient parity on run 0..R
runs = [po.Run(robot, params, [sd], starts[r], goals[r], flavour=fl) for r in range(R)]
worst_g = 0
for it in range(3):
    costs, status = batch.iterate(1)
    G = batch.get_gradient()
    for r in range(R):
        ret, c, tr, gr = runs[r].iterate(1, want_trace=True, want_grads=True)
        rel = np.max(np.abs(G[r] - gr[0])) / np.max(np.abs(gr[0]))
        worst_g = max(worst_g, rel)
        if r < 2:
            print("it", it, "run", r, "grad rel", rel, "traj diff", np.max(np.abs(batch.get_traj()[r] - runs[r].traj())),
                  "cost", costs[r], c)
print("worst gradient rel err", worst_g)
batch.close()
for r in runs: r.close()

# 100-iteration parity
batch = eng.create_batch(robot, params, [sid], starts, goals)
batch.enable_trace(True)
costs, status = batch.iterate(100)
traj = batch.get_traj()
trace = batch.get_trace(100)
worst = 0
for r in range(R):
    run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=fl)
    ret, c, tr, _ = run.iterate(100, want_trace=True)
    d = np.max(np.abs(run.traj() - traj[r]))
    worst = max(worst, d)
    print("run", r, "ret", ret, status[r], "traj maxdiff", d, "cost", c[0], costs[r, 0], "trace maxdiff", np.max(np.abs(tr - trace[r])))
    run.close()
print("worst traj diff after 100 its:", worst)
batch.close()

# timing
for R in (148 * 2, 4096):
    starts, goals = models.random_endpoints(robot, R)
    batch = eng.create_batch(robot, params, [sid], starts, goals)
    batch.iterate(2)
    eng.sync()
    t = time.time()
    batch.iterate_async(100)
    eng.sync()
    dt = time.time() - t
    print("R", R, "100 its:", dt, "s ->", R * 100 / dt, "run-iter/s")
    costs, status = batch.get_costs()
    print("  status nonzero:", int((status != 0).sum()), "mean cost", costs[:, 0].mean())
    batch.close()
eng.close()
