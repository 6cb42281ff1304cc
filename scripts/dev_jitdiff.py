import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params0, gprims, sizes, lengths, pose_world = bench.build_scene()
engA = Engine(0)
obs, sdf = engA.computedistancefield(gprims, sizes, lengths, 0.02)
sd = capi.SdfDesc(sdf, lengths, pose_world)
sidA = engA.upload_sdf(sd)
# something else runs first (as the earlier tests of the file do)
p0 = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
s0, g0 = models.random_endpoints(robot, 8)
b = engA.create_batch(robot, p0, [sidA], s0, g0); b.iterate(100); b.close()
eng = Engine(0)
sid = eng.upload_sdf(sd)
params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
qs, qg = models.random_endpoints(robot, 6, seed0=123, shrink=0.3)
hist = []
for jit in (False, True, False, True):
    eng.enable_jit(jit)
    b = eng.create_batch(robot, params, [sid], qs, qg)
    tr = []
    for it in range(25):
        b.iterate(1)
        tr.append(b.get_traj().copy())
    hist.append(np.array(tr))
    b.close()
for a, bb, name in ((0, 1, "static vs jit"), (0, 2, "static vs static"), (1, 3, "jit vs jit")):
    d = np.abs(hist[a] - hist[bb]).reshape(25, -1).max(axis=1)
    first = int(np.argmax(d > 0)) if d.max() > 0 else -1
    print(name, "max diff", d.max(), "first differing iteration", first)
    if first >= 0:
        idx = np.unravel_index(np.argmax(np.abs(hist[a][first] - hist[bb][first])), hist[a][first].shape)
        print("   at (run, waypoint, dof)", idx)
