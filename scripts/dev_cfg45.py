"""Timing of BASELINE configs[3] (HMC seeds, n_points=256) and configs[4] (dense spheres, n_points=1024)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sd = capi.SdfDesc(sdf, lengths, pose_world)
sid = eng.upload_sdf(sd)
# config 4
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p4 = capi.default_params(n_points=256, lambda_=100.0, obs_factor=500.0, use_momentum=1, use_hmc=1, hmc_resample_lambda=0.02)
starts, goals = models.random_endpoints(robot, 1, shrink=0.3)
st, go = np.repeat(starts, R, 0), np.repeat(goals, R, 0)
b = eng.create_batch(robot, p4, [sid], st, go, seeds=np.arange(1, R + 1))
b.iterate(2); eng.sync(); b.reset()
t = time.time(); b.iterate_async(100); eng.sync(); dt = time.time() - t
costs, status = b.get_costs()
print("cfg4: R=%d P=256 momentum+hmc 100 its: %.3f s -> %.3g run-iter/s; failed %d; best %s" % (R, dt, R * 100 / dt, (status != 0).sum(), b.best()))
b.close()
# config 5
robot5 = models.dense_sphere_arm(200, seed=5)
rng = np.random.default_rng(9)
ids = []
for k in range(4):
    f = rng.uniform(0.05, 0.6, size=(128, 128, 128))
    pose = models.pose_make(rng.uniform(-1.2, -0.6, size=3), models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 1.0)))
    ids.append(eng.upload_sdf(capi.SdfDesc(f, [2.0, 2.0, 2.0], pose)))
R5 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
p5 = capi.default_params(n_points=1024, lambda_=200.0, obs_factor=100.0)
starts, goals = models.random_endpoints(robot5, R5, shrink=0.3)
b = eng.create_batch(robot5, p5, ids, starts, goals)
b.iterate(1); eng.sync(); b.reset()
t = time.time(); b.iterate_async(10); eng.sync(); dt = time.time() - t
costs, status = b.get_costs()
print("cfg5: R=%d P=1024 S=200 K=4, 10 its: %.3f s -> %.3g run-iter/s; failed %d" % (R5, dt, R5 * 10 / dt, (status != 0).sum()))
b.close()
