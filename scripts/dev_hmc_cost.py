import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sid = eng.upload_sdf(capi.SdfDesc(sdf, lengths, pose_world))
R = 8192
starts, goals = models.random_endpoints(robot, 1, shrink=0.3)
st, go = np.repeat(starts, R, 0), np.repeat(goals, R, 0)
for kw in (dict(), dict(use_momentum=1), dict(use_momentum=1, use_hmc=1, hmc_resample_lambda=0.02), dict(use_momentum=1, use_hmc=1, hmc_resample_lambda=0.2)):
    p = capi.default_params(n_points=256, lambda_=100.0, obs_factor=500.0, **kw)
    b = eng.create_batch(robot, p, [sid], st, go, seeds=np.arange(1, R + 1))
    b.iterate(2); eng.sync(); b.reset()
    t = time.time(); b.iterate_async(100); eng.sync(); dt = time.time() - t
    print(kw, "%.3f s -> %.3g run-iter/s" % (dt, R * 100 / dt))
    b.close()
