"""Small launch of the tiled large-robot path for ncu (config-5 shape: 200 spheres, 4 SDFs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
eng = Engine(0)
robot5 = models.dense_sphere_arm(200, seed=5)
rng = np.random.default_rng(9)
ids = []
for k in range(4):
    f = rng.uniform(0.05, 0.6, size=(128, 128, 128))
    pose = models.pose_make(rng.uniform(-1.2, -0.6, size=3), models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 1.0)))
    ids.append(eng.upload_sdf(capi.SdfDesc(f, [2.0, 2.0, 2.0], pose)))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
p5 = capi.default_params(n_points=P, lambda_=200.0, obs_factor=100.0)
starts, goals = models.random_endpoints(robot5, R, shrink=0.3)
b = eng.create_batch(robot5, p5, ids, starts, goals)
b.iterate(2)
eng.sync()
print("done")
