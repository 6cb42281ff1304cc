"""Create / iterate / destroy batches of varying shapes many times; device memory in use must plateau."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sd = capi.SdfDesc(sdf, lengths, pose_world)
rng = np.random.default_rng(0)
used = []
for cycle in range(120):
    sid = eng.upload_sdf(sd)
    R = int(rng.integers(1, 600)); P = int(rng.integers(5, 140))
    p = capi.default_params(n_points=P, lambda_=100.0, obs_factor=300.0, use_momentum=int(rng.integers(0, 2)))
    s, g = models.random_endpoints(robot, R, seed0=cycle, shrink=0.3)
    b = eng.create_batch(robot, p, [sid], s, g)
    b.iterate(3)
    b.close()
    eng.remove_sdf(sid)
    if cycle % 10 == 9:
        free, total = torch.cuda.mem_get_info()
        used.append((total - free) / 2**20)
print("MiB in use every 10 cycles:", [round(u) for u in used])
eng.trim()
free, total = torch.cuda.mem_get_info()
print("after trim:", round((total - free) / 2**20), "MiB")
