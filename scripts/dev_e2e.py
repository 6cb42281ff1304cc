import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sid = eng.upload_sdf(capi.SdfDesc(sdf, lengths, pose_world))
R, P, n = 4096, 100, 7
starts, goals = models.random_endpoints(robot, R)
out = torch.empty((R, P, n), dtype=torch.float64, pin_memory=True).numpy()
def leg(tag):
    for rep in range(4):
        t0 = time.perf_counter(); b = eng.create_batch(robot, params, [sid], starts, goals); t1 = time.perf_counter()
        b.iterate(100); t2 = time.perf_counter()
        b.get_traj(out); t3 = time.perf_counter()
        b.close(); t4 = time.perf_counter()
        print("%s: create %.2f iterate %.2f gettraj %.2f destroy %.2f total %.2f ms (jit %s)" % ((tag,) + tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)) + (b.uses_jit(),)))
leg("static")
eng.enable_jit(True)
leg("jit   ")
eng.enable_jit(False)
leg("static")
