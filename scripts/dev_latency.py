"""Single-run latency through the C ABI and the module (BASELINE configs[0]): create / iterate / gettraj / destroy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models, orcdchomp
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sid = eng.upload_sdf(capi.SdfDesc(sdf, lengths, pose_world))
s, g = np.array([models.WAM7_DEMO_START]), np.array([models.WAM7_DEMO_GOAL])
for rep in range(3):
    t0 = time.perf_counter(); b = eng.create_batch(robot, params, [sid], s, g); t1 = time.perf_counter()
    b.iterate(100); t2 = time.perf_counter()
    tr = b.get_traj(); t3 = time.perf_counter()
    b.close(); t4 = time.perf_counter()
    print("C ABI  R=1: create %.2f ms, iterate(100) %.2f ms, gettraj %.2f ms, destroy %.2f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3)))
b = eng.create_batch(robot, params, [sid], s, g)
for n in (1, 10, 100):
    b.reset(); eng.sync()
    t0 = time.perf_counter(); b.iterate(n); t1 = time.perf_counter()
    print("iterate(%d): %.3f ms" % (n, 1e3 * (t1 - t0)))
b.close()
env = orcdchomp.Environment()
kin_pose, prims, apos, aext = models.table_scene()
table = env.AddKinBody("table", kin_pose, prims)
rb = env.AddRobot("BarrettWAM", models.wam7_robot(), models.WAM7_DEMO_START)
mod = orcdchomp.Module(env, 0)
t0 = time.perf_counter(); mod.computedistancefield(kinbody=table, cube_extent=0.02); t1 = time.perf_counter()
print("module computedistancefield (table): %.2f ms" % (1e3 * (t1 - t0)))
for rep in range(3):
    t0 = time.perf_counter()
    traj = mod.runchomp(robot=rb, n_iter=100, lambda_=100.0, obs_factor=500.0, n_points=100, adofgoal=list(models.WAM7_DEMO_GOAL), no_collision_check=True)
    t1 = time.perf_counter()
    print("module runchomp (100 its): %.2f ms" % (1e3 * (t1 - t0)))
