import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models, orcdchomp
from or_cdchomp_b200.engine import Engine
env = orcdchomp.Environment()
kin_pose, prims, apos, aext = models.table_scene()
table = env.AddKinBody("table", kin_pose, prims)
rb = env.AddRobot("BarrettWAM", models.wam7_robot(), models.WAM7_DEMO_START)
mod = orcdchomp.Module(env, 0)
for rep in range(4):
    t0 = time.perf_counter(); mod.computedistancefield(kinbody=table, cube_extent=0.02); t1 = time.perf_counter()
    print("module computedistancefield (table): %.2f ms" % (1e3 * (t1 - t0)))
    mod.removefield(kinbody=table)
for ce in (0.01, 0.005):
    t0 = time.perf_counter(); mod.computedistancefield(kinbody=table, cube_extent=ce); t1 = time.perf_counter()
    print("cube_extent %g: %.2f ms" % (ce, 1e3 * (t1 - t0)))
    mod.removefield(kinbody=table)
eng = Engine(0)
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
for rep in range(3):
    t0 = time.perf_counter(); obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02); t1 = time.perf_counter()
    print("engine computedistancefield_host:", sizes, "%.2f ms" % (1e3 * (t1 - t0)))
