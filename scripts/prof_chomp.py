"""Small launch of the CHOMP kernel for ncu (one wave: 296 runs, 10 iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sid = eng.upload_sdf(capi.SdfDesc(sdf, lengths, pose_world))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 296
its = int(sys.argv[2]) if len(sys.argv) > 2 else 10
starts, goals = models.random_endpoints(robot, R, shrink=float(os.environ.get('SHRINK', '0.3')))
b = eng.create_batch(robot, params, [sid], starts, goals)
b.iterate(its)
b.reset()
b.iterate(its)
eng.sync()
print("done")
