"""Per-warp cycle totals of the phases of the persistent kernel (OCB_JIT_FLAGS=-DOCB_PHASE_CLOCKS).
Prints the share of each phase, per warp of the block, averaged over the runs of a full bench batch."""
import os, sys
os.environ["OCB_JIT_FLAGS"] = os.environ.get("OCB_JIT_FLAGS", "") + " -DOCB_PHASE_CLOCKS"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from or_cdchomp_b200 import capi, models
from or_cdchomp_b200.engine import Engine
import bench
robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
eng = Engine(0)
eng.enable_jit(True)
obs, sdf = eng.computedistancefield(gprims, sizes, lengths, 0.02)
sid = eng.upload_sdf(capi.SdfDesc(sdf, lengths, pose_world))
R = 4096
starts, goals = models.random_endpoints(robot, R)
b = eng.create_batch(robot, params, [sid], starts, goals)
assert b.uses_jit()
b.enable_trace(True)
b.iterate(100)
b.reset()
b.iterate(100)
tr = b.get_trace(100).reshape(R, 300)
names = ["FK", "B1 wait", "pairs", "spheres", "flush", "stencil", "B2 wait", "solve", "B3 wait", "update", "B4 wait",
         "(limits)", "smooth", "reduce+B5", "-", "loop/hmc"]
nw = 4
acc = tr[:, :nw * 16].reshape(R, nw, 16).mean(axis=0)
tot = acc.sum(axis=1)
print("cycles per iteration per warp:", (tot / 100).round(0))
for k in range(16):
    print("%-10s " % names[k] + "  ".join("%5.1f%%" % (100 * acc[w, k] / tot[w]) for w in range(nw)) + "   %8.0f cyc/iter (warp 0)" % (acc[0, k] / 100))

# distribution over runs: the slowest runs set the launch time of a small batch
per_run = tr[:, :nw * 16].reshape(R, nw, 16)[:, 0, :]
tot_run = per_run.sum(axis=1)
its = b.get_iterations()
rounds = b.get_limit_rounds()
order = np.argsort(-tot_run)
print("total cycles per run: median %.3g  p90 %.3g  p99 %.3g  max %.3g  (sum of the 8 slowest / sum of all: %.3f)" % (
    np.median(tot_run), np.percentile(tot_run, 90), np.percentile(tot_run, 99), tot_run.max(), tot_run[order[:8]].sum() / tot_run.sum()))
for r in order[:8]:
    print("  run %4d: %.3g cycles (%.1fx median), iterations %3d, max limit rounds %4d, limits phase %.0f%%" % (
        r, tot_run[r], tot_run[r] / np.median(tot_run), its[r], rounds[r], 100 * per_run[r, 11] / tot_run[r]))
print("runs over 2x median: %d, over 1.3x: %d" % ((tot_run > 2 * np.median(tot_run)).sum(), (tot_run > 1.3 * np.median(tot_run)).sum()))
