"""CPU-only census: how far do the reference's OWN two builds (oracle/_ref: reference libcd with OpenBLAS's
dgemm / explicit LU inverse; oracle/build: the same algorithm with plain loops) agree on BASELINE configs[1]?
Separates the runs by the number of joint-limit projection steps (chomp.c:608-655) they need.
Writes profiles/r2_limit_chaos_cpu.json.  ~3 minutes on 8 cores."""
import json, os, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np
from or_cdchomp_b200 import capi, models
from oracle import pyoracle as po
import bench

robot, params, gprims, sizes, lengths, pose_world = bench.build_scene()
pa = capi.make_prims(gprims)
_, sdf = po.computedistancefield(pa, len(gprims), sizes, lengths, 0.02, flavour="reference")
sd = capi.SdfDesc(sdf, lengths, pose_world)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
starts, goals = models.random_endpoints(robot, R)


def batch(flavour):
    out = [None] * R
    def work(r):
        run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=flavour)
        ret, c, _, _ = run.iterate(100)
        out[r] = (ret, c, run.traj())
        run.close()
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        list(ex.map(work, range(R)))
    return out

ref = batch("reference")
port = batch("port")
# per-run projection steps from the port, serially (its counter is a global)
rounds = np.zeros(R, dtype=int)
for r in range(R):
    po.debug_limit_rounds(reset=True)
    run = po.Run(robot, params, [sd], starts[r], goals[r], flavour="port")
    run.iterate(100)
    run.close()
    rounds[r] = po.debug_limit_rounds()
rr = np.array([o[0] for o in ref]); pr = np.array([o[0] for o in port])
both = (rr == 0) & (pr == 0)
err = np.zeros(R)
for r in np.where(both)[0]:
    err[r] = np.max(np.abs(ref[r][2] - port[r][2]))
res = dict(runs=R, both_ok=int(both.sum()), both_fail=int(((rr != 0) & (pr != 0)).sum()),
           ref_fail_port_ok=[int(x) for x in np.where((rr != 0) & (pr == 0))[0]],
           port_fail_ref_ok=[int(x) for x in np.where((rr == 0) & (pr != 0))[0]],
           n_err_over_1e9=int((err > 1e-9).sum()), n_err_over_1e6=int((err > 1e-6).sum()), max_err=float(err.max()),
           by_rounds={})
for lo, hi in ((0, 0), (1, 5), (6, 25), (26, 100), (101, 999), (1000, 1000)):
    sel = (rounds >= lo) & (rounds <= hi)
    res["by_rounds"]["%d-%d" % (lo, hi)] = dict(runs=int(sel.sum()), status_mismatch=int((sel & ((rr == 0) != (pr == 0))).sum()),
                                                 err_over_1e6=int((sel & both & (err > 1e-6)).sum()),
                                                 max_err=float(err[sel & both].max()) if (sel & both).any() else 0.0)
json.dump(res, open(os.path.join(ROOT, "profiles", "r2_limit_chaos_cpu.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
