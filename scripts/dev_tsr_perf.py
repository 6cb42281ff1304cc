"""Throughput of the constrained iteration (TSR on every waypoint) next to the CPU oracle.
Development aid; run on the GPU box:  python scripts/dev_tsr_perf.py [R]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from or_cdchomp_b200 import capi, models  # noqa: E402
from or_cdchomp_b200.engine import Engine  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n_iter = 20
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    fl = po.best_flavour()
    obs, sdf = po.computedistancefield(capi.make_prims(gp), len(gp), sizes, lengths, 0.02, flavour=fl)
    sd = capi.SdfDesc(sdf, lengths, models.pose_compose(kin_pose, gpose))
    ee = robot.names.index("wam7")
    base = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    rng = np.random.default_rng(1)
    starts = np.repeat(base[None], R, 0)
    goals = starts.copy()
    goals[:, 0] += rng.uniform(0.6, 1.3, R)
    goals[:, 1:] += rng.uniform(-0.05, 0.05, (R, 6))
    pe = po.fk(robot, base, flavour=fl)[ee]
    T0w, Twe = models.pose_make((0, 0, pe[2])), models.pose_make((0, 0, 0), pe[3:7])
    e = Engine(0)
    sid = e.upload_sdf(sd)
    for label, held, where in (("none", (), "all"), ("start k=3", ("z", "roll", "pitch"), "start"),
                               ("all k=1", ("z",), "all"), ("all k=3", ("z", "roll", "pitch"), "all")):
        Bw = np.tile(np.array([-10.0, 10.0]), (6, 1))
        for h in held:
            Bw[["x", "y", "z", "roll", "pitch", "yaw"].index(h)] = 0.0
        cons = [capi.make_constraint(where, ee, Bw, T0w=T0w, Twe=Twe)] if held else []
        params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0, constraints=cons)
        b = e.create_batch(robot, params, [sid], starts, goals)
        b.iterate(2)
        e.sync()
        t0 = time.perf_counter()
        costs, status = b.iterate(n_iter)
        e.sync()
        dt = time.perf_counter() - t0
        ok = int((status == 0).sum())
        t1 = time.perf_counter()
        n_cpu = 2
        for r in range(n_cpu):
            run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=fl)
            run.iterate(n_iter)
            run.close()
        dc = time.perf_counter() - t1
        print("%-10s rows %4d  GPU %8.0f run-it/s (%d/%d ok)   CPU 1 core %7.1f run-it/s   ratio %.0f" % (
            label, 98 * len(held) if where == "all" else len(held), R * n_iter / dt, ok, R, n_cpu * n_iter / dc,
            (R * n_iter / dt) / (n_cpu * n_iter / dc)), flush=True)
        b.close()
    e.close()


if __name__ == "__main__":
    main()
