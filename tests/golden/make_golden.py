"""Regenerates tests/golden/*.npz from the REFERENCE build of the oracle
(oracle/_ref/liboracle_ref.so = the unmodified /root/reference/src/libcd sources
+ oracle/orcdchomp_port.c).  Only runs where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures let the GPU box (which has no /root/reference) check both the
self-contained oracle port and the CUDA engine against reference outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from or_cdchomp_b200 import capi, models  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

FL = "reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def sdf_kat():
    rng = np.random.default_rng(11)
    grid = rng.normal(size=(6, 7, 5))
    grid[2, 3, 1] = np.inf
    grid[5, 6, 4] = np.inf
    lengths = np.array([1.2, 0.7, 2.0])
    pts = [rng.uniform(-0.05, 1.05, size=(400, 3)) * lengths]
    # boundaries, cell faces and cell-centre planes
    edge = []
    for ax in range(3):
        for v in (0.0, lengths[ax], lengths[ax] * 0.5, lengths[ax] / grid.shape[ax] * 1.5,
                  lengths[ax] / grid.shape[ax] * 2.0, -1e-12, lengths[ax] * (1 + 1e-12)):
            p = 0.37 * lengths
            p = p.copy()
            p[ax] = v
            edge.append(p)
    edge.append(np.zeros(3))
    edge.append(lengths.copy())
    pts.append(np.array(edge))
    pts = np.concatenate(pts)
    vals, grads, errs = po.sdf_sample(grid, lengths, pts, flavour=FL)
    np.savez_compressed(os.path.join(OUT, "sdf_kat.npz"), grid=grid, lengths=lengths, points=pts,
                        values=vals, grads=grads, errs=errs)


def sdf_build():
    rng = np.random.default_rng(5)
    cases = {}
    o = np.where(rng.uniform(size=(14, 11, 9)) < 0.08, np.inf, 0.0)
    cases["iso"] = (o, np.array([1.4, 1.1, 0.9]))
    o = np.where(rng.uniform(size=(9, 13, 10)) < 0.15, np.inf, 0.0)
    cases["aniso"] = (o, np.array([1.0, 2.6, 0.5]))
    o = np.where(rng.uniform(size=(8, 8, 12)) < 0.1, np.inf, 0.0)
    o[rng.uniform(size=o.shape) < 0.05] = 0.3  # finite non-zero heights (treated as obstacles by the flip)
    cases["heights"] = (o, np.array([0.8, 0.8, 1.2]))
    cases["allfree"] = (np.zeros((4, 5, 6)), np.array([1.0, 1.0, 1.0]))
    cases["allobs"] = (np.full((4, 5, 6), np.inf), np.array([1.0, 1.0, 1.0]))
    out = {}
    for k, (o, l) in cases.items():
        out[k + "_obs"] = o
        out[k + "_len"] = l
        out[k + "_sdf"] = po.sdf_from_obsarray(o, l, flavour=FL)
        out[k + "_dt"] = po.dt_sqeuc(o, l, flavour=FL)
    np.savez_compressed(os.path.join(OUT, "sdf_build.npz"), **out)


def occupancy():
    prims, apos, aext = models.clutter_scene(n_boxes=10, n_balls=6, seed=3, half_span=0.9)
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.04, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    occ = po.occupancy(pa, len(gp), sizes, lengths, 0.04, flavour=FL)
    obs, sdf = po.computedistancefield(pa, len(gp), sizes, lengths, 0.04, flavour=FL)
    np.savez_compressed(os.path.join(OUT, "occupancy.npz"), sizes=np.array(sizes), lengths=np.array(lengths),
                        occ_hit=np.packbits(np.isinf(occ)), obs_hit=np.packbits(np.isinf(obs)),
                        sdf=sdf.astype(np.float64))


def mesh():
    prims, apos, aext = models.mesh_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.1)
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    occ = po.occupancy(pa, len(gp), sizes, lengths, 0.02, flavour=FL)
    obs, sdf = po.computedistancefield(pa, len(gp), sizes, lengths, 0.02, flavour=FL)
    np.savez_compressed(os.path.join(OUT, "mesh.npz"), sizes=np.array(sizes), occ_hit=np.packbits(np.isinf(occ)),
                        obs_hit=np.packbits(np.isinf(obs)), sdf=sdf)


def chomp():
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    obs, sdf = po.computedistancefield(pa, len(gp), sizes, lengths, 0.02, flavour=FL)
    pose_world = models.pose_compose(kin_pose, gpose)
    sd = capi.SdfDesc(sdf, lengths, pose_world)
    out = dict(table_sdf=sdf, table_obs_hit=np.packbits(np.isinf(obs)), table_lengths=np.array(lengths),
               table_pose=pose_world)
    # config 1: demo start, 100 iterations
    params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
    starts, goals = models.random_endpoints(robot, 4)
    starts[0], goals[0] = models.WAM7_DEMO_START, models.WAM7_DEMO_GOAL
    trajs, traces, costs, g0 = [], [], [], []
    for r in range(4):
        run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=FL)
        ret, c, tr, gr = run.iterate(100, want_trace=True, want_grads=True)
        assert ret == 0
        trajs.append(run.traj()); traces.append(tr); costs.append(c); g0.append(gr[0])
        run.close()
    out.update(cfg1_starts=starts, cfg1_goals=goals, cfg1_traj=np.array(trajs), cfg1_trace=np.array(traces),
               cfg1_costs=np.array(costs), cfg1_grad0=np.array(g0))
    # momentum + hmc, shorter trajectory
    params = capi.default_params(n_points=40, lambda_=50.0, obs_factor=300.0, use_momentum=1, use_hmc=1,
                                 hmc_resample_lambda=0.05)
    trajs, costs, moms, nexts = [], [], [], []
    for seed in (0, 7, 123456):
        run = po.Run(robot, params, [sd], starts[1], goals[1], seed=seed, flavour=FL)
        ret, c, _, _ = run.iterate(60)
        assert ret == 0
        trajs.append(run.traj()); costs.append(c); moms.append(run.momentum()); nexts.append(run.hmc_next())
        run.close()
    out.update(hmc_traj=np.array(trajs), hmc_costs=np.array(costs), hmc_mom=np.array(moms),
               hmc_next=np.array(nexts), hmc_seeds=np.array([0, 7, 123456]))
    # derivative 2
    params = capi.default_params(n_points=50, lambda_=200.0, derivative=2)
    run = po.Run(robot, params, [sd], starts[2], goals[2], flavour=FL)
    ret, c, _, _ = run.iterate(30)
    assert ret == 0
    out.update(d2_traj=run.traj(), d2_costs=c)
    run.close()
    np.savez_compressed(os.path.join(OUT, "chomp.npz"), **out)


def modes():
    """floating base (mod.cpp:991-1021, 1050-1086, 2424-2464, 2805-2808) and a 200-sphere arm
    (BASELINE configs[4] shape, small), from the reference build"""
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    obs, sdf = po.computedistancefield(capi.make_prims(gp), len(gp), sizes, lengths, 0.02, flavour=FL)
    sd = capi.SdfDesc(sdf, lengths, models.pose_compose(kin_pose, gpose))
    out = {}
    # floating base, plain and momentum
    rng = np.random.default_rng(404)
    starts, goals = models.random_endpoints(robot, 2, seed0=404, shrink=0.3)
    base0 = np.asarray(robot.base_pose, dtype=float)
    qs, qg = [], []
    for r in range(2):
        move = models.pose_make(rng.uniform(-0.2, 0.2, 3), models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(0.2, 0.6)))
        qs.append(np.concatenate([base0, starts[r]]))
        qg.append(np.concatenate([models.pose_compose(base0, move), goals[r]]))
    qs, qg = np.array(qs), np.array(qg)
    for mom in (0, 1):
        params = capi.default_params(n_points=50, lambda_=100.0, obs_factor=300.0, floating_base=1, use_momentum=mom)
        trajs, costs, g0 = [], [], []
        for r in range(2):
            run = po.Run(robot, params, [sd], qs[r], qg[r], flavour=FL)
            ret, c, _, gr = run.iterate(30, want_grads=True)
            assert ret == 0
            trajs.append(run.traj()); costs.append(c); g0.append(gr[0])
            run.close()
        out.update({"float%d_traj" % mom: np.array(trajs), "float%d_costs" % mom: np.array(costs),
                    "float%d_grad0" % mom: np.array(g0)})
    out.update(float_starts=qs, float_goals=qg)
    # dense-sphere arm, fixed base
    robot5 = models.dense_sphere_arm(200, seed=5)
    params = capi.default_params(n_points=48, lambda_=200.0, obs_factor=100.0, epsilon=0.2)
    starts, goals = models.random_endpoints(robot5, 2, seed0=405, shrink=0.4)
    trajs, costs, g0 = [], [], []
    for r in range(2):
        run = po.Run(robot5, params, [sd], starts[r], goals[r], flavour=FL)
        ret, c, _, gr = run.iterate(6, want_grads=True)
        assert ret == 0
        trajs.append(run.traj()); costs.append(c); g0.append(gr[0])
        run.close()
    out.update(dense_starts=starts, dense_goals=goals, dense_traj=np.array(trajs), dense_costs=np.array(costs),
               dense_grad0=np.array(g0), table_sdf=sdf, table_lengths=np.array(lengths),
               table_pose=models.pose_compose(kin_pose, gpose))
    np.savez_compressed(os.path.join(OUT, "modes.npz"), **out)


def constraint_cases(po_mod, robot, flavour):
    """the TSR scenarios of constraints.npz: (name, params keywords, constraint list, starts, goals, iterations).
    Shared with the tests so fixture and checks cannot drift apart."""
    ee = robot.names.index("wam7")
    base = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    rng = np.random.default_rng(77)
    starts = np.repeat(base[None], 2, 0)
    goals = starts.copy()
    goals[:, 0] += rng.uniform(0.6, 1.3, 2)
    goals[:, 1:] += rng.uniform(-0.05, 0.05, (2, 6))
    tool = models.pose_make((0.0, 0.02, 0.12), models.quat_from_axis_angle((0, 1, 0), 0.3))
    pe = models.pose_compose(po_mod.fk(robot, base, flavour=flavour)[ee], tool)
    T0w, Twe = models.pose_make((0, 0, pe[2])), models.pose_make((0, 0, 0), pe[3:7])

    def bw(*held):
        B = np.tile(np.array([-10.0, 10.0]), (6, 1))
        for h in held:
            B[["x", "y", "z", "roll", "pitch", "yaw"].index(h)] = 0.0
        return B

    upright = capi.make_constraint("all", ee, bw("z", "roll", "pitch"), T0w=T0w, Twe=Twe, pose_link_ee=tool)
    elbow = robot.names.index("wam4")
    q1 = starts[0] + (goals[0] - starts[0]) / 39
    pin = capi.make_constraint("start", elbow, bw("x", "y"), T0w=po_mod.fk(robot, q1, flavour=flavour)[elbow])
    slide = capi.make_constraint("start_tsr", ee, bw("x", "z", "roll", "pitch", "yaw"),
                                 T0w=models.pose_make((pe[0], 0, pe[2])), Twe=Twe, pose_link_ee=tool)
    return [("all", dict(n_points=40, lambda_=100.0, obs_factor=500.0), [upright], starts, goals, 30),
            ("two_mom", dict(n_points=40, lambda_=200.0, obs_factor=300.0, use_momentum=1), [upright, pin],
             starts[:1], goals[:1], 30),
            ("start_tsr", dict(n_points=40, lambda_=150.0, obs_factor=500.0), [slide], starts, goals, 30),
            ("d2", dict(n_points=36, lambda_=400.0, derivative=2), [upright], starts[:1], goals[:1], 20)]


def constraints():
    """hard TSR constraints (mod.cpp:1330-1784; chomp.c:553-600) from the reference build: here the
    projection, the dense inverse and LAPACKE_dgesv are the reference's own chomp.c + OpenBLAS"""
    robot = models.wam7_robot()
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    obs, sdf = po.computedistancefield(capi.make_prims(gp), len(gp), sizes, lengths, 0.02, flavour=FL)
    sd = capi.SdfDesc(sdf, lengths, models.pose_compose(kin_pose, gpose))
    out = dict(table_sdf=sdf, table_lengths=np.array(lengths), table_pose=models.pose_compose(kin_pose, gpose))
    for name, kw, cons, starts, goals, n_iter in constraint_cases(po, robot, FL):
        params = capi.default_params(constraints=cons, **kw)
        trajs, costs, vals, jacs = [], [], [], []
        for r in range(len(starts)):
            run = po.Run(robot, params, [sd], starts[r], goals[r], flavour=FL)
            q = starts[r] + 0.37 * (goals[r] - starts[r])
            v, J = run.constraint_eval(0, q)
            ret, c, _, _ = run.iterate(n_iter)
            assert ret == 0
            trajs.append(run.traj()); costs.append(c); vals.append(v); jacs.append(J)
            run.close()
        out.update({name + "_traj": np.array(trajs), name + "_costs": np.array(costs), name + "_val": np.array(vals),
                    name + "_jac": np.array(jacs)})
    np.savez_compressed(os.path.join(OUT, "constraints.npz"), **out)


def mt():
    g = po.MT(0, flavour=FL)
    raw0 = np.array([g.next() for _ in range(1300)], dtype=np.uint64)
    g = po.MT(20260217, flavour=FL)
    gs = np.array([g.gaussian(0.1) for _ in range(200)])
    np.savez_compressed(os.path.join(OUT, "mt.npz"), raw_seed0=raw0, gauss_seed20260217=gs)


if __name__ == "__main__":
    assert po.available("reference"), "build oracle/_ref first (needs /root/reference)"
    sdf_kat(); sdf_build(); occupancy(); mesh(); chomp(); modes(); constraints(); mt()
    print("golden fixtures written to", OUT)
