"""GPU parity tests of the hard end-effector constraints (task space regions): `create ... con_tsr /
everyn_tsr / start_tsr` (src/orcdchomp_mod.cpp:1330-1784, 1930-1996, 2466-2613) and the projection
in cd_chomp_iterate (src/libcd/chomp.c:553-600), through the C ABI, against the CPU oracle -- which for
this part is the reference's own chomp.c / kin.c / spatial.c when oracle/_ref is present."""
import numpy as np
import pytest

from or_cdchomp_b200 import capi, models

pytestmark = pytest.mark.gpu

TRAJ_ATOL = 1e-6
COST_RTOL = 1e-8


def upright_scene(oracle, robot, n_runs, seed=7):
    """Start / goal pairs that differ mostly in the first (vertical) joint, and the TSR that describes
    'keep the tool at the start height with its start roll and pitch': T0w at the start height, Twe the
    tool's start orientation, so the constrained entries are zero at the start and stay close to zero
    along the straight line."""
    rng = np.random.default_rng(seed)
    ee = robot.names.index("wam7")
    base = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    starts = np.repeat(base[None], n_runs, 0)
    goals = starts.copy()
    goals[:, 0] += rng.uniform(0.6, 1.3, n_runs)
    goals[:, 1:] += rng.uniform(-0.05, 0.05, (n_runs, 6))
    pe = oracle.fk(robot, base)[ee]
    T0w = models.pose_make((0, 0, pe[2]))
    Twe = models.pose_make((0, 0, 0), pe[3:7])
    return ee, starts, goals, T0w, Twe


def run_oracle(oracle, flavour, robot, params, sd, starts, goals, n_iter, seeds=None):
    out = []
    for r in range(len(starts)):
        run = oracle.Run(robot, params, [sd], starts[r], goals[r], seed=0 if seeds is None else int(seeds[r]),
                         flavour=flavour)
        ret, c, tr, _ = run.iterate(n_iter, want_trace=True)
        out.append(dict(ret=ret, costs=c, trace=tr, traj=run.traj(), run=run))
    return out


def bounds(*held):
    """Bw with the named rows (x y z roll pitch yaw) held at zero and the rest free"""
    names = ["x", "y", "z", "roll", "pitch", "yaw"]
    Bw = np.tile(np.array([-10.0, 10.0]), (6, 1))
    for h in held:
        Bw[names.index(h)] = 0.0
    return Bw


@pytest.fixture(scope="module")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from or_cdchomp_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("where,held", [("all", ("z", "roll", "pitch")), ("all", ("roll", "pitch")),
                                        ("start", ("z", "roll", "pitch")), ("end", ("x", "y", "z", "roll", "pitch", "yaw"))])
def test_tsr_constraint_matches_oracle(engine, oracle, flavour, wam7, table, where, held):
    """con_tsr 'all' / 'start' / 'end' with different held rows, 40 iterations"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 4)
    if where == "end":
        # hold the whole pose of the last moving waypoint where the straight line puts it
        P = 40
        q = starts[0] + (goals[0] - starts[0]) * (P - 2) / (P - 1)
        starts, goals = starts[:1], goals[:1]
        T0w = oracle.fk(wam7, q)[ee]
        Twe = models.pose_make()
    cons = [capi.make_constraint(where, ee, bounds(*held), T0w=T0w, Twe=Twe)]
    params = capi.default_params(n_points=40, lambda_=100.0, obs_factor=500.0, constraints=cons)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    assert not b.uses_jit()
    b.enable_trace(True)
    costs, status = b.iterate(40)
    traj, trace = b.get_traj(), b.get_trace(40)
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 40)
    for r, o in enumerate(ref):
        assert o["ret"] == 0 and status[r] == 0
        assert np.max(np.abs(traj[r] - o["traj"])) <= TRAJ_ATOL
        assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
        assert np.allclose(trace[r], o["trace"], rtol=1e-7, atol=0)
        # the constraint holds on the GPU's own result (property, not a comparison)
        idx = {"all": range(1, 39), "start": [1], "end": [38]}[where]
        h = np.array([o["run"].constraint_eval(0, traj[r][i])[0] for i in idx])
        h_ref = np.array([o["run"].constraint_eval(0, o["traj"][i])[0] for i in idx])
        assert np.max(np.abs(h - h_ref)) < 1e-5
        if where == "all":  # every point projected in every iteration: the linearisation error is all that is left
            assert np.max(np.abs(h)) < 1e-4
    # and it did something: the unconstrained run ends elsewhere
    p0 = capi.default_params(n_points=40, lambda_=100.0, obs_factor=500.0)
    b0 = engine.create_batch(wam7, p0, [sid], starts, goals)
    b0.iterate(40)
    if where != "end":
        assert np.max(np.abs(b0.get_traj() - traj)) > 1e-3
    b0.close()
    b.close()
    engine.remove_sdf(sid)


def test_two_constraints_tool_offset_momentum(engine, oracle, flavour, wam7, table):
    """everyn_tsr-style constraint on every point through a tool transform, plus a con_tsr on a different link
    at the first point; momentum update (the constraint then sees the accumulated AG, chomp.c:562-565)"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 3, seed=11)
    tool = models.pose_make((0.0, 0.02, 0.12), models.quat_from_axis_angle((0, 1, 0), 0.3))
    pe = models.pose_compose(oracle.fk(wam7, starts[0])[ee], tool)
    T0w = models.pose_make((0, 0, pe[2]))
    Twe = models.pose_make((0, 0, 0), pe[3:7])
    elbow = wam7.names.index("wam4")
    q1 = starts[0] + (goals[0] - starts[0]) / 29
    cons = [capi.make_constraint("all", ee, bounds("z", "pitch"), T0w=T0w, Twe=Twe, pose_link_ee=tool),
            capi.make_constraint("start", elbow, bounds("x"), T0w=oracle.fk(wam7, q1)[elbow])]
    starts, goals = starts[:1].repeat(3, 0), goals[:1].repeat(3, 0)
    goals[1, 3] += 0.02
    goals[2, 5] -= 0.03
    params = capi.default_params(n_points=30, lambda_=200.0, obs_factor=300.0, use_momentum=1, constraints=cons)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    costs, status = b.iterate(30)
    traj = b.get_traj()
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 30)
    for r, o in enumerate(ref):
        assert o["ret"] == 0 and status[r] == 0
        assert np.max(np.abs(traj[r] - o["traj"])) <= TRAJ_ATOL
        assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
    b.close()
    engine.remove_sdf(sid)


def test_constraint_with_wider_metric(engine, oracle, flavour, wam7, table):
    """derivative = 2: the dense inverse of a penta-diagonal metric couples the constrained waypoints"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 2, seed=5)
    cons = [capi.make_constraint("all", ee, bounds("roll", "pitch"), T0w=T0w, Twe=Twe)]
    params = capi.default_params(n_points=36, lambda_=400.0, derivative=2, constraints=cons)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    costs, status = b.iterate(25)
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 25)
    for r, o in enumerate(ref):
        assert o["ret"] == 0 and status[r] == 0
        assert np.max(np.abs(b.get_traj()[r] - o["traj"])) <= TRAJ_ATOL
        assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
    b.close()
    engine.remove_sdf(sid)


def test_start_tsr_frees_the_start_point(engine, oracle, flavour, wam7, table):
    """start_tsr (mod.cpp:2316, 2320-2323, 2521, 2571-2576; sphere_cost_pre 1040-1043, 1108-1114, 1126-1128):
    m = n_points - 1, the start row is optimised under its constraint, no initial boundary row in the metric,
    one-sided velocity and borrowed acceleration at the start; plain and momentum + HMC updates"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 3)
    sid = engine.upload_sdf(table["desc"])
    for kw, seeds in ((dict(), None), (dict(use_momentum=1, use_hmc=1, hmc_resample_lambda=0.1), [3, 4, 5])):
        cons = [capi.make_constraint("start_tsr", ee, bounds("x", "z", "roll", "pitch", "yaw"), T0w=T0w, Twe=Twe)]
        pe = oracle.fk(wam7, starts[0])[ee]
        cons[0].T0w[0] = pe[0]  # the start may slide along y only
        params = capi.default_params(n_points=40, lambda_=150.0, obs_factor=500.0, constraints=cons, **kw)
        b = engine.create_batch(wam7, params, [sid], starts, goals, seeds=seeds)
        assert b.m == 39
        b.enable_trace(True)
        b.capture_gradient(1)
        costs, status = b.iterate(30)
        traj, trace, grad = b.get_traj(), b.get_trace(30), b.get_gradient()
        assert grad.shape == (3, 39, 7)
        ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 30, seeds=seeds)
        for r, o in enumerate(ref):
            assert o["ret"] == 0 and status[r] == 0
            assert np.max(np.abs(traj[r] - o["traj"])) <= TRAJ_ATOL
            assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
            assert np.allclose(trace[r], o["trace"], rtol=1e-7, atol=0)
            assert np.max(np.abs(traj[r][0] - starts[r])) > 1e-3      # the start moved ...
            h = o["run"].constraint_eval(0, traj[r][0])[0]
            assert np.max(np.abs(h)) < 5e-3                            # ... inside its region (to first order)
            assert np.array_equal(traj[r][-1], goals[r])               # the goal never does
        b.close()
    engine.remove_sdf(sid)


def test_floating_base_constraint(engine, oracle, flavour, wam7, table):
    """floating_base with a constraint on every point: the pose columns of the constraint Jacobian come from
    cd_spatial_pose_jac (mod.cpp:1432-1437), unscaled"""
    rng = np.random.default_rng(21)
    ee = wam7.names.index("wam7")
    base0 = np.asarray(wam7.base_pose, dtype=float)
    arm = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    qs, qg = [], []
    for r in range(3):
        move = models.pose_make(rng.uniform(-0.15, 0.15, 3), models.quat_from_axis_angle((0, 0, 1), rng.uniform(0.2, 0.5)))
        qs.append(np.concatenate([base0, arm]))
        qg.append(np.concatenate([models.pose_compose(base0, move), arm + rng.uniform(-0.1, 0.1, 7)]))
    qs, qg = np.array(qs), np.array(qg)
    pe = oracle.fk(wam7, arm)[ee]
    cons = [capi.make_constraint("all", ee, bounds("z", "pitch"), T0w=models.pose_make((0, 0, pe[2])),
                                 Twe=models.pose_make((0, 0, 0), pe[3:7]))]
    params = capi.default_params(n_points=36, lambda_=200.0, obs_factor=300.0, floating_base=1, constraints=cons)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], qs, qg)
    costs, status = b.iterate(25)
    traj = b.get_traj()
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], qs, qg, 25)
    for r, o in enumerate(ref):
        assert o["ret"] == 0 and status[r] == 0
        assert np.max(np.abs(traj[r] - o["traj"])) <= TRAJ_ATOL
        assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
        assert np.allclose(np.linalg.norm(traj[r][:, 3:7], axis=1), 1.0, atol=1e-14)
    b.close()
    engine.remove_sdf(sid)


def test_constraint_argument_errors(engine, wam7, table):
    """what create refuses (mod.cpp:2100; the tiled path and the run-time specialised kernel do not carry constraints)"""
    ee = wam7.names.index("wam7")
    sid = engine.upload_sdf(table["desc"])
    qs, qg = np.array(models.WAM7_DEMO_START), np.array(models.WAM7_DEMO_GOAL)
    two = [capi.make_constraint("start_tsr", ee, bounds("z")), capi.make_constraint("start_tsr", ee, bounds("x"))]
    with pytest.raises(RuntimeError, match="at most one start_tsr"):
        engine.create_batch(wam7, capi.default_params(n_points=20, constraints=two), [sid], qs, qg)
    with pytest.raises(RuntimeError, match="floating_base and start_tsr"):
        engine.create_batch(wam7, capi.default_params(n_points=20, floating_base=1, constraints=two[:1]), [sid],
                            np.concatenate([wam7.base_pose, qs]), np.concatenate([wam7.base_pose, qg]))
    bad = [capi.make_constraint("all", 99, bounds("z"))]
    with pytest.raises(RuntimeError, match="bad link"):
        engine.create_batch(wam7, capi.default_params(n_points=20, constraints=bad), [sid], qs, qg)
    # a constraint whose bounds hold nothing adds no rows: same result as no constraint
    free = [capi.make_constraint("all", ee, bounds())]
    b1 = engine.create_batch(wam7, capi.default_params(n_points=30, lambda_=100.0, constraints=free), [sid], qs, qg)
    b2 = engine.create_batch(wam7, capi.default_params(n_points=30, lambda_=100.0), [sid], qs, qg)
    b1.iterate(10)
    b2.iterate(10)
    assert np.max(np.abs(b1.get_traj() - b2.get_traj())) < 1e-12
    b1.close()
    b2.close()
    engine.remove_sdf(sid)


def test_golden_constraints(engine, oracle, wam7):
    """against outputs of the reference's own chomp.c / kin.c / spatial.c / LAPACK build
    (tests/golden/constraints.npz, generated by tests/golden/make_golden.py)"""
    import importlib.util
    from conftest import golden_path
    spec = importlib.util.spec_from_file_location("make_golden", golden_path("make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(golden_path("constraints.npz"))
    sid = engine.upload_sdf(capi.SdfDesc(gold["table_sdf"], gold["table_lengths"], gold["table_pose"]))
    for name, kw, cons, starts, goals, n_iter in mg.constraint_cases(oracle, wam7, "port"):
        params = capi.default_params(constraints=cons, **kw)
        b = engine.create_batch(wam7, params, [sid], starts, goals)
        costs, status = b.iterate(n_iter)
        assert (status == 0).all()
        assert np.max(np.abs(b.get_traj() - gold[name + "_traj"])) <= TRAJ_ATOL
        assert np.allclose(costs, gold[name + "_costs"], rtol=COST_RTOL, atol=0)
        b.close()
    engine.remove_sdf(sid)


def test_cd_chomp_facade_with_constraints(oracle, flavour, wam7, table):
    """libcd_b200's cd_chomp facade: constraints travel as data in the params of cd_chomp_b200_set_sphere_cost"""
    from or_cdchomp_b200 import libcd
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 1)
    cons = [capi.make_constraint("all", ee, bounds("z", "roll", "pitch"), T0w=T0w, Twe=Twe)]
    params = capi.default_params(n_points=32, lambda_=150.0, obs_factor=400.0, constraints=cons)
    run = oracle.Run(wam7, params, [table["desc"]], starts[0], goals[0], flavour=flavour)
    fac = libcd.ChompRun(wam7, params, [table["desc"]], starts[0], goals[0])
    for it in range(10):
        ret, c, tr, _ = run.iterate(1, want_trace=True)
        rc, cf = fac.iterate(1)
        assert rc == ret == 0
        assert np.allclose(cf, tr[0], rtol=1e-8, atol=0)
        assert np.max(np.abs(fac.traj - run.traj())) <= TRAJ_ATOL
    fac.close()
    run.close()
    bad = capi.default_params(n_points=32, constraints=[capi.make_constraint("start_tsr", ee, bounds("z"))])
    with pytest.raises(RuntimeError, match="start_tsr is not offered"):
        libcd.ChompRun(wam7, bad, [table["desc"]], starts[0], goals[0])


def test_dependent_rows_are_skipped_and_counted(engine, oracle, wam7, table):
    """the same constraint twice: every waypoint's rows are linearly dependent.  The reference's dgesv then fails
    ("constraint inversion error!", chomp.c:582-590) and it moves on with an unsolved right-hand side; the engine
    (tridiagonal metric) leaves those waypoints unconstrained, counts them, and stays finite"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 2)
    c = capi.make_constraint("all", ee, bounds("z", "pitch"), T0w=T0w, Twe=Twe)
    params = capi.default_params(n_points=30, lambda_=100.0, obs_factor=500.0, constraints=[c, c])
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    costs, status = b.iterate(5)
    assert (status == 0).all() and np.isfinite(b.get_traj()).all() and np.isfinite(costs).all()
    assert (b.get_constraint_skips() == 5 * 28).all()
    b.close()
    # the same constraint once: nothing skipped
    params = capi.default_params(n_points=30, lambda_=100.0, obs_factor=500.0, constraints=[c])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    b.iterate(5)
    assert (b.get_constraint_skips() == 0).all()
    b.close()
    engine.remove_sdf(sid)


@pytest.mark.parametrize("n_points", [4, 5, 32, 33, 66])
def test_short_trajectories_both_projection_forms(engine, oracle, flavour, wam7, table, n_points, monkeypatch):
    """block sizes of one and two warps, one to three moving waypoints: the two-sided sweep degenerates to a
    one-sided one / to the middle step alone; both forms (OCB_CON_DENSE=0 sweep, =1 dense system) against the oracle"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 2, seed=n_points)
    cons = [capi.make_constraint("all", ee, bounds("z", "roll", "pitch"), T0w=T0w, Twe=Twe)]
    params = capi.default_params(n_points=n_points, lambda_=100.0, obs_factor=500.0, constraints=cons)
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 10)
    sid = engine.upload_sdf(table["desc"])
    for form in ("0", "1"):
        monkeypatch.setenv("OCB_CON_DENSE", form)
        b = engine.create_batch(wam7, params, [sid], starts, goals)
        costs, status = b.iterate(10)
        for r, o in enumerate(ref):
            assert o["ret"] == 0 and status[r] == 0
            assert np.max(np.abs(b.get_traj()[r] - o["traj"])) <= TRAJ_ATOL, form
            assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0), form
        b.close()
    engine.remove_sdf(sid)


def test_long_trajectory_sweep_in_global_scratch(engine, oracle, flavour, wam7, table):
    """n_points = 256 (cfg4's shape, 8 warps per run), momentum, every waypoint constrained: 762 rows; the sweep's
    matrices no longer fit the run's shared workspace and live in global scratch"""
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 2, seed=256)
    cons = [capi.make_constraint("all", ee, bounds("z", "roll", "pitch"), T0w=T0w, Twe=Twe)]
    params = capi.default_params(n_points=256, lambda_=100.0, obs_factor=500.0, use_momentum=1, constraints=cons)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    costs, status = b.iterate(8)
    ref = run_oracle(oracle, flavour, wam7, params, table["desc"], starts, goals, 8)
    for r, o in enumerate(ref):
        assert o["ret"] == 0 and status[r] == 0
        assert np.max(np.abs(b.get_traj()[r] - o["traj"])) <= TRAJ_ATOL
        assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0)
    assert (b.get_constraint_skips() == 0).all()
    b.close()
    engine.remove_sdf(sid)


def test_constraints_through_the_multi_engine_interface(oracle, wam7, table):
    """ocb_multi_*: a constrained batch dealt over several engines in one process gives, in global run order,
    exactly what one engine gives (devices 0 and 1 when the box has them, else two engines on device 0)"""
    import torch
    from or_cdchomp_b200.engine import Engine, MultiEngine
    ee, starts, goals, T0w, Twe = upright_scene(oracle, wam7, 9, seed=3)
    cons = [capi.make_constraint("all", ee, bounds("z", "roll", "pitch"), T0w=T0w, Twe=Twe),
            capi.make_constraint("start_tsr", ee, bounds("x", "z", "roll", "pitch", "yaw"),
                                 T0w=models.pose_make((oracle.fk(wam7, starts[0])[ee][0], 0, T0w[2])), Twe=Twe)]
    params = capi.default_params(n_points=40, lambda_=150.0, obs_factor=500.0, constraints=cons)
    e = Engine(0)
    sid = e.upload_sdf(table["desc"])
    b = e.create_batch(wam7, params, [sid], starts, goals)
    c1, s1 = b.iterate(15)
    t1 = b.get_traj()
    b.close()
    e.close()
    assert (s1 == 0).all() and np.max(np.abs(t1[:, 0] - starts)) > 1e-4
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    m = MultiEngine(devs)
    msid = m.upload_sdf(table["desc"])
    mb = m.create_batch(wam7, params, [msid], starts, goals)
    c2, s2 = mb.iterate(15)
    assert np.array_equal(s1, s2) and np.array_equal(c1, c2) and np.array_equal(mb.get_traj(), t1)
    mb.close()
    m.remove_sdf(msid)
    m.close()


def test_tree_robot_prismatic_mimic_branch(engine, oracle, flavour, table):
    """a constraint on each branch of a tree with a prismatic joint and a mimic joint: the frames come through a
    reloaded branch frame, the columns include a prismatic one (zero angular part) and a dof that two joints share"""
    robot = models.prismatic_test_robot()
    sd = capi.SdfDesc(table["sdf"], table["lengths"], models.pose_make((-0.5, -0.6, 0.1)))
    start = np.array([0.1, 0.3, 0.2, -0.4])
    goals = start[None] + np.array([[0.12, 0.25, 0.1, 0.15], [-0.1, -0.2, 0.15, 0.2], [0.1, 0.2, -0.12, -0.25]])
    starts = np.repeat(start[None], 3, 0)
    tip, tool = robot.names.index("armA2"), robot.names.index("tool")
    fk0 = oracle.fk(robot, start)
    offset = models.pose_make((0.0, 0.05, 0.0), models.quat_from_axis_angle((1, 0, 0), 0.2))
    cons = [capi.make_constraint("all", tip, bounds("z"), T0w=fk0[tip]),
            capi.make_constraint("start", tool, bounds("x"), T0w=models.pose_compose(fk0[tool], offset), pose_link_ee=offset)]
    params = capi.default_params(n_points=33, lambda_=200.0, obs_factor=300.0, epsilon=0.15, constraints=cons)
    sid = engine.upload_sdf(sd)
    for env in ("0", "1"):
        import os
        os.environ["OCB_CON_DENSE"] = env
        try:
            b = engine.create_batch(robot, params, [sid], starts, goals)
        finally:
            del os.environ["OCB_CON_DENSE"]
        costs, status = b.iterate(25)
        ref = run_oracle(oracle, flavour, robot, params, sd, starts, goals, 25)
        for r, o in enumerate(ref):
            assert (o["ret"] == 0) == (status[r] == 0)
            if o["ret"] == 0:
                assert np.max(np.abs(b.get_traj()[r] - o["traj"])) <= TRAJ_ATOL, env
                assert np.allclose(costs[r], o["costs"], rtol=COST_RTOL, atol=0), env
        assert any(o["ret"] == 0 for o in ref)
        b.close()
    engine.remove_sdf(sid)


def test_constraints_on_the_tiled_path(engine, oracle, flavour, table):
    """a 200-sphere arm does not fit the persistent kernel: the constrained update runs in the tiled path's
    update kernel, with the branch frames, the rows and the sweep's matrices in global scratch; both projection
    forms, plain and momentum, against the oracle"""
    robot = models.dense_sphere_arm(200, seed=5)
    ee = robot.names.index("wam7")
    base = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    starts = np.repeat(base[None], 2, 0)
    goals = starts + np.array([[0.5, 0.03, -0.02, 0.04, 0.0, -0.03, 0.02], [0.7, -0.04, 0.03, -0.02, 0.03, 0.02, -0.04]])
    pe = oracle.fk(robot, base)[ee]
    cons = [capi.make_constraint("all", ee, bounds("z", "roll", "pitch"), T0w=models.pose_make((0, 0, pe[2])),
                                 Twe=models.pose_make((0, 0, 0), pe[3:7]))]
    sid = engine.upload_sdf(table["desc"])
    import os
    for kw in (dict(), dict(use_momentum=1)):
        params = capi.default_params(n_points=48, lambda_=300.0, obs_factor=100.0, epsilon=0.2, constraints=cons, **kw)
        ref = run_oracle(oracle, flavour, robot, params, table["desc"], starts, goals, 5)
        for form in ("0", "1"):
            os.environ["OCB_CON_DENSE"] = form
            try:
                b = engine.create_batch(robot, params, [sid], starts, goals)
            finally:
                del os.environ["OCB_CON_DENSE"]
            assert b.tile_width() > 0 and not b.uses_jit()
            costs, status = b.iterate(5)
            for r, o in enumerate(ref):
                assert o["ret"] == 0 and status[r] == 0
                assert np.max(np.abs(b.get_traj()[r] - o["traj"])) <= TRAJ_ATOL, (kw, form)
                assert np.allclose(costs[r], o["costs"], rtol=1e-7, atol=0), (kw, form)
            b.close()
    # start_tsr stays with the persistent kernel
    bad = capi.default_params(n_points=48, constraints=[capi.make_constraint("start_tsr", ee, bounds("z"))])
    with pytest.raises(RuntimeError, match="start_tsr needs the run in one SM"):
        engine.create_batch(robot, bad, [sid], starts, goals)
    engine.remove_sdf(sid)
