"""The reference's module commands end to end on the GPU (scripts/test_wam7.py flow):
computedistancefield -> runchomp = create / iterate / gettraj / destroy, plus error behaviour."""
import ctypes as C
import os

import numpy as np
import pytest

from or_cdchomp_b200 import capi, models, orcdchomp

pytestmark = pytest.mark.gpu


@pytest.fixture()
def world(engine):
    env = orcdchomp.Environment()
    kin_pose, prims, apos, aext = models.table_scene()
    table = env.AddKinBody("table", kin_pose, prims)
    robot_desc = models.wam7_robot()
    robot = env.AddRobot("BarrettWAM", robot_desc, models.WAM7_DEMO_START)
    mod = orcdchomp.Module(env, 0)
    yield env, mod, table, robot, robot_desc
    mod.close()
    env.close()


def oracle_field_for_table(oracle, flavour, cube_extent=0.02, padding=0.2):
    """what computedistancefield(kinbody=table) must produce, via the oracle"""
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, cube_extent, padding)
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    obs, sdf = oracle.computedistancefield(pa, len(gp), sizes, lengths, cube_extent, flavour=flavour)
    return capi.SdfDesc(sdf, lengths, models.pose_compose(kin_pose, gpose))


def test_runchomp_matches_oracle(world, oracle, flavour):
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    cost = [None]
    traj = mod.runchomp(robot=robot, n_iter=100, lambda_=100.0, obs_factor=500.0, n_points=100,
                        adofgoal=list(models.WAM7_DEMO_GOAL), no_collision_exception=True, cost=cost)
    sd = oracle_field_for_table(oracle, flavour)
    params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
    run = oracle.Run(robot_desc, params, [sd], models.WAM7_DEMO_START, models.WAM7_DEMO_GOAL, flavour=flavour)
    ret, c, _, _ = run.iterate(100)
    assert ret == 0 and traj.shape == (100, 7)
    assert np.max(np.abs(traj - run.traj())) <= 1e-6
    assert abs(cost[0] - c[0]) <= 1e-5 * abs(c[0])  # iterate prints 6 significant digits (mod.cpp:2849)
    run.close()
    # default n_points is 101 (mod.cpp:1840)
    h = mod.create(robot=robot, adofgoal=list(models.WAM7_DEMO_GOAL))
    assert mod.gettraj(run=h).shape == (101, 7)
    mod.destroy(run=h)


def test_field_attached_to_disabled_robot(world, oracle, flavour):
    """scripts/test_wam7.py:76-80: the robot is disabled and the field for everything else is
    attached to the robot body."""
    env, mod, table, robot, robot_desc = world
    env.Enable(robot, False)
    mod.computedistancefield(kinbody=robot, cube_extent=0.04)
    env.Enable(robot, True)
    assert "BarrettWAM" in mod.viewfields()
    traj = mod.runchomp(robot=robot, n_iter=5, lambda_=100.0, adofgoal=list(models.WAM7_DEMO_GOAL))
    assert np.isfinite(traj).all()
    mod.removefield(kinbody=robot)
    with pytest.raises(RuntimeError, match="No signed distance fields"):
        mod.create(robot=robot, adofgoal=list(models.WAM7_DEMO_GOAL))


def test_addfield_fromobsarray_cache_and_errors(world, oracle, flavour, tmp_path):
    env, mod, table, robot, robot_desc = world
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    rng = np.random.default_rng(0)
    obs = np.where(rng.uniform(size=(12, 10, 8)) < 0.1, np.inf, 0.0)
    p = libc.malloc(obs.nbytes)                       # the module takes ownership and frees it (mod.cpp:703-704)
    C.memmove(p, obs.ctypes.data, obs.nbytes)
    box = env.AddKinBody("box", models.pose_make((0.5, 0.0, 0.3)), [])
    mod.addfield_fromobsarray(kinbody=box, obsarray="0x%x" % p, sizes=obs.shape, lengths=[0.6, 0.5, 0.4],
                              pose=[-0.3, -0.25, -0.2, 0, 0, 0, 2.0])   # quaternion gets normalised
    with pytest.raises(RuntimeError, match="already have an sdf"):
        mod.addfield_fromobsarray(kinbody=box, obsarray="0x%x" % p, sizes=obs.shape, lengths=[0.6, 0.5, 0.4])
    # the same run through the oracle
    sdf = oracle.sdf_from_obsarray(obs, [0.6, 0.5, 0.4], flavour=flavour)
    pose_world = models.pose_compose(models.pose_make((0.5, 0.0, 0.3)), models.pose_make((-0.3, -0.25, -0.2)))
    sd = capi.SdfDesc(sdf, [0.6, 0.5, 0.4], pose_world)
    params = capi.default_params(n_points=40, lambda_=50.0)
    goal = [0.3, 0.8, 0.1, 1.2, 0.0, 0.3, 0.2]
    traj = mod.runchomp(robot=robot, n_iter=30, lambda_=50.0, n_points=40, adofgoal=goal)
    run = oracle.Run(robot_desc, params, [sd], models.WAM7_DEMO_START, goal, flavour=flavour)
    run.iterate(30)
    assert np.max(np.abs(traj - run.traj())) <= 1e-6
    run.close()
    # cache file: raw doubles, written then read back (mod.cpp:416-444, 571-580)
    cache = str(tmp_path / "sdf_table.dat")
    mod.computedistancefield(kinbody=table, cube_extent=0.04, cache_filename=cache)
    assert os.path.getsize(cache) % 8 == 0
    mod.removefield(kinbody=table)
    mod.computedistancefield(kinbody=table, cube_extent=0.04, cache_filename=cache, require_cache=True)
    mod.removefield(kinbody=table)
    with pytest.raises(RuntimeError, match="require_cache"):
        mod.computedistancefield(kinbody=table, cube_extent=0.03, cache_filename=cache, require_cache=True)
    # error texts of the reference
    with pytest.raises(RuntimeError, match="Bad arguments!"):
        mod.SendCommand("create robot BarrettWAM adofgoal '0 0 0 0 0 0 0' no_report_cost")
    with pytest.raises(RuntimeError, match="lambda must be >=0.01!"):
        mod.create(robot=robot, adofgoal=goal, lambda_=0.001)
    with pytest.raises(RuntimeError, match="n_points must be >=3!"):
        mod.create(robot=robot, adofgoal=goal, n_points=2)
    with pytest.raises(RuntimeError, match="size of adofgoal does not match"):
        mod.create(robot=robot, adofgoal=[0, 1])
    with pytest.raises(RuntimeError, match="Cannot parse start_tsr TSR!"):
        mod.create(robot=robot, adofgoal=goal, start_tsr="x")
    with pytest.raises(RuntimeError, match="not supported by the B200 engine"):
        mod.create(robot=robot, adofgoal=goal, start_cost="0x1 0x2")
    with pytest.raises(RuntimeError, match="you must pass a created run!"):
        mod.iterate(run="0x1234", n_iter=1)
    with pytest.raises(RuntimeError, match="Could not find kinbody"):
        mod.computedistancefield(kinbody="nothing")


def test_createbatch_and_dat_file(world, tmp_path):
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    starts, goals = models.random_endpoints(robot_desc, 16, shrink=0.3)
    h = mod.createbatch(robot=robot, adofgoals=goals, adofstarts=starts, lambda_=100.0, n_points=100,
                        obs_factor=500.0)
    costs = [float(x) for x in mod.iterate(run=h, n_iter=20).split()]
    trajs = mod.gettraj(run=h)
    assert len(costs) == 16 and trajs.shape == (16, 100, 7)
    assert np.array_equal(trajs[:, 0], starts)
    mod.destroy(run=h)
    dat = str(tmp_path / "run.dat")
    h = mod.create(robot=robot, adofgoal=list(goals[0]), lambda_=100.0, dat_filename=dat)
    mod.iterate(run=h, n_iter=7)
    mod.destroy(run=h)
    rows = [l.split() for l in open(dat)]
    assert len(rows) == 7 and [int(r[0]) for r in rows] == list(range(7)) and all(len(r) == 5 for r in rows)


def test_starttraj_seeding(world):
    """create starttraj=<trajectory XML> (mod.cpp:2005-2012, 2375-2415): the run starts from the
    passed trajectory, sampled at n_points evenly spaced times over its duration."""
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    goal = list(models.WAM7_DEMO_GOAL)
    h = mod.create(robot=robot, adofgoal=goal, n_points=60, lambda_=100.0, obs_factor=300.0)
    mod.iterate(run=h, n_iter=7)
    xml = mod.SendCommand("gettraj run %s no_collision_check" % h)
    t7 = mod.gettraj(run=h, no_collision_check=True)
    # same number of points: the seed is reproduced and the optimisation continues identically
    h2 = mod.create(robot=robot, starttraj=xml, n_points=60, lambda_=100.0, obs_factor=300.0)
    assert np.max(np.abs(mod.gettraj(run=h2, no_collision_check=True) - t7)) <= 1e-14
    mod.iterate(run=h, n_iter=5)
    mod.iterate(run=h2, n_iter=5)
    a, b = mod.gettraj(run=h, no_collision_check=True), mod.gettraj(run=h2, no_collision_check=True)
    assert np.max(np.abs(a - b)) <= 1e-11 and np.max(np.abs(a - t7)) > 1e-4
    mod.destroy(run=h2)
    # resampling to another length: linear interpolation in time, same end points
    h3 = mod.create(robot=robot, starttraj=xml, n_points=119)
    t = mod.gettraj(run=h3, no_collision_check=True)
    assert t.shape == (119, 7)
    assert np.max(np.abs(t[0] - t7[0])) == 0.0 and np.max(np.abs(t[-1] - t7[-1])) <= 1e-15
    assert np.max(np.abs(t[::2] - t7)) <= 1e-13   # 118 = 2 * 59: every second sample is a seed waypoint
    assert np.max(np.abs(t[1::2] - 0.5 * (t7[:-1] + t7[1:]))) <= 1e-13
    mod.destroy(run=h3)
    mod.destroy(run=h)
    for bad, text in ((dict(starttraj=xml, adofgoal=goal), "Cannot pass both adofgoal and starttraj!"),
                      (dict(starttraj="<trajectory></trajectory>"), "joint_values"),
                      (dict(starttraj=xml.replace('dof="7" interpolation="linear"', 'dof="6" interpolation="linear"')),
                       "does not match")):
        with pytest.raises(RuntimeError) as ei:
            mod.create(robot=robot, **bad)
        assert text in str(ei.value)


def test_floating_base_through_module(world, oracle, flavour):
    """create floating_base basegoal '...' (mod.cpp:1912-1923, 2093, 2424-2443): rows are the base
    pose from the robot's transform to basegoal followed by the active dofs; gettraj carries the
    affine_transform group (mod.cpp:2912-2956)."""
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    base0 = np.asarray(robot_desc.base_pose, dtype=float)
    base1 = models.pose_compose(base0, models.pose_make((0.1, 0.15, -0.05), models.quat_from_axis_angle((0, 0.3, 1), 0.5)))
    goal = list(models.WAM7_DEMO_GOAL)
    traj = mod.runchomp(robot=robot, n_iter=20, lambda_=100.0, obs_factor=300.0, n_points=40, adofgoal=goal,
                        floating_base=True, basegoal=list(base1), no_collision_check=True)
    assert traj.shape == (40, 14)
    sd = oracle_field_for_table(oracle, flavour)
    params = capi.default_params(n_points=40, lambda_=100.0, obs_factor=300.0, floating_base=1)
    run = oracle.Run(robot_desc, params, [sd], np.concatenate([base0, models.WAM7_DEMO_START]),
                     np.concatenate([base1, goal]), flavour=flavour)
    ret, c, _, _ = run.iterate(20)
    assert ret == 0 and np.max(np.abs(traj - run.traj())) <= 1e-6
    run.close()
    for bad, text in ((dict(floating_base=True), "Passed floating_base with no basegoal!"),
                      (dict(floating_base=True, basegoal=[0, 0, 0, 1]), "basegoal argument must be length 7!")):
        with pytest.raises(RuntimeError) as ei:
            mod.create(robot=robot, adofgoal=goal, **bad)
        assert text in str(ei.value)


def test_trajs_fileformstr_dumps(world, tmp_path):
    """iterate trajs_fileformstr FMT (mod.cpp:2769-2795): the trajectory as it stands before every
    iteration is written to sprintf(FMT, iter); stepping this way changes nothing in the result."""
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    goal = list(models.WAM7_DEMO_GOAL)
    kw = dict(robot=robot, adofgoal=goal, n_points=30, lambda_=100.0, obs_factor=300.0)
    h = mod.create(**kw)
    fmt = str(tmp_path / "traj_%03d.xml")
    mod.iterate(run=h, n_iter=4, trajs_fileformstr=fmt)
    final = mod.gettraj(run=h, no_collision_check=True)
    files = sorted(os.listdir(tmp_path))
    assert files == ["traj_%03d.xml" % k for k in range(4)]
    import re
    dumps = []
    for f in files:
        text = open(tmp_path / f).read()
        assert 'count="30"' in text and "joint_values BarrettWAM" in text
        body = re.search(r"<data[^>]*>\s*(.*?)\s*</data>", text, re.S).group(1)
        dumps.append(np.array(body.split(), dtype=np.float64).reshape(30, 7))
    h2 = mod.create(**kw)
    assert np.array_equal(dumps[0], mod.gettraj(run=h2, no_collision_check=True))   # before iteration 0: the straight line
    mod.iterate(run=h2, n_iter=3)
    assert np.array_equal(dumps[3], mod.gettraj(run=h2, no_collision_check=True))   # before iteration 3
    mod.iterate(run=h2, n_iter=1)
    assert np.array_equal(final, mod.gettraj(run=h2, no_collision_check=True))
    mod.destroy(run=h)
    mod.destroy(run=h2)


def test_tsr_arguments_of_create(world, oracle, flavour):
    """create ... con_tsr 'all manipee arm' TSR / con_tsr 'start link wam4' TSR / everyn_tsr TSR / start_tsr TSR
    (mod.cpp:1930-1997, tsr_create_parse 3068-3110) against the oracle given the same constraints as numbers"""
    env, mod, table, robot, robot_desc = world
    mod.computedistancefield(kinbody=table, cube_extent=0.02)
    sd = oracle_field_for_table(oracle, flavour)
    ee = robot_desc.names.index("wam7")
    tool = models.pose_make((0.0, 0.0, 0.1), models.quat_from_axis_angle((0, 0, 1), 0.4))
    env.AddManipulator(robot, "arm", "wam7", tool)
    start = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    goal = start + np.array([1.0, 0.03, -0.02, 0.04, 0.0, -0.03, 0.02])
    env.SetActiveDOFValues(robot, start)
    pe = models.pose_compose(oracle.fk(robot_desc, start)[ee], tool)
    T0w, Twe = models.pose_make((0, 0, pe[2])), models.pose_make((0, 0, 0), pe[3:7])
    Bw = np.tile(np.array([-10.0, 10.0]), (6, 1))
    Bw[[2, 3, 4]] = 0.0  # z, roll, pitch
    tsr = orcdchomp.TSR(T0w, Twe, Bw)
    assert len(tsr.serialize().split()) == 38
    con = capi.make_constraint("all", ee, Bw, T0w=T0w, Twe=Twe, pose_link_ee=tool)

    def reference(cons, n_iter=25):
        params = capi.default_params(n_points=40, lambda_=100.0, obs_factor=500.0, constraints=cons)
        run = oracle.Run(robot_desc, params, [sd], start, goal, flavour=flavour)
        ret, c, _, _ = run.iterate(n_iter)
        assert ret == 0
        return run.traj()

    want = reference([con])
    for kw in (dict(con_tsr=("all manipee arm", tsr)), dict(con_tsr=("all", tsr)), dict(everyn_tsr=tsr)):
        traj = mod.runchomp(robot=robot, n_iter=25, lambda_=100.0, obs_factor=500.0, n_points=40, adofgoal=list(goal),
                            no_collision_check=True, **kw)
        assert np.max(np.abs(traj - want)) <= 1e-6
    # a bare link, first moving point only
    elbow = robot_desc.names.index("wam4")
    q1 = start + (goal - start) / 39
    Bx = np.tile(np.array([-10.0, 10.0]), (6, 1))
    Bx[0] = 0.0
    T_el = oracle.fk(robot_desc, q1)[elbow]
    tsr_el = orcdchomp.TSR(T_el, models.pose_make(), Bx)
    traj = mod.runchomp(robot=robot, n_iter=25, lambda_=100.0, obs_factor=500.0, n_points=40, adofgoal=list(goal),
                        no_collision_check=True, con_tsrs=[("start link wam4", tsr_el), ("all manipee arm", tsr)])
    want2 = reference([capi.make_constraint("start", elbow, Bx, T0w=T_el), con])
    assert np.max(np.abs(traj - want2)) <= 1e-6
    # start_tsr: the start row comes back changed
    traj = mod.runchomp(robot=robot, n_iter=25, lambda_=100.0, obs_factor=500.0, n_points=40, adofgoal=list(goal),
                        no_collision_check=True, start_tsr=tsr)
    want3 = reference([capi.make_constraint("start_tsr", ee, Bw, T0w=T0w, Twe=Twe, pose_link_ee=tool)])
    assert np.max(np.abs(traj - want3)) <= 1e-6 and np.max(np.abs(traj[0] - start)) > 1e-4
    # the reference's messages
    with pytest.raises(RuntimeError, match="con_tsr manip not found!"):
        mod.create(robot=robot, adofgoal=list(goal), con_tsr=("all manipee hand", tsr))
    with pytest.raises(RuntimeError, match="con_tsr link not found!"):
        mod.create(robot=robot, adofgoal=list(goal), con_tsr=("all link nothing", tsr))
    with pytest.raises(RuntimeError, match="con_tsr first arg must be start, end, or all!"):
        mod.create(robot=robot, adofgoal=list(goal), con_tsr=("middle", tsr))
    with pytest.raises(RuntimeError, match="Cannot parse constraint TSR!"):
        mod.create(robot=robot, adofgoal=list(goal), con_tsr=("all", "0 NULL 1 2 3"))
    with pytest.raises(RuntimeError, match="You must pass robot before any con_tsrs!"):
        mod.SendCommand("create con_tsr 'all' '%s' robot BarrettWAM adofgoal '0 0 0 0 0 0 0'" % tsr.serialize())
