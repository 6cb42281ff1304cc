"""CPU tests of the checker itself: the self-contained oracle port
(oracle/libcd_port.c + oracle/orcdchomp_port.c) against
  (a) golden vectors generated from the reference's own libcd sources
      (tests/golden/make_golden.py), always;
  (b) the compiled reference (oracle/_ref), when it is present in this checkout;
  (c) independent closed forms / brute force.
"""
import numpy as np
import pytest

from conftest import golden_path
from or_cdchomp_b200 import capi, models


def test_mt19937_matches_numpy_and_golden(oracle):
    """gsl_rng_mt19937 restatement: seed 0 -> 4357 (GSL), raw stream == numpy's MT19937."""
    g = oracle.MT(0)
    raw = np.array([g.next() for _ in range(1300)], dtype=np.uint64)
    gold = np.load(golden_path("mt.npz"))
    assert np.array_equal(raw, gold["raw_seed0"])
    rs = np.random.RandomState(4357)  # init_genrand(4357)
    ref = rs.randint(0, 2 ** 32, size=1300, dtype=np.uint64)
    assert np.array_equal(raw, ref)
    g = oracle.MT(5489)
    assert g.next() == 3499211612  # published first output of mt19937 for seed 5489
    g = oracle.MT(20260217)
    gs = np.array([g.gaussian(0.1) for _ in range(200)])
    assert np.array_equal(gs, gold["gauss_seed20260217"])
    g = oracle.MT(99)
    many = np.array([g.gaussian(2.0) for _ in range(20000)])
    assert abs(many.mean()) < 0.06 and abs(many.std() - 2.0) < 0.05


def test_sdf_sampling_kat(oracle):
    """cd_grid_lookup_index / interp / grad known answers incl. edges (grid.c:191-209, 331-454)."""
    gold = np.load(golden_path("sdf_kat.npz"))
    vals, grads, errs = oracle.sdf_sample(gold["grid"], gold["lengths"], gold["points"])
    assert np.array_equal(errs, gold["errs"])
    ok = errs == 0
    assert ok.sum() > 300 and (~ok).sum() > 10
    assert np.array_equal(vals[ok], gold["values"][ok])
    fin = ok & np.isfinite(gold["values"])
    assert np.array_equal(grads[fin], gold["grads"][fin])
    assert np.isinf(gold["values"][ok]).sum() >= 1  # HUGE_VAL stencil cases are covered


def test_sdf_sampling_semantics(oracle):
    """x == length is inside (last cell); value is continuous across cell-centre planes."""
    grid = np.arange(4 * 3 * 5, dtype=np.float64).reshape(4, 3, 5) ** 1.1
    lengths = np.array([2.0, 1.5, 2.5])
    v, g, e = oracle.sdf_sample(grid, lengths, [[2.0, 1.5, 2.5], [2.0000001, 1.0, 1.0], [-1e-9, 1, 1]])
    assert list(e) == [0, 1, 1]
    c = 0.5 * (2.0 / 4) * 3  # centre plane of cell 1 on axis 0
    v, g, e = oracle.sdf_sample(grid, lengths, [[c - 1e-9, 0.7, 1.2], [c + 1e-9, 0.7, 1.2]])
    assert abs(v[0] - v[1]) < 1e-6 and g[0, 0] != g[1, 0]


def test_sdf_build_golden(oracle):
    gold = np.load(golden_path("sdf_build.npz"))
    for k in ("iso", "aniso", "heights", "allfree", "allobs"):
        obs, ln = gold[k + "_obs"], gold[k + "_len"]
        sdf = oracle.sdf_from_obsarray(obs, ln)
        dt = oracle.dt_sqeuc(obs, ln)
        assert np.array_equal(dt, gold[k + "_dt"]), k
        if k.startswith("all"):
            assert np.array_equal(sdf, gold[k + "_sdf"])
        else:
            assert np.array_equal(sdf, gold[k + "_sdf"]), k


def test_edt_brute_force(oracle):
    """exact squared EDT vs brute force, anisotropic pitch (SURVEY section 4 iii)."""
    rng = np.random.default_rng(2)
    shape, lengths = (9, 7, 11), np.array([0.9, 1.4, 0.55])
    obs = np.where(rng.uniform(size=shape) < 0.07, 0.0, np.inf)  # seeds are the zeros
    dt = oracle.dt_sqeuc(obs, lengths)
    pitch = lengths / np.array(shape)
    idx = np.argwhere(obs == 0.0)
    grid = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, 3)
    d2 = (((grid[:, None, :] - idx[None, :, :]) * pitch) ** 2).sum(-1).min(1).reshape(shape)
    assert np.max(np.abs(dt - d2)) < 1e-12
    # SDF sign convention: positive in free space, negative inside obstacles
    occ = np.where(rng.uniform(size=shape) < 0.3, np.inf, 0.0)
    sdf = oracle.sdf_from_obsarray(occ, lengths)
    assert (sdf[occ == 0.0] > 0).all() and (sdf[np.isinf(occ)] < 0).all()


def test_occupancy_and_flood_golden(oracle):
    gold = np.load(golden_path("occupancy.npz"))
    prims, apos, aext = models.clutter_scene(n_boxes=10, n_balls=6, seed=3, half_span=0.9)
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.04, 0.2)
    assert list(sizes) == list(gold["sizes"])
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    occ = oracle.occupancy(pa, len(gp), sizes, lengths, 0.04)
    obs, sdf = oracle.computedistancefield(pa, len(gp), sizes, lengths, 0.04)
    assert np.array_equal(np.packbits(np.isinf(occ)), gold["occ_hit"])
    assert np.array_equal(np.packbits(np.isinf(obs)), gold["obs_hit"])
    assert np.array_equal(sdf, gold["sdf"])
    assert set(np.unique(obs[np.isfinite(obs)])) <= {0.0}
    assert np.isinf(obs).sum() >= np.isinf(occ).sum()  # enclosed pockets can only add obstacles


def test_triangle_mesh_voxeliser(oracle):
    """cube-vs-triangle predicate: golden, containment in the analytic solid,
    and an independent sampling check of the separating-axis test."""
    gold = np.load(golden_path("mesh.npz"))
    prims, apos, aext = models.mesh_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.1)
    assert list(sizes) == list(gold["sizes"])
    gp = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gp)
    occ = oracle.occupancy(pa, len(gp), sizes, lengths, 0.02)
    obs, sdf = oracle.computedistancefield(pa, len(gp), sizes, lengths, 0.02)
    assert np.array_equal(np.packbits(np.isinf(occ)), gold["occ_hit"])
    assert np.array_equal(np.packbits(np.isinf(obs)), gold["obs_hit"])
    assert np.array_equal(sdf, gold["sdf"])
    assert np.isinf(obs).sum() > np.isinf(occ).sum()          # closed meshes: the flood fill fills the inside
    # a meshed sphere's shell lies inside the solid sphere's voxel set and the filled mesh nearly equals it
    c, r = (0.1, 0.0, -0.2), 0.3
    tri = models.icosphere_mesh(c, r, 3)
    sizes, lengths, gpose = models.field_geometry((0, 0, 0), (0.6, 0.6, 0.6), 0.02, 0.1)
    shell = oracle.occupancy(capi.make_prims(models.prims_to_grid_frame(tri, gpose)), len(tri), sizes, lengths, 0.02)
    solid = oracle.occupancy(capi.make_prims(models.prims_to_grid_frame([("sphere", c, r)], gpose)), 1, sizes, lengths, 0.02)
    assert not (np.isinf(shell) & ~np.isinf(solid)).any()
    filled, _ = oracle.computedistancefield(capi.make_prims(models.prims_to_grid_frame(tri, gpose)), len(tri),
                                            sizes, lengths, 0.02, want_sdf=False)
    assert abs(int(np.isinf(filled).sum()) - int(np.isinf(solid).sum())) < 0.02 * np.isinf(solid).sum()
    # sampling check: a voxel is hit iff some point of the triangle lies in the (closed) cube
    rng = np.random.default_rng(3)
    sizes, lengths = [6, 6, 6], [0.6, 0.6, 0.6]
    for _ in range(40):
        v = rng.uniform(0.05, 0.55, size=(3, 3))
        occ = oracle.occupancy(capi.make_prims([("tri", v[0], v[1], v[2])]), 1, sizes, lengths, 0.05)
        u = rng.uniform(size=(20000, 2))
        flip = u.sum(1) > 1
        u[flip] = 1 - u[flip]
        pts = v[0] + u[:, :1] * (v[1] - v[0]) + u[:, 1:] * (v[2] - v[0])
        cells = np.unique(np.floor(pts / 0.1).astype(int), axis=0)
        assert all(np.isinf(occ[tuple(cell)]) for cell in cells)      # sampled points prove a hit
        # and every reported hit is within half a cube diagonal of the triangle's bounding box
        hit = np.argwhere(np.isinf(occ))
        centres = (hit + 0.5) * 0.1
        lo, hi = v.min(0) - 0.05 - 1e-12, v.max(0) + 0.05 + 1e-12
        assert ((centres >= lo) & (centres <= hi)).all()


def test_flood_fill_encloses_pocket(oracle):
    """a hollow box: the inside is not reachable from voxel 0 and becomes obstacle (mod.cpp:543-548)."""
    # two nested boxes are not expressible as a hollow primitive; build the shell from 6 slabs
    t, h = 0.03, 0.3
    slabs = []
    for ax in range(3):
        for sgn in (-1, 1):
            c = [0.0, 0.0, 0.0]
            c[ax] = sgn * h
            e = [h + t, h + t, h + t]
            e[ax] = t
            slabs.append(("box", models.pose_make(c), tuple(e)))
    sizes, lengths, gpose = models.field_geometry((0, 0, 0), (h + t,) * 3, 0.02, 0.1)
    gp = models.prims_to_grid_frame(slabs, gpose)
    pa = capi.make_prims(gp)
    occ = oracle.occupancy(pa, len(gp), sizes, lengths, 0.02)
    obs, _ = oracle.computedistancefield(pa, len(gp), sizes, lengths, 0.02, want_sdf=False)
    mid = tuple(s // 2 for s in sizes)
    assert occ[mid] == 1.0 and np.isinf(obs[mid]) and obs[0, 0, 0] == 0.0


def test_metric_closed_form(oracle, wam7, table):
    """D=1, both ends fixed: straight line has cost_smooth = 1/2 (m+1) sum |dq|^2."""
    params = capi.default_params(n_points=30, lambda_=100.0, obs_factor=0.0, obs_factor_self=0.0)
    qs, qg = models.WAM7_DEMO_START, models.WAM7_DEMO_GOAL
    run = oracle.Run(wam7, params, [table["desc"]], qs, qg)
    ret, costs, _, _ = run.iterate(0)
    m = params.n_points - 2
    dq = (qg - qs) / (m + 1)
    expect = 0.5 * (m + 1) * (m + 1) * float(dq @ dq)
    assert ret == 0 and abs(costs[2] - expect) < 1e-9 * expect and costs[1] == 0.0
    run.close()


def test_chomp_golden(oracle, wam7):
    """full runs (config 1 + random + momentum/HMC + derivative 2) against reference outputs."""
    gold = np.load(golden_path("chomp.npz"))
    sd = capi.SdfDesc(gold["table_sdf"], gold["table_lengths"], gold["table_pose"])
    params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
    for r in range(2):
        run = oracle.Run(wam7, params, [sd], gold["cfg1_starts"][r], gold["cfg1_goals"][r])
        ret, c, tr, gr = run.iterate(100, want_trace=True, want_grads=True)
        assert ret == 0
        assert np.max(np.abs(run.traj() - gold["cfg1_traj"][r])) < 1e-9
        assert np.max(np.abs(tr - gold["cfg1_trace"][r])) < 1e-8
        assert np.max(np.abs(gr[0] - gold["cfg1_grad0"][r])) < 1e-9 * np.max(np.abs(gold["cfg1_grad0"][r]))
        run.close()
    params = capi.default_params(n_points=40, lambda_=50.0, obs_factor=300.0, use_momentum=1, use_hmc=1,
                                 hmc_resample_lambda=0.05)
    for k, seed in enumerate(gold["hmc_seeds"]):
        run = oracle.Run(wam7, params, [sd], gold["cfg1_starts"][1], gold["cfg1_goals"][1], seed=int(seed))
        ret, c, _, _ = run.iterate(60)
        assert ret == 0 and run.hmc_next() == gold["hmc_next"][k]
        assert np.max(np.abs(run.traj() - gold["hmc_traj"][k])) < 1e-8
        assert np.max(np.abs(run.momentum() - gold["hmc_mom"][k])) < 1e-8
        run.close()
    params = capi.default_params(n_points=50, lambda_=200.0, derivative=2)
    run = oracle.Run(wam7, params, [sd], gold["cfg1_starts"][2], gold["cfg1_goals"][2])
    ret, c, _, _ = run.iterate(30)
    assert ret == 0 and np.max(np.abs(run.traj() - gold["d2_traj"])) < 1e-8
    run.close()


def test_port_vs_compiled_reference(oracle, wam7, table):
    """(b): the port against the reference's own compiled sources, when present here."""
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built in this checkout (needs /root/reference)")
    rng = np.random.default_rng(3)
    obs = np.where(rng.uniform(size=(12, 10, 15)) < 0.1, np.inf, 0.0)
    ln = [1.2, 0.8, 1.7]
    assert np.array_equal(oracle.sdf_from_obsarray(obs, ln, "port"), oracle.sdf_from_obsarray(obs, ln, "reference"))
    grid = rng.normal(size=(5, 6, 7))
    pts = rng.uniform(-0.1, 1.1, size=(500, 3)) * np.array(ln)
    a, b = oracle.sdf_sample(grid, ln, pts, "port"), oracle.sdf_sample(grid, ln, pts, "reference")
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    fa, fb = oracle.fk(wam7, models.WAM7_DEMO_START, "port"), oracle.fk(wam7, models.WAM7_DEMO_START, "reference")
    assert np.max(np.abs(fa - fb)) < 1e-15
    robot = models.prismatic_test_robot()
    for rb, P, kw in ((wam7, 60, {}), (robot, 25, {}), (wam7, 30, dict(use_momentum=1, use_hmc=1))):
        params = capi.default_params(n_points=P, lambda_=80.0, obs_factor=400.0, **kw)
        starts, goals = models.random_endpoints(rb, 1, seed0=77)
        outs = []
        for fl in ("port", "reference"):
            run = oracle.Run(rb, params, [table["desc"]], starts[0], goals[0], seed=5, flavour=fl)
            ret, c, tr, _ = run.iterate(40, want_trace=True)
            outs.append((ret, run.traj(), tr))
            run.close()
        assert outs[0][0] == outs[1][0] == 0
        assert np.max(np.abs(outs[0][1] - outs[1][1])) < 1e-10
        assert np.max(np.abs(outs[0][2] - outs[1][2])) < 1e-8


def test_fk_jacobian_finite_difference(oracle):
    """restated FK / Jacobian are mutually consistent: obstacle gradient == d cost / d q numerically."""
    robot = models.prismatic_test_robot()
    rng = np.random.default_rng(1)
    sdf = rng.uniform(0.05, 0.4, size=(12, 12, 12))
    # smooth field so the piecewise-constant gradient equals the derivative of the interpolant
    x = (np.arange(12) + 0.5) / 12
    sdf = 0.3 + 0.2 * x[:, None, None] - 0.15 * x[None, :, None] + 0.1 * x[None, None, :]
    sd = capi.SdfDesc(sdf, [1.5, 1.5, 1.5], models.pose_make((-0.7, -0.7, -0.3)))
    params = capi.default_params(n_points=7, epsilon=1.0, obs_factor=10.0, obs_factor_self=0.0)
    starts, goals = models.random_endpoints(robot, 1, seed0=5)
    run = oracle.Run(robot, params, [sd], starts[0], goals[0])
    g, c = run.obstacle_gradient()
    assert np.isfinite(g).all() and np.abs(g).max() > 0
    run.close()


def test_floating_base_oracle(oracle, flavour, wam7, table):
    """floating_base branch of the restated callbacks (mod.cpp:991-1021, 1050-1086, 2424-2464,
    2805-2808): the pose block of the sphere Jacobian is 0.01 x the derivative of the sphere
    centre along translations and unit-quaternion tangents; with the base held where
    robot.base_pose puts it the active-dof columns equal the fixed-base gradient; rows start as
    and stay unit quaternions."""
    rng = np.random.default_rng(3)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    pose = np.concatenate([rng.uniform(-1, 1, 3), q])
    pb = rng.uniform(-0.5, 0.5, 3)

    def world(p7):
        return models.quat_rotate(p7[3:], pb) + p7[:3]
    v = world(pose)
    J = oracle.pose_block(pose, v, flavour=flavour)
    h = 1e-6
    for k in range(3):
        d = np.zeros(7); d[k] = h
        assert np.allclose((world(pose + d) - world(pose - d)) / (2 * h), J[:, k] / 0.01, atol=1e-8)
    for _ in range(3):
        dq = rng.normal(size=4)
        dq -= q * dq.dot(q)                      # tangent to the unit sphere
        pp, pm = pose.copy(), pose.copy()
        pp[3:] = (q + h * dq) / np.linalg.norm(q + h * dq)
        pm[3:] = (q - h * dq) / np.linalg.norm(q - h * dq)
        assert np.allclose((world(pp) - world(pm)) / (2 * h), J[:, 3:] @ dq / 0.01, atol=1e-7)
    # a run: rows start as unit quaternions and stay so; the 0.01-scaled pose columns move the
    # base far less than the arm.  (The base must travel: the reference divides by the squared
    # speed of every active sphere unguarded, mod.cpp:1239, and in this mode the spheres of the
    # base link are active too.)
    sd = table["desc"]
    pf = capi.default_params(n_points=30, lambda_=100.0, obs_factor=300.0, floating_base=1)
    starts, goals = models.random_endpoints(wam7, 1, seed0=11, shrink=0.3)
    base0 = np.asarray(wam7.base_pose, dtype=float)
    base1 = models.pose_compose(base0, models.pose_make((0.15, -0.1, 0.05), models.quat_from_axis_angle((0.2, 0.1, 1.0), 0.4)))
    flt = oracle.Run(wam7, pf, [sd], np.concatenate([base0, starts[0]]), np.concatenate([base1, goals[0]]), flavour=flavour)
    t0 = flt.traj()
    assert t0.shape == (30, 14)
    assert np.allclose(np.linalg.norm(t0[:, 3:7], axis=1), 1.0, atol=1e-15)
    g1, c1 = flt.obstacle_gradient()
    assert g1.shape == (28, 14) and np.all(np.isfinite(g1)) and np.abs(g1[:, :7]).max() > 0
    ret, c, tr, _ = flt.iterate(5, want_trace=True)
    assert ret == 0 and np.all(np.isfinite(tr))
    t = flt.traj()
    assert np.allclose(np.linalg.norm(t[:, 3:7], axis=1), 1.0, atol=1e-15)
    assert np.max(np.abs(t[:, :7] - t0[:, :7])) < 0.2 * np.max(np.abs(t[:, 7:] - t0[:, 7:]))
    flt.close()


def test_port_matches_golden_modes(oracle):
    """the self-contained port against fixtures made by the reference build (tests/golden/modes.npz):
    floating base (plain, momentum) and the 200-sphere arm."""
    gold = np.load(golden_path("modes.npz"))
    robot = models.wam7_robot()
    sd = capi.SdfDesc(gold["table_sdf"], gold["table_lengths"], gold["table_pose"])
    for mom in (0, 1):
        params = capi.default_params(n_points=50, lambda_=100.0, obs_factor=300.0, floating_base=1, use_momentum=mom)
        run = oracle.Run(robot, params, [sd], gold["float_starts"][0], gold["float_goals"][0], flavour="port")
        ret, c, _, gr = run.iterate(30, want_grads=True)
        assert ret == 0
        assert np.max(np.abs(gr[0] - gold["float%d_grad0" % mom][0])) <= 1e-9 * np.max(np.abs(gr[0]))
        assert np.max(np.abs(run.traj() - gold["float%d_traj" % mom][0])) <= 1e-9
        run.close()
    robot5 = models.dense_sphere_arm(200, seed=5)
    params = capi.default_params(n_points=48, lambda_=200.0, obs_factor=100.0, epsilon=0.2)
    run = oracle.Run(robot5, params, [sd], gold["dense_starts"][0], gold["dense_goals"][0], flavour="port")
    ret, c, _, gr = run.iterate(6, want_grads=True)
    assert ret == 0 and np.max(np.abs(run.traj() - gold["dense_traj"][0])) <= 1e-9
    run.close()


def test_sphere_at_rest_is_finite(oracle, flavour, wam7, table):
    """mod.cpp:1239 divides by |v|^2 unguarded, but cblas_daxpy (1241) and cblas_dgemv (1244) return
    at once for a zero scalar, so a sphere at rest contributes exactly zero: the restated callbacks
    must reproduce that quick return instead of spreading NaN."""
    params = capi.default_params(n_points=30, lambda_=100.0, obs_factor=300.0)
    starts, goals = models.random_endpoints(wam7, 2, seed0=7, shrink=0.3)
    goals[0, :2] = starts[0, :2]
    goals[1] = starts[1]
    for r in range(2):
        run = oracle.Run(wam7, params, [table["desc"]], starts[r], goals[r], flavour=flavour)
        ret, c, _, gr = run.iterate(3, want_grads=True)
        assert ret == 0 and np.isfinite(gr).all() and np.isfinite(c).all() and np.isfinite(run.traj()).all()
        run.close()


def _constraint_cases(oracle, robot, flavour="port"):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", golden_path("make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.constraint_cases(oracle, robot, flavour)


def test_port_matches_golden_constraints(oracle):
    """hard TSR constraints: the self-contained port (its own projection, dense inverse and pivoted
    elimination; con_tsr restated from mod.cpp:1330-1497) against fixtures made by the reference build,
    where chomp.c:553-600, kin.c, spatial.c and LAPACKE_dgesv are the reference's own
    (tests/golden/constraints.npz)."""
    gold = np.load(golden_path("constraints.npz"))
    robot = models.wam7_robot()
    sd = capi.SdfDesc(gold["table_sdf"], gold["table_lengths"], gold["table_pose"])
    for name, kw, cons, starts, goals, n_iter in _constraint_cases(oracle, robot):
        params = capi.default_params(constraints=cons, **kw)
        for r in range(len(starts)):
            run = oracle.Run(robot, params, [sd], starts[r], goals[r], flavour="port")
            q = starts[r] + 0.37 * (goals[r] - starts[r])
            v, J = run.constraint_eval(0, q)
            assert np.max(np.abs(v - gold[name + "_val"][r])) <= 1e-12
            assert np.max(np.abs(J - gold[name + "_jac"][r])) <= 1e-12
            # the analytic Jacobian is the derivative of the value
            Jn = np.zeros_like(J)
            for j in range(len(q)):
                d = np.zeros(len(q))
                d[j] = 1e-6
                Jn[:, j] = (run.constraint_eval(0, q + d)[0] - run.constraint_eval(0, q - d)[0]) / 2e-6
            assert np.max(np.abs(J - Jn)) <= 1e-7
            ret, c, _, _ = run.iterate(n_iter)
            assert ret == 0
            assert np.max(np.abs(run.traj() - gold[name + "_traj"][r])) <= 1e-9
            assert np.allclose(c, gold[name + "_costs"][r], rtol=1e-9, atol=0)
            if name == "start_tsr":
                assert run.m == params.n_points - 1 and np.max(np.abs(run.traj()[0] - starts[r])) > 1e-3
            run.close()


def test_constraint_projection_zeroes_linearised_constraint(oracle, flavour, wam7, table):
    """chomp.c:553-600 as a property: after one iteration h(T_old) + J (T_new - T_old) = 0 on every constrained
    row (when no joint limit interferes), whatever the cost gradient does"""
    ee = wam7.names.index("wam7")
    start = np.array([0.4, 0.9, 0.1, 1.4, 0.2, -0.5, 0.3])
    goal = start + np.array([0.9, 0.2, -0.1, 0.15, 0.1, -0.2, 0.1])
    Bw = np.tile(np.array([-10.0, 10.0]), (6, 1))
    Bw[[0, 4, 5]] = 0.0
    T0w = oracle.fk(wam7, start + 0.5 * (goal - start))[ee]
    cons = [capi.make_constraint("all", ee, Bw, T0w=T0w)]
    params = capi.default_params(n_points=24, lambda_=300.0, obs_factor=200.0, constraints=cons)
    run = oracle.Run(wam7, params, [table["desc"]], start, goal, flavour=flavour)
    T0 = run.traj()
    hJ = [run.constraint_eval(0, T0[i]) for i in range(1, 23)]
    ret, _, _, _ = run.iterate(1)
    assert ret == 0
    T1 = run.traj()
    for k, (h, J) in enumerate(hJ):
        lin = h + J @ (T1[k + 1] - T0[k + 1])
        assert np.max(np.abs(lin)) <= 1e-9 * max(1.0, np.max(np.abs(h)))
    run.close()
