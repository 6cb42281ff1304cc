"""GPU parity tests of the SDF build pipeline through the C ABI: bit-exact occupancy,
SDF within 1e-12 (BASELINE.json), against the CPU oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from or_cdchomp_b200 import capi, models

pytestmark = pytest.mark.gpu
SDF_ATOL = 1e-12


def test_sdf_build_golden(engine):
    gold = np.load(golden_path("sdf_build.npz"))
    for k in ("iso", "aniso", "heights", "allfree", "allobs"):
        sdf = engine.sdf_build(gold[k + "_obs"], gold[k + "_len"])
        ref = gold[k + "_sdf"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(sdf), fin), k
        assert np.array_equal(sdf[~fin], ref[~fin]), k
        assert np.max(np.abs(sdf[fin] - ref[fin]), initial=0.0) <= SDF_ATOL, k


def test_sdf_build_random_shapes(engine, oracle, flavour):
    rng = np.random.default_rng(17)
    cases = [((2, 2, 2), (1, 1, 1), 0.3), ((33, 17, 40), (3.3, 1.7, 4.0), 0.05), ((40, 40, 40), (0.8, 0.8, 0.8), 0.02),
             ((5, 64, 7), (1.0, 3.0, 0.2), 0.2), ((64, 3, 31), (0.64, 0.09, 0.5), 0.01)]
    for shape, lens, frac in cases:
        obs = np.where(rng.uniform(size=shape) < frac, np.inf, 0.0)
        sdf = engine.sdf_build(obs, lens)
        ref = oracle.sdf_from_obsarray(obs, lens, flavour=flavour)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(sdf), fin)
        assert np.max(np.abs(sdf[fin] - ref[fin]), initial=0.0) <= SDF_ATOL, shape


def test_fast_integer_path_equals_general_and_oracle(engine, oracle, flavour):
    """binary grids with cubic cells take the exact integer path (sdf_fast.cu); it must agree
    with the general fp64 path and with the oracle for awkward sizes, empty / full grids,
    isolated seeds and lines without any seed."""
    rng = np.random.default_rng(23)
    cases = []
    for shape, frac in (((2, 2, 2), 0.5), ((7, 5, 33), 0.1), ((40, 37, 70), 0.01), ((33, 64, 31), 0.3),
                        ((16, 16, 100), 0.002), ((50, 3, 65), 0.05)):
        cases.append(np.where(rng.uniform(size=shape) < frac, np.inf, 0.0))
    one = np.zeros((20, 20, 40)); one[3, 17, 38] = np.inf
    hole = np.full((20, 20, 40), np.inf); hole[10, 2, 0] = 0.0
    cases += [one, hole, np.zeros((5, 6, 7)), np.full((5, 6, 7), np.inf)]
    for obs in cases:
        pitch = 0.013
        lens = [pitch * k for k in obs.shape]
        engine.force_general_sdf(False)
        fast = engine.sdf_build(obs, lens)
        engine.force_general_sdf(True)
        gen = engine.sdf_build(obs, lens)
        engine.force_general_sdf(False)
        ref = oracle.sdf_from_obsarray(obs, lens, flavour=flavour)
        fin = np.isfinite(ref)
        for got in (fast, gen):
            assert np.array_equal(np.isfinite(got), fin), obs.shape
            assert np.array_equal(got[~fin], ref[~fin]), obs.shape
            assert np.max(np.abs(got[fin] - ref[fin]), initial=0.0) <= SDF_ATOL, obs.shape
    # a grid that is not 0 / HUGE_VAL falls back to the general path on its own
    o = cases[2].copy()
    o[1, 1, 1] = 0.25
    lens = [0.02 * k for k in o.shape]
    got = engine.sdf_build(o, lens)
    ref = oracle.sdf_from_obsarray(o, lens, flavour=flavour)
    assert np.max(np.abs(got - ref)) <= SDF_ATOL


def test_dt_sqeuc_device(engine, oracle, flavour):
    """cd_grid_double_dt_sqeuc alone, on HBM pointers, arbitrary finite heights."""
    rng = np.random.default_rng(4)
    f = rng.uniform(0, 0.5, size=(12, 20, 9))
    f[rng.uniform(size=f.shape) < 0.6] = np.inf
    lens = (1.2, 1.0, 1.8)
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty_like(d_in)
    engine.dt_sqeuc_device(d_in.data_ptr(), f.shape, lens, d_out.data_ptr())
    engine.sync()
    ref = oracle.dt_sqeuc(f, lens, flavour=flavour)
    assert np.max(np.abs(d_out.cpu().numpy() - ref)) <= SDF_ATOL


def test_occupancy_flood_sdf_pipeline(engine, oracle, flavour):
    """computedistancefield after the sizing: occupancy bit-exact, flood + relabel bit-exact, SDF 1e-12."""
    gold = np.load(golden_path("occupancy.npz"))
    prims, apos, aext = models.clutter_scene(n_boxes=10, n_balls=6, seed=3, half_span=0.9)
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.04, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    n = int(np.prod(sizes))
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    engine.occupancy_device(gp, sizes, lengths, 0.04, d.data_ptr())
    occ = d.cpu().numpy().reshape(sizes)
    assert set(np.unique(occ[np.isfinite(occ)])) == {1.0}
    assert np.array_equal(np.packbits(np.isinf(occ)), gold["occ_hit"])
    engine.flood_relabel_device(d.data_ptr(), sizes, 0)
    engine.sync()
    obs = d.cpu().numpy().reshape(sizes)
    assert np.array_equal(np.packbits(np.isinf(obs)), gold["obs_hit"])
    obs2, sdf = engine.computedistancefield(gp, sizes, lengths, 0.04)
    assert np.array_equal(obs2, obs)
    assert np.max(np.abs(sdf - gold["sdf"])) <= SDF_ATOL
    # a second, larger random scene directly against the oracle
    prims, apos, aext = models.clutter_scene(n_boxes=24, n_balls=12, seed=8, half_span=1.2)
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.025, 0.2)
    gp = models.prims_to_grid_frame(prims, gpose)
    obs_g, sdf_g = engine.computedistancefield(gp, sizes, lengths, 0.025)
    pa = capi.make_prims(gp)
    obs_r, sdf_r = oracle.computedistancefield(pa, len(gp), sizes, lengths, 0.025, flavour=flavour)
    assert np.array_equal(obs_g, obs_r)
    assert np.max(np.abs(sdf_g - sdf_r)) <= SDF_ATOL


def test_triangle_mesh_pipeline(engine, oracle, flavour):
    """triangle-mesh kinbody (icospheres + box meshes): occupancy, flood fill and SDF against the
    golden vectors and the oracle; the voxel set must be bit-exact."""
    gold = np.load(golden_path("mesh.npz"))
    prims, apos, aext = models.mesh_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.1)
    gp = models.prims_to_grid_frame(prims, gpose)
    n = int(np.prod(sizes))
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    engine.occupancy_device(gp, sizes, lengths, 0.02, d.data_ptr())
    occ = d.cpu().numpy().reshape(sizes)
    assert np.array_equal(np.packbits(np.isinf(occ)), gold["occ_hit"])
    obs, sdf = engine.computedistancefield(gp, sizes, lengths, 0.02)
    assert np.array_equal(np.packbits(np.isinf(obs)), gold["obs_hit"])
    assert np.max(np.abs(sdf - gold["sdf"])) <= SDF_ATOL
    # a finer, larger mesh directly against the oracle (many small triangles per voxel and vice versa)
    tri = models.icosphere_mesh((0.05, -0.1, 0.0), 0.45, 4) + models.box_mesh(
        models.pose_make((-0.2, 0.3, 0.1), models.quat_from_axis_angle((1, 2, 3), 0.7)), (0.5, 0.05, 0.3))
    sizes, lengths, gpose = models.field_geometry((0, 0, 0), (0.7, 0.7, 0.7), 0.0125, 0.1)
    gp = models.prims_to_grid_frame(tri, gpose)
    obs_g, _ = engine.computedistancefield(gp, sizes, lengths, 0.0125, want_sdf=False)
    obs_r, _ = oracle.computedistancefield(capi.make_prims(gp), len(gp), sizes, lengths, 0.0125, flavour=flavour,
                                           want_sdf=False)
    assert np.array_equal(obs_g, obs_r)


def test_flood_encloses_pocket(engine):
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
    obs, sdf = engine.computedistancefield(gp, sizes, lengths, 0.02)
    mid = tuple(s // 2 for s in sizes)
    assert np.isinf(obs[mid]) and obs[0, 0, 0] == 0.0 and sdf[mid] < 0 and sdf[0, 0, 0] > 0


def test_sdf_sampling_matches_oracle_through_cost(engine, oracle, flavour, wam7, table):
    """cd_grid_double_interp / grad on the device are exercised through the obstacle-only
    gradient with self-collision off, against the oracle (covers in-range / out-of-range spheres)."""
    params = capi.default_params(n_points=60, lambda_=100.0, obs_factor=500.0, obs_factor_self=0.0)
    starts, goals = models.random_endpoints(wam7, 8, seed0=555)
    sid = engine.upload_sdf(table["desc"])
    b = engine.create_batch(wam7, params, [sid], starts, goals)
    b.capture_gradient(2)
    b.iterate(1)
    g = b.get_gradient()
    hit = 0
    for r in range(8):
        run = oracle.Run(wam7, params, [table["desc"]], starts[r], goals[r], flavour=flavour)
        og, oc = run.obstacle_gradient()
        og = og / (params.n_points - 2)
        if np.abs(og).max() > 0:
            hit += 1
            assert np.max(np.abs(g[r] - og)) <= 1e-9 * np.abs(og).max()
        run.close()
    assert hit >= 4
    b.close()
    engine.remove_sdf(sid)


def test_full_size_sdf_properties(engine):
    """BASELINE configs[2] scale (400^3, cube_extent 0.005 -> 0.01 m voxels): properties that do
    not need the CPU oracle at this size."""
    prims, apos, aext = models.clutter_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.005, 0.2)
    assert list(sizes) == [400, 400, 400]
    gp = models.prims_to_grid_frame(prims, gpose)
    n = int(np.prod(sizes))
    d_obs = torch.empty(n, dtype=torch.float64, device="cuda")
    d_sdf = torch.empty(n, dtype=torch.float64, device="cuda")
    engine.occupancy_device(gp, sizes, lengths, 0.005, d_obs.data_ptr())
    engine.flood_relabel_device(d_obs.data_ptr(), sizes, 0)
    engine.sdf_build_device(d_obs.data_ptr(), sizes, lengths, d_sdf.data_ptr())
    engine.sync()
    obs = d_obs.view(400, 400, 400)
    sdf = d_sdf.view(400, 400, 400)
    occ = torch.isinf(obs)
    assert bool(torch.isfinite(sdf).all())
    assert 0.005 < float(occ.double().mean()) < 0.5
    assert bool((sdf[occ] < 0).all()) and bool((sdf[~occ] > 0).all())          # sign convention
    pitch = 0.01
    for ax in range(3):                          # 1-Lipschitz inside each region; the centre-to-centre
        a = sdf.narrow(ax, 0, 399)               # convention gives +1 / -1 voxel across the obstacle boundary
        bb = sdf.narrow(ax, 1, 399)
        same = (a > 0) == (bb > 0)
        diff = (a - bb).abs()
        assert float(diff[same].max()) <= pitch * (1 + 1e-9)
        assert float(diff[~same].max()) <= 2 * pitch * (1 + 1e-9)
    # free cells next to an obstacle are exactly one voxel away, and vice versa
    near = occ[1:, :, :] ^ occ[:-1, :, :]
    assert float((sdf[1:, :, :][near].abs() - pitch).abs().max()) < 1e-12
    # exactness on a random sample of cells against brute force over the obstacle set
    oc = occ.nonzero().double()
    g = torch.Generator(device="cpu").manual_seed(0)
    pick = torch.randint(0, 400, (64, 3), generator=g).cuda()
    for p in pick:
        if bool(occ[p[0], p[1], p[2]]):
            continue
        d = ((oc - p.double()) * pitch).pow(2).sum(1).min().sqrt()
        assert abs(float(d) - float(sdf[p[0], p[1], p[2]])) < 1e-12


def test_libcd_named_entry_points(oracle, flavour):
    """libcd_b200.so: cd_grid_double_bin_sdf / dt_sqeuc / sedt under libcd's names and struct
    layout (grid.h:29-41, 84-93) return freshly malloc'ed grids holding the oracle's values;
    the fused flood + relabel call works in place on a caller-owned grid."""
    from or_cdchomp_b200 import libcd
    rng = np.random.default_rng(31)
    gold = np.load(golden_path("sdf_build.npz"))
    for k in ("iso", "aniso", "heights"):
        sdf, lens = libcd.bin_sdf(gold[k + "_obs"], gold[k + "_len"])
        ref = gold[k + "_sdf"]
        fin = np.isfinite(ref)
        assert np.allclose(lens, gold[k + "_len"], rtol=0, atol=0)
        assert np.array_equal(np.isfinite(sdf), fin) and np.max(np.abs(sdf[fin] - ref[fin]), initial=0.0) <= SDF_ATOL
    obs = np.where(rng.uniform(size=(21, 34, 18)) < 0.04, np.inf, 0.0)
    lens = [0.42, 0.68, 0.36]
    sdf, _ = libcd.bin_sdf(obs, lens)
    assert np.max(np.abs(sdf - oracle.sdf_from_obsarray(obs, lens, flavour=flavour))) <= SDF_ATOL
    func = np.where(rng.uniform(size=(12, 9, 20)) < 0.1, 0.0, np.inf)
    func[rng.uniform(size=func.shape) < 0.05] = 0.37  # parabola heights (grid.c:274-304)
    ref = oracle.dt_sqeuc(func, [1.2, 0.9, 2.0], flavour=flavour)
    for legacy in (False, True):
        dt, _ = libcd.dt_sqeuc(func, [1.2, 0.9, 2.0], legacy_name=legacy)
        assert np.max(np.abs(dt - ref)) <= SDF_ATOL
    # flood + relabel: closed shell around a pocket (mod.cpp:536-548)
    g = np.ones((16, 16, 16))
    g[4:12, 4:12, 4:12] = np.inf
    g[6:10, 6:10, 6:10] = 1.0
    out = libcd.flood_relabel(g, [1, 1, 1], 0)
    want = np.zeros_like(g)
    want[4:12, 4:12, 4:12] = np.inf
    assert np.array_equal(out, want)
    # error paths keep libcd's codes
    with pytest.raises(libcd.LibcdError) as ei:
        libcd.bin_sdf(np.zeros((4, 4)), [1, 1])
    assert ei.value.code == -2


def test_resident_fields_and_pose_aliases(engine, wam7, table):
    """fields built without a host round trip (ocb_computedistancefield_resident,
    ocb_sdf_build_resident) hold the same values as the host variants; an alias of a resident
    grid with another pose behaves like an upload of the grid with that pose."""
    sizes, lengths, gprims = table["sizes"], table["lengths"], table["gprims"]
    obs, sdf = engine.computedistancefield(gprims, sizes, lengths, 0.02)
    sid = engine.computedistancefield_resident(gprims, sizes, lengths, 0.02, table["pose_world"])
    assert np.array_equal(engine.download_sdf(sid, sizes), sdf)
    sid2 = engine.sdf_build_resident(obs, lengths, table["pose_world"])
    assert np.array_equal(engine.download_sdf(sid2, sizes), sdf)
    other = models.pose_compose(models.pose_make((0.1, -0.05, 0.02), models.quat_from_axis_angle((0, 0, 1), 0.3)),
                                table["pose_world"])
    alias = engine.alias_sdf(sid, other)
    up = engine.upload_sdf(capi.SdfDesc(sdf, lengths, other))
    params = capi.default_params(n_points=40, lambda_=100.0, obs_factor=400.0)
    starts, goals = models.random_endpoints(wam7, 3, seed0=5, shrink=0.3)
    res = []
    for s in (alias, up):
        b = engine.create_batch(wam7, params, [s], starts, goals)
        b.iterate(6)
        res.append(b.get_traj())
        b.close()
    assert np.array_equal(res[0], res[1])
    for s in (alias, up, sid2, sid):
        engine.remove_sdf(s)
    engine.trim()  # the pool hands its unused memory back; the engine stays usable
    again = engine.computedistancefield_resident(gprims, sizes, lengths, 0.02, table["pose_world"])
    assert np.array_equal(engine.download_sdf(again, sizes), sdf)
    engine.remove_sdf(again)


def test_flood_fill_random_grids(engine):
    """the bit-mask flood fill against 6-connected component labelling (scipy) on random grids:
    axes that are not multiples of 32, a single z word, thin grids, winding corridors, start cells
    anywhere, a start cell that is not fillable (grid_flood.c:30-111, mod.cpp:543-548)."""
    from scipy import ndimage
    rng = np.random.default_rng(77)
    six = ndimage.generate_binary_structure(3, 1)
    shapes = [(5, 7, 3), (9, 4, 31), (6, 5, 32), (4, 6, 33), (12, 10, 70), (3, 3, 129), (40, 37, 45), (1, 1, 50), (2, 64, 2)]
    for shape in shapes:
        for frac in (0.25, 0.45, 0.62):   # around the site-percolation threshold: long winding paths
            open_cells = rng.uniform(size=shape) > frac
            grid = np.where(open_cells, 1.0, np.inf)
            grid[rng.uniform(size=shape) < 0.02] = 0.37            # other values conduct nothing and stay
            open_cells = grid == 1.0
            labels, _ = ndimage.label(open_cells, structure=six)
            flat_open = np.flatnonzero(open_cells.ravel())
            starts = [0, int(np.prod(shape)) - 1]
            if len(flat_open):
                starts += [int(rng.choice(flat_open)) for _ in range(2)]
            for start in starts:
                got = engine.flood_relabel(grid, start)
                want = grid.copy()
                lab = labels.ravel()[start]
                if lab > 0:
                    want[labels == lab] = 0.0
                    want[open_cells & (labels != lab)] = np.inf
                else:                                              # start cell not fillable: nothing is reached
                    want[open_cells] = np.inf
                assert np.array_equal(got, want), (shape, frac, start)


def test_fast_path_solid_obstacles(engine):
    """the sparse free-cell field of the integer path (only zero-height samples at the ends of a
    run are kept, results only between the first and last obstacle cell of a line) on thick solid
    obstacles with cavities, thin walls, slabs touching the grid faces and lines that are entirely
    obstacle -- against the general fp64 path, which follows grid.c:462-569 operation for operation."""
    rng = np.random.default_rng(91)
    for shape in ((70, 65, 90), (33, 100, 47), (64, 64, 64)):
        obs = np.zeros(shape)
        for _ in range(6):                                  # solid boxes, some touching the faces
            lo = [rng.integers(0, s - 2) for s in shape]
            hi = [min(s, l + rng.integers(2, max(3, s // 2))) for s, l in zip(shape, lo)]
            obs[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = np.inf
        for _ in range(3):                                  # cavities inside them
            c = [rng.integers(2, s - 2) for s in shape]
            obs[c[0] - 1:c[0] + 2, c[1] - 1:c[1] + 2, c[2] - 1:c[2] + 2] = 0.0
        obs[:, shape[1] // 3, :] = np.inf                   # a full wall: whole lines of obstacle
        obs[shape[0] // 2, :, shape[2] // 2] = np.inf       # a rod
        obs[rng.uniform(size=shape) < 0.002] = np.inf       # specks
        pitch = 0.01
        lens = [pitch * k for k in shape]
        engine.force_general_sdf(False)
        fast = engine.sdf_build(obs, lens)
        engine.force_general_sdf(True)
        gen = engine.sdf_build(obs, lens)
        engine.force_general_sdf(False)
        assert np.all(np.isfinite(gen)) and np.all((fast < 0) == (obs != 0))
        assert np.max(np.abs(fast - gen)) <= SDF_ATOL, shape
