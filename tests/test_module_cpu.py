"""Host logic of the module boundary that needs no GPU: the Python command builders emit
exactly the strings the reference's pythonsrc/orcdchomp/orcdchomp.py emits (same keywords,
order and number formatting), and the module header's symbols are exported."""
import os
import re

import numpy as np

from conftest import ROOT
from or_cdchomp_b200 import orcdchomp


class FakeMod:
    def __init__(self):
        self.sent = []

    def SendCommand(self, cmd, releasegil=False):
        self.sent.append(cmd)
        return "0x1234"

    def GetName(self):
        return "BarrettWAM"


def test_command_strings_match_reference_formatting():
    m = FakeMod()
    orcdchomp.bind(m)
    m.computedistancefield(kinbody=m, cube_extent=0.005, aabb_padding=0.2, cache_filename="sdf it's.dat",
                           require_cache=True)
    assert m.sent[-1] == ("computedistancefield kinbody 'BarrettWAM' cube_extent 0.005000 aabb_padding 0.200000 "
                          "cache_filename 'sdf it'\\''s.dat' require_cache")
    m.create(robot="r", adofgoal=[0.5, 1, -2.25], lambda_=100.0, n_points=100, use_momentum=True, use_hmc=True,
             hmc_resample_lambda=0.02, seed=7, epsilon=0.1, epsilon_self=0.04, obs_factor=500.0,
             obs_factor_self=10.0, derivative=1, dat_filename="x.dat")
    assert m.sent[-1] == ("create robot 'r' adofgoal '0.5 1 -2.25' lambda 100.0000 n_points 100 derivative 1 "
                          "use_momentum use_hmc hmc_resample_lambda 0.020000 seed 7 epsilon 0.100000 "
                          "epsilon_self 0.040000 obs_factor 500.000000 obs_factor_self 10.000000 "
                          "dat_filename 'x.dat'")
    cost = [None]
    m.SendCommand = lambda cmd, releasegil=False: (m.sent.append(cmd), "12.5")[1]
    m.iterate(run="0xdead", n_iter=100, max_time=2.5, cost=cost)
    assert m.sent[-1] == "iterate run 0xdead n_iter 100 max_time 2.500000" and cost[0] == 12.5
    m.destroy(run="0xdead")
    assert m.sent[-1] == "destroy run 0xdead"
    m.removefield(kinbody="table")
    assert m.sent[-1] == "removefield kinbody 'table'"
    m.addfield_fromobsarray(kinbody="k", obsarray="0x10", sizes=[2, 3, 4], lengths=[1.0, 1.5, 2.0],
                            pose=[0, 0, 0, 0, 0, 0, 1])
    assert m.sent[-1] == ("addfield_fromobsarray kinbody 'k' obsarray 0x10 sizes '2 3 4' lengths '1.0 1.5 2.0' "
                          "pose '0 0 0 0 0 0 1'")
    m.viewspheres(robot="r")
    assert m.sent[-1] == "viewspheres robot 'r'"


def test_runchomp_sequence():
    m = FakeMod()
    orcdchomp.bind(m)
    xml = ('<trajectory>\n<configuration>\n<group name="joint_values r 0 1" offset="0" dof="2" '
           'interpolation="linear"/>\n<group name="deltatime" offset="2" dof="1" interpolation=""/>\n'
           '</configuration>\n<data count="3">\n0 1 0 0.5 1.5 0.5 1 2 0.5 \n</data>\n</trajectory>\n')

    def send(cmd, releasegil=False):
        m.sent.append(cmd)
        return {"create": "0xabc", "iterate": "3.5", "gettraj": xml, "destroy": ""}[cmd.split()[0]]

    m.SendCommand = send
    traj = m.runchomp(robot="r", adofgoal=[1, 2], n_iter=10, lambda_=100.0, no_collision_exception=True)
    assert [c.split()[0] for c in m.sent] == ["create", "iterate", "gettraj", "destroy"]
    assert m.sent[1] == "iterate run 0xabc n_iter 10" and m.sent[2] == "gettraj run 0xabc no_collision_exception"
    assert np.array_equal(traj, [[0, 1], [0.5, 1.5], [1, 2]])


def test_module_header_symbols_exported():
    src = open(os.path.join(ROOT, "include", "orcdchomp_b200_module.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", src)))
    assert names == sorted(orcdchomp._MODULE_EXPORTS)
    lib = orcdchomp._lib()
    for n in names:
        assert getattr(lib, n) is not None


def test_tsr_text_round_trip():
    """orcdchomp.TSR.serialize() -> the library's reader of the create command's TSR arguments
    (tsr_create_parse, src/orcdchomp_mod.cpp:3068-3110): rotation by columns, translation, bounds; poses come
    back as [x y z qx qy qz qw] with the quaternion of cd_kin_quat_from_R (largest component positive)"""
    import numpy as np
    from or_cdchomp_b200 import models, orcdchomp
    rng = np.random.default_rng(3)
    for trial in range(20):
        q0 = models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(-3.1, 3.1))
        q1 = models.quat_from_axis_angle(rng.normal(size=3), rng.uniform(-3.1, 3.1))
        T0w, Twe = models.pose_make(rng.uniform(-1, 1, 3), q0), models.pose_make(rng.uniform(-1, 1, 3), q1)
        Bw = rng.uniform(-1, 1, (6, 2))
        Bw[rng.integers(0, 6)] = 0.0
        text = orcdchomp.TSR(T0w, Twe, Bw, manipindex=trial % 3, bodyandlink="NULL").serialize()
        assert len(text.split()) == 38
        a, b, bw = orcdchomp.parse_tsr(text)
        for got, want in ((a, T0w), (b, Twe)):
            assert np.allclose(got[:3], want[:3], atol=0)
            assert min(np.abs(got[3:] - want[3:]).max(), np.abs(got[3:] + want[3:]).max()) < 1e-15 * 8
            assert got[3 + int(np.argmax(np.abs(got[3:])))] > 0
        assert np.array_equal(bw, Bw)
    # 4 x 4 matrices are accepted as well
    M = np.eye(4)
    M[:3, 3] = [0.1, 0.2, 0.3]
    a, b, bw = orcdchomp.parse_tsr(orcdchomp.TSR(M, None, np.zeros((6, 2))).serialize())
    assert np.allclose(a, [0.1, 0.2, 0.3, 0, 0, 0, 1]) and np.allclose(b, [0, 0, 0, 0, 0, 0, 1])
    import pytest
    for bad in ("", "0 NULL 1 2 3", "x NULL " + " ".join(["0"] * 36), "0 NULL " + " ".join(["0"] * 37)):
        with pytest.raises(ValueError):
            orcdchomp.parse_tsr(bad)
