"""the OpenRAVE-side adapters.

* include/orcdchomp_b200_openrave.h (chain extractor, written against OpenRAVE's public API) is
  compiled against a declaration-only stand-in (include/openrave_min/) and run on a random tree
  robot that implements those accessors (tests/cpp/adapter_probe.cpp).  The extracted struct
  ocb_robot -- what the kernels evaluate -- must place every link exactly where OpenRAVE's own rule
  T_child = T_parent * Left * motion(axis, value) * Right places it, for random active-dof vectors.
* the <orcdchomp><spheres> block of a robot XML file is read by the library's reader
  (ocb_kdata_parse_spheres, the element / attribute handling of src/orcdchomp_kdata.cpp:65-98).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from or_cdchomp_b200 import capi, models


def _run_probe(tmp_path, seed):
    exe = os.path.join(str(tmp_path), "adapter_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include", "openrave_min"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_probe.cpp"),
                           "-o", exe])
    return subprocess.check_output([exe, str(seed)], text=True).splitlines()


@pytest.mark.parametrize("seed", [7, 19])
def test_chain_extractor_matches_openrave_rule(oracle, tmp_path, seed):
    lines = _run_probe(tmp_path, seed)
    head = dict(zip(lines[0].split()[0::2], [int(x) for x in lines[0].split()[1::2]]))
    links = [ln.split()[1:] for ln in lines if ln.startswith("link ")]
    assert len(links) == head["n_links"] == 2 * head["n_or_links"] - 1   # a joint frame + a link frame per non-root link
    parent = [int(f[0]) for f in links]
    assert all(p < i for i, p in enumerate(parent))
    base = [float(x) for x in next(ln for ln in lines if ln.startswith("base")).split()[1:]]
    lim = [float(x) for x in next(ln for ln in lines if ln.startswith("limits")).split()[1:]]
    sph = [ln.split()[1:] for ln in lines if ln.startswith("sphere ")]
    link_map = [int(x) for x in next(ln for ln in lines if ln.startswith("map")).split()[1:]]
    robot = capi.RobotDesc(
        names=["l%d" % i for i in range(len(links))], parent=parent,
        pose_parent=[[float(x) for x in f[5:12]] for f in links], joint_type=[int(f[1]) for f in links],
        axis=[[float(x) for x in f[12:15]] for f in links], dof_index=[int(f[2]) for f in links],
        dof_coeff=[[float(f[3]), float(f[4])] for f in links], base_pose=base,
        limit_lower=lim[0::2], limit_upper=lim[1::2],
        sphere_link=[int(f[0]) for f in sph], sphere_pos=[[float(x) for x in f[1:4]] for f in sph],
        sphere_radius=[float(f[4]) for f in sph])
    assert robot.n_dof == head["n_dof"] == 4 and robot.n_spheres == 6
    # the probe's active dofs are 5, 3, 2, 0 of dofs with limits [-2-d, 1.5+d]
    assert np.allclose(robot.limit_lower, [-7, -5, -4, -2]) and np.allclose(robot.limit_upper, [6.5, 4.5, 3.5, 1.5])
    assert all(int(f[0]) in link_map for f in sph)
    # forward kinematics through the extracted description == OpenRAVE's rule
    i = 0
    trials = 0
    while i < len(lines):
        if lines[i].startswith("q"):
            q = np.array([float(x) for x in lines[i].split()[1:]])
            world = {}
            i += 1
            while i < len(lines) and lines[i].startswith("world"):
                f = lines[i].split()
                world[int(f[1])] = np.array([float(x) for x in f[2:]])
                i += 1
            poses = oracle.fk(robot, q)
            for li, w in world.items():
                p = poses[link_map[li]]
                assert np.allclose(p[:3], w[:3], atol=1e-13), (li, p, w)
                # same rotation: quaternions equal up to sign
                assert min(np.abs(p[3:] - w[3:]).max(), np.abs(p[3:] + w[3:]).max()) < 1e-13, (li, p, w)
            trials += 1
        else:
            i += 1
    assert trials == 3
    # ocb_or::tsr_constraint: the manipulator's end effector (OpenRAVE link 5) with its tool transform, a bare link (3)
    cons = [ln.split()[1:] for ln in lines if ln.startswith("con ")]
    tool = [float(x) for x in next(ln for ln in lines if ln.startswith("tool")).split()[1:]]
    assert [int(cons[0][0]), int(cons[0][1])] == [capi.CON_ALL, link_map[5]]
    assert np.allclose([float(x) for x in cons[0][2:]], tool, atol=0)
    assert [int(cons[1][0]), int(cons[1][1])] == [capi.CON_START, link_map[3]]
    assert [float(x) for x in cons[1][2:]] == [0, 0, 0, 0, 0, 0, 1]


def test_sphere_table_xml_reader():
    table = models.wam7_spheres()
    assert len(table) == 16 and table[0] == ("wam0", (0.22, 0.14, 0.346), 0.15)
    assert [t[0] for t in table[10:]] == ["Finger0-1", "Finger1-1", "Finger2-1", "Finger0-2", "Finger1-2", "Finger2-2"]
    xml = """<?xml version="1.0"?><Robot><!-- <sphere link="commented" pos="1 1 1" radius="1"/> -->
      <sphere link="outside" pos="0 0 0" radius="1"/>
      <KinBody><orcdchomp><spheres>
         <sphere link='a b' radius="0.5"  pos=" 1  -2.5 3e-1 " />
         <sphere link="c" pos="0 0 0" radius="0.25"></sphere>
      </spheres><sphere link="stray" pos="0 0 0" radius="1"/></orcdchomp></KinBody></Robot>"""
    assert models.parse_spheres_xml(xml) == [("a b", (1.0, -2.5, 0.3), 0.5), ("c", (0.0, 0.0, 0.0), 0.25)]
    with pytest.raises(ValueError, match="unknown attribute colour=red"):
        models.parse_spheres_xml('<orcdchomp><spheres><sphere link="a" colour="red"/></spheres></orcdchomp>')
    with pytest.raises(ValueError, match="inside <spheres>"):
        models.parse_spheres_xml("<orcdchomp><spheres><spheres></spheres></spheres></orcdchomp>")
    assert models.parse_spheres_xml("<Robot/>") == []
