"""Synthetic robots and scenes for the BASELINE.json configurations.

OpenRAVE's stock models (robots/wam7.kinbody.xml, barretthand.kinbody.xml, the
table / mug meshes; scripts/test_wam7.py:23-35 in the reference) are not part of
the reference repository, so the kinematic tree is stated here explicitly.  The
sphere table is the reference's own (scripts/barrettwam_withspheres.robot.xml:
24-45), the start configuration and CHOMP parameters are the demo's
(scripts/test_wam7.py:66,83-84).  The tree below IS the ground truth both the
CUDA engine and the CPU oracle consume; it is "a WAM-like 7-dof arm with a
3-finger hand", not a calibrated Barrett model.
"""
import math
import os

import numpy as np

from .capi import (JOINT_FIXED, JOINT_PRISMATIC, JOINT_REVOLUTE, RobotDesc)

# --------------------------------------------------------------------------
# pose helpers (libcd pose = [x y z qx qy qz qw], src/libcd/kin.c:42-52)


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = math.sin(0.5 * angle)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(0.5 * angle)])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz])


def quat_rotate(q, v):
    qx, qy, qz, qw = q
    x, y, z = v
    return np.array([
        x * (qx * qx - qy * qy - qz * qz + qw * qw) + 2 * y * (qx * qy - qz * qw) + 2 * z * (qx * qz + qy * qw),
        2 * x * (qx * qy + qz * qw) + y * (-qx * qx + qy * qy - qz * qz + qw * qw) + 2 * z * (qy * qz - qx * qw),
        2 * x * (qx * qz - qy * qw) + 2 * y * (qy * qz + qx * qw) + z * (-qx * qx - qy * qy + qz * qz + qw * qw)])


def pose_make(xyz=(0, 0, 0), quat=(0, 0, 0, 1)):
    return np.array(list(xyz) + list(quat), dtype=np.float64)


def pose_compose(ab, bc):
    """cd_kin_pose_compose semantics (src/libcd/kin.c:136-178)."""
    q = quat_mul(ab[3:], bc[3:])
    t = quat_rotate(ab[3:], bc[:3]) + ab[:3]
    return np.concatenate([t, q])


def pose_invert(p):
    q = np.array([-p[3], -p[4], -p[5], p[6]])
    t = -quat_rotate(q, p[:3])
    return np.concatenate([t, q])


# --------------------------------------------------------------------------
# WAM7 + BarrettHand-like model

WAM7_LIMITS = np.array([
    [-2.6, 2.6], [-2.0, 2.0], [-2.8, 2.8], [-0.9, 3.1],
    [-4.76, 1.24], [-1.6, 1.6], [-3.0, 3.0]])

# scripts/test_wam7.py:66
WAM7_DEMO_START = np.array([2.5, -1.8, 0.0, 2.0, 0.0, 0.2, 0.0])
# the demo's goal comes from ikfast (not reproducible offline); fixed stand-in
WAM7_DEMO_GOAL = np.array([0.6, 0.9, 0.3, 1.4, -0.4, 0.5, 0.8])

def parse_spheres_xml(text):
    """The <orcdchomp><spheres> block of a robot XML text -> [(link, (x, y, z), radius), ...] in
    document order, read by the library's own reader (ocb_kdata_parse_spheres: the element and
    attribute handling of src/orcdchomp_kdata.cpp:65-98 in the reference)."""
    import ctypes as C
    from . import capi
    lib = capi.load_library()
    fn = lib.ocb_kdata_parse_spheres
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.c_int, C.c_char_p, capi.c_double_p, capi.c_double_p, capi.c_int_p, C.c_char_p, C.c_size_t]
    cap = 512
    names = C.create_string_buffer(cap * 64)
    pos = np.zeros((cap, 3))
    rad = np.zeros(cap)
    n = C.c_int()
    err = C.create_string_buffer(256)
    rc = fn(text.encode() if isinstance(text, str) else text, cap, names, capi.dptr(pos), capi.dptr(rad),
            C.byref(n), err, len(err))
    if rc:
        raise ValueError(err.value.decode())
    out = []
    for k in range(min(n.value, cap)):
        link = names.raw[k * 64:(k + 1) * 64].split(b"\0", 1)[0].decode()
        out.append((link, (float(pos[k, 0]), float(pos[k, 1]), float(pos[k, 2])), float(rad[k])))
    return out


def wam7_spheres():
    """The WAM + BarrettHand sphere model, parsed from the robot XML's <orcdchomp> block
    (data/barrettwam_spheres.robot.xml; values of scripts/barrettwam_withspheres.robot.xml:24-45)."""
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "data", "barrettwam_spheres.robot.xml")) as f:
        return parse_spheres_xml(f.read())


def wam7_robot(base_pose=None, hand_values=(0.5, 0.5, 0.5, 0.3)):
    """17-link tree: wam0..wam7 (7 active revolute dofs, all frames axis-aligned
    at q=0 with z along the arm, axes z y z y z y z), a fixed hand base and a
    3-finger hand whose 4 dofs are INACTIVE (frozen at hand_values =
    (curl0, curl1, curl2, spread)), as when the demo activates only the arm
    (scripts/test_wam7.py:47)."""
    if base_pose is None:
        base_pose = pose_make()
    c0, c1, c2, spread = hand_values
    rx90 = quat_from_axis_angle((1, 0, 0), math.pi / 2)
    rz180 = quat_from_axis_angle((0, 0, 1), math.pi)
    links = []  # (name, parent, pose_parent, type, axis, dof, coeff)

    def add(name, parent, xyz, quat, jtype, axis, dof, coeff):
        links.append((name, parent, pose_make(xyz, quat), jtype, axis, dof, coeff))

    I = (0, 0, 0, 1)
    add("wam0", -1, (0, 0, 0), I, JOINT_FIXED, (0, 0, 1), -1, (0, 0))
    add("wam1", 0, (0.22, 0.14, 0.346), I, JOINT_REVOLUTE, (0, 0, 1), 0, (1, 0))
    add("wam2", 1, (0, 0, 0), I, JOINT_REVOLUTE, (0, 1, 0), 1, (1, 0))
    add("wam3", 2, (0.045, 0, 0.55), I, JOINT_REVOLUTE, (0, 0, 1), 2, (1, 0))
    add("wam4", 3, (-0.045, 0, 0), I, JOINT_REVOLUTE, (0, 1, 0), 3, (1, 0))
    add("wam5", 4, (0, 0, 0.3), I, JOINT_REVOLUTE, (0, 0, 1), 4, (1, 0))
    add("wam6", 5, (0, 0, 0), I, JOINT_REVOLUTE, (0, 1, 0), 5, (1, 0))
    add("wam7", 6, (0, 0, 0.06), I, JOINT_REVOLUTE, (0, 0, 1), 6, (1, 0))
    add("handbase", 7, (0, 0, 0), I, JOINT_FIXED, (0, 0, 1), -1, (0, 0))
    # finger 0 (spread +), finger 1 (spread -), finger 2 (no spread, opposing)
    add("Finger0-0", 8, (0.0, -0.025, 0.0415), I, JOINT_REVOLUTE, (0, 0, 1), -1, (0, spread))
    add("Finger0-1", 9, (0.05, 0, 0.0339), rx90, JOINT_REVOLUTE, (0, 0, 1), -1, (0, c0))
    add("Finger0-2", 10, (0.07, 0, 0), I, JOINT_REVOLUTE, (0, 0, 1), -1, (0, c0 / 3.0 + 0.7))
    add("Finger1-0", 8, (0.0, 0.025, 0.0415), I, JOINT_REVOLUTE, (0, 0, 1), -1, (0, -spread))
    add("Finger1-1", 12, (0.05, 0, 0.0339), rx90, JOINT_REVOLUTE, (0, 0, 1), -1, (0, c1))
    add("Finger1-2", 13, (0.07, 0, 0), I, JOINT_REVOLUTE, (0, 0, 1), -1, (0, c1 / 3.0 + 0.7))
    add("Finger2-1", 8, (-0.05, 0, 0.0754), quat_mul(rz180, rx90), JOINT_REVOLUTE, (0, 0, 1), -1, (0, c2))
    add("Finger2-2", 15, (0.07, 0, 0), I, JOINT_REVOLUTE, (0, 0, 1), -1, (0, c2 / 3.0 + 0.7))

    names = [l[0] for l in links]
    spheres = wam7_spheres()
    return RobotDesc(
        names=names,
        parent=[l[1] for l in links],
        pose_parent=[l[2] for l in links],
        joint_type=[l[3] for l in links],
        axis=[l[4] for l in links],
        dof_index=[l[5] for l in links],
        dof_coeff=[l[6] for l in links],
        base_pose=base_pose,
        limit_lower=WAM7_LIMITS[:, 0], limit_upper=WAM7_LIMITS[:, 1],
        sphere_link=[names.index(s[0]) for s in spheres],
        sphere_pos=[s[1] for s in spheres],
        sphere_radius=[s[2] for s in spheres])


def dense_sphere_arm(n_spheres=200, seed=5):
    """BASELINE config 5: the 7-dof arm (no hand) carrying `n_spheres` spheres
    spread over its 7 moving links, radii U[0.02, 0.06]."""
    rng = np.random.default_rng(seed)
    base = wam7_robot()
    keep = 8  # wam0..wam7
    link_len = {1: 0.05, 2: 0.55, 3: 0.05, 4: 0.3, 5: 0.05, 6: 0.06, 7: 0.12}
    sl, sp, sr = [], [], []
    for i in range(n_spheres):
        link = 1 + (i % 7)
        z = rng.uniform(0.0, link_len[link])
        xy = rng.uniform(-0.03, 0.03, size=2)
        sl.append(link)
        sp.append((xy[0], xy[1], z))
        sr.append(rng.uniform(0.02, 0.06))
    return RobotDesc(
        names=base.names[:keep], parent=base.parent[:keep], pose_parent=base.pose_parent[:keep],
        joint_type=base.joint_type[:keep], axis=base.axis[:keep], dof_index=base.dof_index[:keep],
        dof_coeff=base.dof_coeff[:keep], base_pose=base.base_pose,
        limit_lower=WAM7_LIMITS[:, 0], limit_upper=WAM7_LIMITS[:, 1],
        sphere_link=sl, sphere_pos=sp, sphere_radius=sr)


def prismatic_test_robot():
    """Small tree with a prismatic joint, a mimic joint and a branch; only used by
    the parity tests to exercise the general kinematics path."""
    I = (0, 0, 0, 1)
    ry = quat_from_axis_angle((0, 1, 0), 0.4)
    links = [
        ("base", -1, pose_make(), JOINT_FIXED, (0, 0, 1), -1, (0, 0)),
        ("slide", 0, pose_make((0.1, 0.0, 0.2)), JOINT_PRISMATIC, (0, 0, 1), 0, (1, 0)),
        ("yaw", 1, pose_make((0, 0, 0.1), ry), JOINT_REVOLUTE, (0, 0, 1), 1, (1, 0)),
        ("armA", 2, pose_make((0.2, 0, 0)), JOINT_REVOLUTE, (0, 1, 0), 2, (1, 0.1)),
        ("armA2", 3, pose_make((0.25, 0, 0)), JOINT_REVOLUTE, (0, 1, 0), 2, (-0.5, 0.2)),  # mimic of dof 2
        ("armB", 2, pose_make((-0.2, 0, 0.05)), JOINT_REVOLUTE, (1, 0, 0), 3, (1, 0)),
        ("tool", 5, pose_make((0, 0.2, 0)), JOINT_FIXED, (0, 0, 1), -1, (0, 0)),
    ]
    names = [l[0] for l in links]
    spheres = [("base", (0, 0, 0.1), 0.1), ("yaw", (0.1, 0, 0), 0.05), ("armA", (0.12, 0, 0), 0.05),
               ("armA2", (0.1, 0, 0), 0.04), ("armA2", (0.2, 0, 0), 0.04), ("armB", (0, 0.1, 0), 0.05),
               ("tool", (0, 0.05, 0.02), 0.04), ("slide", (0, 0, 0), 0.06)]
    return RobotDesc(
        names=names, parent=[l[1] for l in links], pose_parent=[l[2] for l in links],
        joint_type=[l[3] for l in links], axis=[l[4] for l in links],
        dof_index=[l[5] for l in links], dof_coeff=[l[6] for l in links],
        base_pose=pose_make((0.05, -0.02, 0.0), quat_from_axis_angle((0, 0, 1), 0.3)),
        limit_lower=[-0.2, -2.5, -1.5, -2.0], limit_upper=[0.5, 2.5, 1.5, 2.0],
        sphere_link=[names.index(s[0]) for s in spheres],
        sphere_pos=[s[1] for s in spheres], sphere_radius=[s[2] for s in spheres])


# --------------------------------------------------------------------------
# scenes


def field_geometry(aabb_pos, aabb_extents, cube_extent=0.02, aabb_padding=0.2):
    """Grid sizing of computedistancefield (src/orcdchomp_mod.cpp:386-410):
    sizes = ceil((extent+padding)/cube_extent), lengths = sizes*2*cube_extent,
    grid origin (kinbody frame) = aabb.pos - lengths/2, identity rotation."""
    sizes = [int(math.ceil((aabb_extents[i] + aabb_padding) / cube_extent)) for i in range(3)]
    lengths = [sizes[i] * 2.0 * cube_extent for i in range(3)]
    pose = pose_make([aabb_pos[i] - 0.5 * lengths[i] for i in range(3)])
    return sizes, lengths, pose


def table_scene():
    """Config 1/2: one table slab kinbody.  Returns (kinbody pose in world,
    primitives in the kinbody frame, aabb pos, aabb half extents)."""
    kin_pose = pose_make((0.75, 0.14, 0.30), quat_from_axis_angle((0, 0, 1), 0.3))
    half = (0.4, 0.6, 0.02)
    prims = [("box", pose_make(), half)]
    return kin_pose, prims, (0.0, 0.0, 0.0), half


def clutter_scene(n_boxes=64, n_balls=32, seed=3, half_span=1.8):
    """Config 3: 64 random boxes (half of them rotated) + 32 spheres inside a
    3.6 m cube; primitives in the kinbody frame, AABB = the cube."""
    rng = np.random.default_rng(seed)
    prims = []
    for i in range(n_boxes):
        half = rng.uniform(0.05, 0.25, size=3)
        c = rng.uniform(-half_span + 0.45, half_span - 0.45, size=3)
        if i % 2:
            axis = rng.normal(size=3)
            q = quat_from_axis_angle(axis, rng.uniform(0, math.pi))
        else:
            q = np.array([0, 0, 0, 1.0])
        prims.append(("box", pose_make(c, q), tuple(half)))
    for i in range(n_balls):
        r = rng.uniform(0.05, 0.3)
        c = rng.uniform(-half_span + 0.35, half_span - 0.35, size=3)
        prims.append(("sphere", tuple(c), r))
    return prims, (0.0, 0.0, 0.0), (half_span, half_span, half_span)


def prims_to_grid_frame(prims, pose_grid_in_kinbody):
    """Express kinbody-frame primitives in the grid frame (the frame the
    occupancy cube is axis-aligned in, mod.cpp:507-518)."""
    inv = pose_invert(pose_grid_in_kinbody)
    out = []
    for p in prims:
        if p[0] == "box":
            out.append(("box", pose_compose(inv, np.asarray(p[1], dtype=np.float64)), p[2]))
        elif p[0] == "tri":
            out.append(("tri",) + tuple(tuple(quat_rotate(inv[3:], np.asarray(v, dtype=np.float64)) + inv[:3])
                                        for v in p[1:4]))
        else:
            c = quat_rotate(inv[3:], np.asarray(p[1], dtype=np.float64)) + inv[:3]
            out.append(("sphere", tuple(c), p[2]))
    return out


def random_endpoints(robot, n_runs, seed0=20260217, shrink=0.05):
    """Config 2: start / goal uniform in the joint limits shrunk by 5 %, one
    numpy Generator per run seeded seed0 + run."""
    lo = robot.limit_lower
    hi = robot.limit_upper
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * (1.0 - shrink)
    starts = np.empty((n_runs, robot.n_dof))
    goals = np.empty((n_runs, robot.n_dof))
    for r in range(n_runs):
        rng = np.random.default_rng(seed0 + r)
        starts[r] = mid + half * rng.uniform(-1, 1, size=robot.n_dof)
        goals[r] = mid + half * rng.uniform(-1, 1, size=robot.n_dof)
    return starts, goals


def icosphere_mesh(centre, radius, subdivisions=2):
    """Triangle mesh of a sphere (subdivided icosahedron): list of ('tri', v0, v1, v2)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, new_faces = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    c = np.asarray(centre, dtype=np.float64)
    return [("tri", tuple(c + radius * verts[a]), tuple(c + radius * verts[b]), tuple(c + radius * verts[d]))
            for a, b, d in faces]


def box_mesh(pose, half):
    """12 triangles of an oriented box given as (pose7, half extents)."""
    pose = np.asarray(pose, dtype=np.float64)
    corners = {}
    for sx in (-1, 1):
        for sy in (-1, 1):
            for sz in (-1, 1):
                local = np.array([sx * half[0], sy * half[1], sz * half[2]])
                corners[(sx, sy, sz)] = tuple(quat_rotate(pose[3:], local) + pose[:3])
    tris = []
    for ax in range(3):
        for sgn in (-1, 1):
            o = [a for a in range(3) if a != ax]
            def key(u, v):
                k = [0, 0, 0]
                k[ax], k[o[0]], k[o[1]] = sgn, u, v
                return tuple(k)
            a, b, c, d = corners[key(-1, -1)], corners[key(1, -1)], corners[key(1, 1)], corners[key(-1, 1)]
            tris += [("tri", a, b, c), ("tri", a, c, d)]
    return tris


def mesh_scene(seed=11, half_span=0.8):
    """A small triangle-mesh kinbody: two icospheres and three rotated box meshes."""
    rng = np.random.default_rng(seed)
    prims = []
    for _ in range(2):
        prims += icosphere_mesh(rng.uniform(-half_span + 0.3, half_span - 0.3, size=3), rng.uniform(0.1, 0.25), 2)
    for _ in range(3):
        q = quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, math.pi))
        prims += box_mesh(pose_make(rng.uniform(-half_span + 0.3, half_span - 0.3, size=3), q),
                          rng.uniform(0.05, 0.2, size=3))
    return prims, (0.0, 0.0, 0.0), (half_span, half_span, half_span)
