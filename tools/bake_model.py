#!/usr/bin/env python3
"""Bake the reference's URDF robot models into constant tables (include/xarm_model_tables.h).

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU box):

    python tools/bake_model.py            # rewrites include/xarm_model_tables.h
    python tools/bake_model.py --check    # verifies the committed header is up to date

What it restates (PyBullet URDF import; behaviour RECALLED, no Bullet source here):

* fixed joints stay separate bodies dynamically, but a rigidly attached body moves with its nearest moving
  ancestor, so every URDF link becomes a "part" (mass, COM, inertia) owned by one moving link;
* a link without <inertial> gets mass 1 (mass 0 if it is called "world")  [REF xarm7_pd.urdf:241 link_eef];
* loadURDF is called WITHOUT URDF_USE_INERTIA_FROM_FILE [REF xarm_pick_and_place.py:77], so Bullet replaces
  the URDF inertia tensor by the box inertia of the link's collision AABB (convex-hull margin 0.001), taken in
  the inertial frame (URDF inertial origin, rotated to the principal axes of the URDF tensor when that tensor
  is not diagonal).  A link with mass but no collision shape gets zero rotational inertia (point mass).
  `--urdf-inertia` switches to the tensors written in the URDF instead (kept as a calibration switch).

Only numbers (model constants) are emitted; no reference code is copied.
"""
import argparse
import math
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/gym_xarm/envs/urdf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "xarm_model_tables.h")
HULL_MARGIN = 0.001  # gUrdfDefaultCollisionMargin [RECALLED]


def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def fvec(s, n=3, default=0.0):
    if s is None:
        return np.full(n, default)
    return np.array([float(x) for x in s.split()])


def mesh_vertices(path):
    if path.lower().endswith(".obj"):
        v = []
        with open(path) as f:
            for line in f:
                if line.startswith("v "):
                    v.append([float(x) for x in line.split()[1:4]])
        return np.array(v)
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * n == len(data):  # binary STL
        arr = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), count=n, offset=84)
        return arr["v"].reshape(-1, 3).astype(np.float64)
    v = []  # ascii STL
    for line in data.decode("ascii", "ignore").splitlines():
        t = line.split()
        if t and t[0] == "vertex":
            v.append([float(x) for x in t[1:4]])
    return np.array(v)


def jacobi_principal(I, threshold=1e-6, max_steps=30):
    """Jacobi diagonalisation picking the largest off-diagonal element each sweep (btMatrix3x3::diagonalize
    restated from memory); returns (principal values, rotation whose columns are the principal axes)."""
    A = I.copy()
    R = np.eye(3)
    for _ in range(max_steps):
        p, q, r = 0, 1, 2
        mx = abs(A[0, 1])
        if abs(A[0, 2]) > mx:
            p, q, r, mx = 0, 2, 1, abs(A[0, 2])
        if abs(A[1, 2]) > mx:
            p, q, r, mx = 1, 2, 0, abs(A[1, 2])
        t = threshold * (abs(A[0, 0]) + abs(A[1, 1]) + abs(A[2, 2]))
        if mx <= t:
            break
        theta = (A[q, q] - A[p, p]) / (2 * A[p, q])
        tt = (1.0 / (theta + math.copysign(math.sqrt(1 + theta * theta), theta))) if theta != 0 else 1.0
        c = 1 / math.sqrt(1 + tt * tt)
        s = c * tt
        G = np.eye(3)
        G[p, p] = c
        G[q, q] = c
        G[p, q] = s
        G[q, p] = -s
        A = G.T @ A @ G
        R = R @ G
    return np.diag(A).copy(), R


class Link:
    pass


def load_urdf(path, urdf_inertia):
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(path)
    links = {}
    for le in root.findall("link"):
        L = Link()
        L.name = le.get("name")
        ine = le.find("inertial")
        if ine is None:
            L.mass = 0.0 if L.name == "world" else 1.0
            L.com = np.zeros(3)
            L.Rin = np.eye(3)
            L.I_urdf = np.eye(3) if L.mass else np.zeros((3, 3))
            L.has_inertial = False
        else:
            o = ine.find("origin")
            L.com = fvec(o.get("xyz") if o is not None else None)
            L.Rin = rpy_to_R(fvec(o.get("rpy") if o is not None else None))
            L.mass = float(ine.find("mass").get("value"))
            i = ine.find("inertia")
            g = lambda k: float(i.get(k))
            L.I_urdf = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            L.has_inertial = True
        L.collisions = []
        for ce in le.findall("collision"):
            o = ce.find("origin")
            xyz = fvec(o.get("xyz") if o is not None else None)
            R = rpy_to_R(fvec(o.get("rpy") if o is not None else None))
            geo = ce.find("geometry")
            if geo.find("box") is not None:
                half = fvec(geo.find("box").get("size")) / 2
                L.collisions.append(("box", xyz, R, -half, half))
            elif geo.find("mesh") is not None:
                m = geo.find("mesh")
                scale = fvec(m.get("scale"), default=1.0) if m.get("scale") else np.ones(3)
                v = mesh_vertices(os.path.join(base_dir, m.get("filename"))) * scale
                L.collisions.append(("hull", xyz, R, v.min(0) - HULL_MARGIN, v.max(0) + HULL_MARGIN))
        links[L.name] = L
    joints = []
    for je in root.findall("joint"):
        if je.get("type") is None:
            continue
        J = Link()
        J.name = je.get("name")
        J.type = je.get("type")
        J.parent = je.find("parent").get("link")
        J.child = je.find("child").get("link")
        o = je.find("origin")
        J.xyz = fvec(o.get("xyz") if o is not None else None)
        J.R = rpy_to_R(fvec(o.get("rpy") if o is not None else None))
        a = je.find("axis")
        J.axis = fvec(a.get("xyz")) if a is not None else np.array([1.0, 0, 0])
        lim = je.find("limit")
        J.lo = float(lim.get("lower", 0)) if lim is not None else 0.0
        J.hi = float(lim.get("upper", 0)) if lim is not None else 0.0
        dyn = je.find("dynamics")
        J.damping = float(dyn.get("damping", 0)) if dyn is not None else 0.0
        joints.append(J)
    # inertia per link in the LINK frame about its COM
    for L in links.values():
        if L.mass == 0.0:
            L.I_link = np.zeros((3, 3))
            continue
        # inertial frame orientation: URDF rpy, then principal axes when the tensor is not diagonal
        offd = abs(L.I_urdf[0, 1]) + abs(L.I_urdf[0, 2]) + abs(L.I_urdf[1, 2])
        if offd == 0.0:
            princ, Rp = np.diag(L.I_urdf).copy(), np.eye(3)
        else:
            princ, Rp = jacobi_principal(L.I_urdf)
        Rf = L.Rin @ Rp  # inertial frame -> link frame
        if urdf_inertia:
            diag = princ
        elif not L.collisions:
            diag = np.zeros(3)
        else:
            lo = np.full(3, 1e30)
            hi = np.full(3, -1e30)
            for (_, xyz, R, bmin, bmax) in L.collisions:
                # child transform = inertial^-1 * collision origin; AABB of the transformed local AABB
                Rc = Rf.T @ R
                tc = Rf.T @ (xyz - L.com)
                c = Rc @ ((bmin + bmax) / 2) + tc
                e = np.abs(Rc) @ ((bmax - bmin) / 2)
                lo = np.minimum(lo, c - e)
                hi = np.maximum(hi, c + e)
            l = hi - lo
            diag = L.mass / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])
        L.I_link = Rf @ np.diag(diag) @ Rf.T
    return links, joints


def build_model(path, urdf_inertia):
    """Flatten to moving links + rigidly attached parts, in PyBullet joint-index order."""
    links, joints = load_urdf(path, urdf_inertia)
    child_names = {j.child for j in joints}
    rootname = [n for n in links if n not in child_names][0]
    # PyBullet numbers joints depth-first in URDF child order; these URDFs list them so that file order works.
    order = []

    def visit(name):
        for j in joints:
            if j.parent == name:
                order.append(j)
                visit(j.child)

    visit(rootname)
    mov = []  # moving links
    parts = []
    frame_of = {rootname: (-1, np.eye(3), np.zeros(3))}  # link name -> (moving owner, R, t) of link frame in owner frame
    joint_index = {}
    for idx, j in enumerate(order):
        joint_index[j.name] = idx
        owner, Rp, tp = frame_of[j.parent]
        R = Rp @ j.R
        t = tp + Rp @ j.xyz
        if j.type in ("revolute", "prismatic", "continuous"):
            M = Link()
            M.name = j.child
            M.joint = j.name
            M.pb_index = idx
            M.parent = owner
            M.R0, M.t0 = R, t
            M.type = 0 if j.type != "prismatic" else 1
            M.axis = j.axis / np.linalg.norm(j.axis)
            M.lo, M.hi, M.damping = j.lo, j.hi, j.damping
            mov.append(M)
            frame_of[j.child] = (len(mov) - 1, np.eye(3), np.zeros(3))
        else:
            frame_of[j.child] = (owner, R, t)
        L = links[j.child]
        own, Rl, tl = frame_of[j.child]
        if own >= 0 and L.mass > 0:
            P = Link()
            P.name, P.owner, P.mass = L.name, own, L.mass
            P.com = tl + Rl @ L.com
            P.I = Rl @ L.I_link @ Rl.T
            P.pb_index = idx
            parts.append(P)
    # composite per moving link
    for i, M in enumerate(mov):
        ps = [p for p in parts if p.owner == i]
        m = sum(p.mass for p in ps)
        com = sum(p.mass * p.com for p in ps) / m
        I = np.zeros((3, 3))
        for p in ps:
            d = p.com - com
            I += p.I + p.mass * (d @ d * np.eye(3) - np.outer(d, d))
        M.mass, M.com, M.I = m, com, I
        M.Ig = sum(p.I for p in ps)  # sum of the parts' central inertias (no parallel-axis terms)
    return links, joints, mov, parts, frame_of, joint_index


def fmt(x):
    return repr(float(x))


def arr(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return "{" + ", ".join(fmt(x) for x in a) + "}"
    return "{" + ", ".join(arr(r) for r in a) + "}"


def sym6(I):
    return [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]


def emit_model(prefix, mov, parts, extra):
    o = []
    n, npart = len(mov), len(parts)
    o.append(f"#define {prefix}_NDOF {n}")
    o.append(f"#define {prefix}_NPART {npart}")
    o.append(f"#define {prefix}_PARENT {{" + ", ".join(str(m.parent) for m in mov) + "}")
    o.append(f"#define {prefix}_JTYPE {{" + ", ".join(str(m.type) for m in mov) + "}  /* 0 revolute, 1 prismatic */")
    o.append(f"#define {prefix}_PB_INDEX {{" + ", ".join(str(m.pb_index) for m in mov) + "}  /* PyBullet joint index of each DoF */")
    o.append(f"#define {prefix}_R0 " + arr([m.R0.reshape(9) for m in mov]) + "  /* joint frame rotation in parent moving link, row-major */")
    o.append(f"#define {prefix}_T0 " + arr([m.t0 for m in mov]))
    o.append(f"#define {prefix}_AXIS " + arr([m.axis for m in mov]))
    o.append(f"#define {prefix}_LIMIT_LO " + arr([m.lo for m in mov]))
    o.append(f"#define {prefix}_LIMIT_HI " + arr([m.hi for m in mov]))
    o.append(f"#define {prefix}_DAMPING " + arr([m.damping for m in mov]))
    o.append(f"#define {prefix}_MASS " + arr([m.mass for m in mov]) + "  /* composite of the parts a DoF carries */")
    o.append(f"#define {prefix}_COM " + arr([m.com for m in mov]))
    o.append(f"#define {prefix}_INERTIA " + arr([sym6(m.I) for m in mov]) + "  /* xx xy xz yy yz zz about COM, link frame */")
    o.append(f"#define {prefix}_CENTRAL_INERTIA " + arr([sym6(m.Ig) for m in mov]) + "  /* sum of part inertias about their own COMs */")
    o.append(f"#define {prefix}_PART_OWNER {{" + ", ".join(str(p.owner) for p in parts) + "}")
    o.append(f"#define {prefix}_PART_MASS " + arr([p.mass for p in parts]))
    o.append(f"#define {prefix}_PART_COM " + arr([p.com for p in parts]))
    o.append(f"#define {prefix}_PART_INERTIA " + arr([sym6(p.I) for p in parts]))
    o.append(f"/* parts: " + ", ".join(f"{p.name}->dof{p.owner}" for p in parts) + " */")
    for k, v in extra.items():
        o.append(f"#define {prefix}_{k} {v}")
    return o


def generate(urdf_inertia=False):
    out = []
    out.append("/* GENERATED by tools/bake_model.py from the reference's URDF files - do not edit.")
    out.append(" * Model constants only (numbers read from gym_xarm/envs/urdf/{xarm7_pd,xarm7,my_door}.urdf and the")
    out.append(" * AABBs of their collision meshes).  Initialiser-list macros so that the C oracle (double) and the CUDA")
    out.append(" * kernels (float) each declare their own typed tables.  Inertia rule: "
               + ("URDF tensors" if urdf_inertia else "collision-AABB box inertia (PyBullet default)") + ". */")
    out.append("#ifndef XARM_MODEL_TABLES_H")
    out.append("#define XARM_MODEL_TABLES_H")
    out.append("")
    # --- Panda-gripper arm [REF xarm7_pd.urdf] ---
    links, joints, mov, parts, frame_of, jidx = build_model(os.path.join(REF, "xarm7_pd.urdf"), urdf_inertia)
    names = [m.name for m in mov]
    l7 = names.index("link7")
    own, R, t = frame_of["link_eef"]
    assert own == l7
    ownh, Rh, th = frame_of["panda_hand"]
    hand_com = th + Rh @ links["panda_hand"].com
    extra = {
        "EEF_DOF": l7,
        "EEF_POS": arr(t),
        "HAND_COM": arr(hand_com) + "  /* PyBullet link 9 COM in the link7 frame [REF xarm7_pd.urdf:325] */",
        "FINGER1_DOF": names.index("panda_leftfinger"),
        "FINGER2_DOF": names.index("panda_rightfinger"),
    }
    # collider boxes from the collision hull AABBs: finger hull and hand hull as boxes
    def hull_box(link):
        (_, xyz, Rc, bmin, bmax) = links[link].collisions[0]
        bmin = bmin + HULL_MARGIN
        bmax = bmax - HULL_MARGIN
        c = Rc @ ((bmin + bmax) / 2) + xyz
        e = np.abs(Rc) @ ((bmax - bmin) / 2)
        return c, e
    c1, e1 = hull_box("panda_leftfinger")
    c2, e2 = hull_box("panda_rightfinger")
    ch, eh = hull_box("panda_hand")
    ch = th + Rh @ ch
    extra["FINGER1_BOX_C"] = arr(c1)
    extra["FINGER1_BOX_H"] = arr(e1)
    extra["FINGER2_BOX_C"] = arr(c2)
    extra["FINGER2_BOX_H"] = arr(e2)
    extra["HAND_BOX_C"] = arr(ch) + "  /* in the link7 frame */"
    extra["HAND_BOX_H"] = arr(eh)
    out.append("/* ---- xArm7 + Panda hand: 7 revolute + 2 prismatic DoF [REF gym_xarm/envs/urdf/xarm7_pd.urdf] ---- */")
    out += emit_model("XARM_PD", mov, parts, extra)
    out.append("")
    # --- xArm-gripper arm [REF xarm7.urdf] ---
    links, joints, mov, parts, frame_of, jidx = build_model(os.path.join(REF, "xarm7.urdf"), urdf_inertia)
    names = [m.name for m in mov]
    l7 = names.index("link7")
    own, R, t = frame_of["link_eef"]
    ownh, Rh, th = frame_of["xarm_gripper_base_link"]
    hand_com = th + Rh @ links["xarm_gripper_base_link"].com
    extra = {
        "EEF_DOF": l7,
        "EEF_POS": arr(t),
        "HAND_COM": arr(hand_com) + "  /* PyBullet link 9 COM in the link7 frame [REF xarm7.urdf:365] */",
        "DRIVE_DOF": names.index("left_outer_knuckle"),
    }
    out.append("/* ---- xArm7 + xArm gripper: 7 + 6 revolute DoF [REF gym_xarm/envs/urdf/xarm7.urdf] ---- */")
    out += emit_model("XARM_XG", mov, parts, extra)
    out.append("")
    # --- door [REF my_door.urdf] ---
    links, joints, mov, parts, frame_of, jidx = build_model(os.path.join(REF, "my_door.urdf"), urdf_inertia)
    assert len(mov) == 1
    D = mov[0]
    out.append("/* ---- door: two fixed bars + one sliding bar [REF gym_xarm/envs/urdf/my_door.urdf:1-110] ---- */")
    out.append("#define XARM_DOOR_AXIS " + arr(D.axis))
    out.append("#define XARM_DOOR_ORIGIN " + arr(D.t0))
    out.append("#define XARM_DOOR_LIMIT_LO " + fmt(D.lo))
    out.append("#define XARM_DOOR_LIMIT_HI " + fmt(D.hi))
    out.append("#define XARM_DOOR_DAMPING " + fmt(D.damping))
    out.append("#define XARM_DOOR_MASS " + fmt(D.mass))
    out.append("#define XARM_DOOR_BAR_HALF " + arr(links["doorLink"].collisions[0][4]))
    f1 = frame_of["doorFrameLink1"][2]
    f2 = frame_of["doorFrameLink2"][2]
    out.append("#define XARM_DOOR_FIXED_BAR1 " + arr(f1))
    out.append("#define XARM_DOOR_FIXED_BAR2 " + arr(f2))
    out.append("#define XARM_DOOR_FRICTION 10.0  /* <lateral_friction> [REF my_door.urdf:5,39,73] */")
    out.append("")
    out.append("#endif /* XARM_MODEL_TABLES_H */")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--urdf-inertia", action="store_true")
    a = ap.parse_args()
    text = generate(a.urdf_inertia)
    if a.check:
        cur = open(OUT).read()
        sys.exit(0 if cur == text else 1)
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", os.path.normpath(OUT), len(text), "bytes")
