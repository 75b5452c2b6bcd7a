#!/usr/bin/env python3
"""Offline model compiler: MJCF scene + Panda model -> flat constant tables.

Reads the two MJCF files the reference loads (``mujoco_manip/env.py:15-69`` resolves
``pick_and_place_scene.xml`` + the included ``franka_emika_panda/panda.xml``) with plain
``xml.etree`` (MuJoCo itself is not installed in this image) and emits

* ``mujoco_manip_b200/csrc/model_gen.h``  - C tables shared (as *data*) by the CUDA kernels and
  the CPU oracle,
* ``mujoco_manip_b200/model.json``        - the same numbers for the Python host side.

Only what the step hot path needs is compiled: kinematic tree, explicit inertials, joint / dof
parameters, position actuators, the fixed tendon + joint equality of the gripper, collision
geoms (boxes, plane, cylinders, convex hulls of the collision meshes), the candidate geom-pair
list after MuJoCo's filters, the ``scene_start`` keyframe and the compile-time constants MuJoCo
derives at ``qpos0`` (``dof_invweight0``, ``body_invweight0``, ``meaninertia``) [SURVEY App. A4].

Run:  python tools/modelc.py [/root/reference/mujoco_manip/data]
The generated files are committed; the GPU box never needs the reference tree.
"""
from __future__ import annotations

import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# MuJoCo geom type enum order (decides which geom is "geom1" in a pair)
GT_PLANE, GT_CYLINDER, GT_BOX, GT_MESH = 0, 5, 6, 7
JT_FREE, JT_SLIDE, JT_HINGE = 0, 2, 3


def fl(s, n=None):
    v = [float(x) for x in s.split()]
    if n is not None:
        assert len(v) == n, (s, n)
    return v


def qnorm(q):
    q = np.asarray(q, float)
    return q / np.linalg.norm(q)


def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def q2R(q):
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


# ---------------------------------------------------------------------------------------------
# mesh loading
# ---------------------------------------------------------------------------------------------
def load_mesh_vertices(path):
    if path.lower().endswith(".stl"):
        with open(path, "rb") as f:
            data = f.read()
        ntri = struct.unpack_from("<I", data, 80)[0]
        assert len(data) == 84 + 50 * ntri, "ASCII STL not supported"
        v = np.zeros((ntri * 3, 3))
        for t in range(ntri):
            vals = struct.unpack_from("<12f", data, 84 + 50 * t)
            v[3 * t : 3 * t + 3] = np.array(vals[3:12]).reshape(3, 3)
        return v
    verts = []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
    return np.array(verts)


def hull_vertices(v):
    v = np.unique(np.round(v.astype(np.float64), 9), axis=0)
    h = ConvexHull(v)
    idx = np.sort(h.vertices)
    return v[idx]


# ---------------------------------------------------------------------------------------------
# MJCF parsing (just what these two files use)
# ---------------------------------------------------------------------------------------------
class Defaults:
    """Nested <default class=...> tables for joint / geom / general."""

    def __init__(self):
        self.cls = {"main": {"joint": {}, "geom": {}, "general": {}, "parent": None}}

    def load(self, node, parent="main"):
        name = node.get("class", "main" if parent is None else None)
        if name is None:
            name = "main"
        if name not in self.cls:
            self.cls[name] = {"joint": {}, "geom": {}, "general": {}, "parent": parent}
        for tag in ("joint", "geom", "general"):
            for el in node.findall(tag):
                self.cls[name][tag].update(el.attrib)
        for sub in node.findall("default"):
            self.load(sub, name)

    def resolve(self, tag, cname):
        chain = []
        c = cname or "main"
        while c is not None:
            chain.append(c)
            c = self.cls[c]["parent"]
        out = {}
        for c in reversed(chain):
            out.update(self.cls[c][tag])
        return out


def compile_model(data_dir):
    scene_path = os.path.join(data_dir, "pick_and_place_scene.xml")
    panda_dir = os.path.join(data_dir, "franka_emika_panda")
    scene = ET.parse(scene_path).getroot()
    panda = ET.parse(os.path.join(panda_dir, "panda.xml")).getroot()
    mesh_dir = os.path.join(panda_dir, panda.find("compiler").get("meshdir", ""))

    opt = {"timestep": 0.002, "gravity": [0, 0, -9.81], "integrator": "Euler"}
    for root in (scene, panda):
        o = root.find("option")
        if o is not None:
            if o.get("timestep"):
                opt["timestep"] = float(o.get("timestep"))
            if o.get("gravity"):
                opt["gravity"] = fl(o.get("gravity"), 3)
            if o.get("integrator"):
                opt["integrator"] = o.get("integrator")
    assert opt["integrator"] == "implicitfast"

    dfl = Defaults()
    for root in (panda, scene):
        d = root.find("default")
        if d is not None:
            # the outer <default> is the unnamed main class
            for sub in d.findall("default"):
                dfl.load(sub, "main")
            for tag in ("joint", "geom", "general"):
                for el in d.findall(tag):
                    dfl.cls["main"][tag].update(el.attrib)

    meshes = {}
    for m in panda.find("asset").findall("mesh"):
        f = m.get("file")
        name = m.get("name") or os.path.splitext(os.path.basename(f))[0]
        meshes[name] = os.path.join(mesh_dir, f)

    bodies = [
        dict(name="world", parent=-1, pos=[0, 0, 0], quat=[1, 0, 0, 0], mass=0.0, ipos=[0, 0, 0],
             inertia=np.zeros((3, 3)), joint=-1)
    ]
    joints, geoms = [], []

    def add_geom(el, bid, childclass):
        cname = el.get("class") or childclass
        a = dfl.resolve("geom", cname)
        a.update(el.attrib)
        contype = int(a.get("contype", 1))
        conaff = int(a.get("conaffinity", 1))
        if contype == 0 and conaff == 0:
            return  # visual only; all robot bodies carry explicit inertials
        gtype = a.get("type", "sphere")
        g = dict(
            name=a.get("name", ""), body=bid, contype=contype, conaffinity=conaff,
            pos=fl(a.get("pos", "0 0 0"), 3), quat=list(qnorm(fl(a.get("quat", "1 0 0 0"), 4))),
            friction=fl(a.get("friction", "1 0.005 0.0001"), 3), condim=int(a.get("condim", 3)),
            size=[0.0, 0.0, 0.0], mesh="", mass=float(a.get("mass", 0) or 0), gtype_name=gtype,
        )
        if gtype == "box":
            g["type"] = GT_BOX
            g["size"] = fl(a["size"], 3)
            g["rbound"] = float(np.linalg.norm(g["size"]))
        elif gtype == "plane":
            g["type"] = GT_PLANE
            g["size"] = fl(a["size"], 3)
            g["rbound"] = 0.0
        elif gtype == "cylinder":
            g["type"] = GT_CYLINDER
            s = fl(a["size"])
            g["size"] = [s[0], s[1], 0.0]
            g["rbound"] = float(np.hypot(s[0], s[1]))
        elif gtype == "mesh":
            g["type"] = GT_MESH
            g["mesh"] = a["mesh"]
        else:
            raise ValueError(f"unsupported geom type {gtype}")
        geoms.append(g)

    def add_body(el, parent, childclass):
        childclass = el.get("childclass") or childclass
        b = dict(
            name=el.get("name", ""), parent=parent, pos=fl(el.get("pos", "0 0 0"), 3),
            quat=list(qnorm(fl(el.get("quat", "1 0 0 0"), 4))), mass=0.0, ipos=[0, 0, 0],
            inertia=np.zeros((3, 3)), joint=-1,
        )
        bid = len(bodies)
        bodies.append(b)
        ine = el.find("inertial")
        if ine is not None:
            b["mass"] = float(ine.get("mass"))
            b["ipos"] = fl(ine.get("pos", "0 0 0"), 3)
            if ine.get("fullinertia"):
                xx, yy, zz, xy, xz, yz = fl(ine.get("fullinertia"), 6)
                b["inertia"] = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
            else:
                b["inertia"] = np.diag(fl(ine.get("diaginertia"), 3))
            assert ine.get("quat") is None
        jels = el.findall("joint") + el.findall("freejoint")
        assert len(jels) <= 1
        for j in jels:
            if j.tag == "freejoint":
                jd = dict(name=j.get("name", ""), type=JT_FREE, body=bid, axis=[0, 0, 1],
                          range=[0.0, 0.0], limited=0, armature=0.0, damping=0.0)
            else:
                a = dfl.resolve("joint", j.get("class") or childclass)
                a.update(j.attrib)
                jt = {"hinge": JT_HINGE, "slide": JT_SLIDE}[a.get("type", "hinge")]
                rng = fl(a.get("range", "0 0"), 2)
                jd = dict(name=a.get("name", ""), type=jt, body=bid,
                          axis=list(qnorm(fl(a.get("axis", "0 0 1"), 3))), range=rng,
                          limited=int(rng[0] < rng[1]),  # autolimits="true"
                          armature=float(a.get("armature", 0)), damping=float(a.get("damping", 0)))
                assert a.get("pos") is None and a.get("ref") is None
            b["joint"] = len(joints)
            joints.append(jd)
        for g in el.findall("geom"):
            add_geom(g, bid, childclass)
        for c in el.findall("body"):
            add_body(c, bid, childclass)

    # include precedes the scene's own worldbody (SURVEY 2.1 body order)
    for root in (panda, scene):
        wb = root.find("worldbody")
        for g in wb.findall("geom"):
            add_geom(g, 0, None)
        for c in wb.findall("body"):
            add_body(c, 0, None)

    # free-body inertia from its single box geom (inertiafromgeom=auto, no <inertial>)
    for g in geoms:
        b = bodies[g["body"]]
        if b["joint"] >= 0 and joints[b["joint"]]["type"] == JT_FREE:
            assert g["type"] == GT_BOX and np.allclose(g["pos"], 0)
            m = g["mass"]
            sx, sy, sz = g["size"]
            b["mass"] = m
            b["inertia"] = np.diag([m / 3 * (sy * sy + sz * sz), m / 3 * (sx * sx + sz * sz),
                                    m / 3 * (sx * sx + sy * sy)])

    # qpos / dof addresses
    nq = nv = 0
    for j in joints:
        j["qposadr"], j["dofadr"] = nq, nv
        if j["type"] == JT_FREE:
            nq, nv = nq + 7, nv + 6
        else:
            nq, nv = nq + 1, nv + 1

    # weld ids (bodies without joints are welded to their parent)
    for i, b in enumerate(bodies):
        if i == 0:
            b["weld"] = 0
        elif b["joint"] < 0:
            b["weld"] = bodies[b["parent"]]["weld"]
        else:
            b["weld"] = i
    for i, b in enumerate(bodies):
        w = b["weld"]
        b["weldparent"] = 0 if w == 0 else bodies[bodies[w]["parent"]]["weld"]

    # collision hulls
    hull_off, hull_pts = {}, []
    for g in geoms:
        if g["type"] != GT_MESH:
            g["vadr"], g["vnum"] = 0, 0
            continue
        if g["mesh"] not in hull_off:
            hv = hull_vertices(load_mesh_vertices(meshes[g["mesh"]]))
            hull_off[g["mesh"]] = (len(hull_pts), len(hv))
            hull_pts.extend(hv.tolist())
        g["vadr"], g["vnum"] = hull_off[g["mesh"]]
        hv = np.array(hull_pts[g["vadr"] : g["vadr"] + g["vnum"]])
        # bounding sphere about the vertex-box centre, expressed as geom-frame offset
        c = 0.5 * (hv.min(0) + hv.max(0))
        g["bcenter"] = c.tolist()
        g["rbound"] = float(np.linalg.norm(hv - c, axis=1).max())
        g["size"] = (0.5 * (hv.max(0) - hv.min(0))).tolist()
    for g in geoms:
        g.setdefault("bcenter", [0.0, 0.0, 0.0])

    # excludes
    name2body = {b["name"]: i for i, b in enumerate(bodies)}
    excludes = set()
    for root in (panda, scene):
        c = root.find("contact")
        if c is not None:
            for e in c.findall("exclude"):
                a, b_ = name2body[e.get("body1")], name2body[e.get("body2")]
                excludes.add((min(a, b_), max(a, b_)))

    # candidate geom pairs after MuJoCo's static filters
    pairs = []
    for i in range(len(geoms)):
        for k in range(i + 1, len(geoms)):
            g1, g2 = geoms[i], geoms[k]
            b1, b2 = g1["body"], g2["body"]
            if b1 == b2:
                continue
            if not ((g1["contype"] & g2["conaffinity"]) or (g2["contype"] & g1["conaffinity"])):
                continue
            w1, w2 = bodies[b1]["weld"], bodies[b2]["weld"]
            if w1 == w2:
                continue
            if w1 != 0 and w2 != 0 and (w1 == bodies[b2]["weldparent"] or w2 == bodies[b1]["weldparent"]):
                continue
            if (min(b1, b2), max(b1, b2)) in excludes:
                continue
            a, b_ = (i, k) if g1["type"] <= g2["type"] else (k, i)
            pairs.append((a, b_))

    # actuators
    acts = []
    for a_el in panda.find("actuator").findall("general"):
        a = dfl.resolve("general", a_el.get("class"))
        a.update(a_el.attrib)
        gain = fl(a.get("gainprm", "1"))[0]
        bias = (fl(a.get("biasprm", "0 0 0")) + [0, 0, 0])[:3]
        act = dict(name=a.get("name"), gain=gain, bias=bias, ctrlrange=fl(a["ctrlrange"], 2),
                   forcerange=fl(a["forcerange"], 2))
        if a.get("joint"):
            jn = [j["name"] for j in joints].index(a["joint"])
            act["trntype"], act["trnid"] = 0, joints[jn]["dofadr"]
        else:
            act["trntype"], act["trnid"] = 1, 0
        acts.append(act)

    ten = panda.find("tendon").find("fixed")
    tendon = [([j["name"] for j in joints].index(e.get("joint")), float(e.get("coef")))
              for e in ten.findall("joint")]
    tendon = [(joints[j]["dofadr"], c) for j, c in tendon]

    eq_el = panda.find("equality").find("joint")
    jn = [j["name"] for j in joints]
    eq = dict(dof1=joints[jn.index(eq_el.get("joint1"))]["dofadr"],
              dof2=joints[jn.index(eq_el.get("joint2"))]["dofadr"],
              solref=fl(eq_el.get("solref"), 2),
              solimp=(fl(eq_el.get("solimp")) + [0.5, 2.0])[:5])
    if len(fl(eq_el.get("solimp"))) == 3:
        eq["solimp"] = fl(eq_el.get("solimp")) + [0.5, 2.0]

    key = None
    for k in scene.find("keyframe").findall("key"):
        if k.get("name") == "scene_start":
            key = dict(qpos=fl(k.get("qpos"), nq), ctrl=fl(k.get("ctrl"), len(acts)))

    model = dict(opt=opt, bodies=bodies, joints=joints, geoms=geoms, pairs=pairs, actuators=acts,
                 tendon=tendon, equality=eq, key=key, nq=nq, nv=nv, hull=hull_pts)
    derive_constants(model)
    return model


# ---------------------------------------------------------------------------------------------
# compile-time constants at qpos0  (numpy CRB; independent of the oracle's C++ implementation)
# ---------------------------------------------------------------------------------------------
def fk_qpos0(model):
    """World poses of all bodies at qpos0 (all joint coordinates zero, free bodies at body pos)."""
    X = []
    for i, b in enumerate(model["bodies"]):
        if i == 0:
            X.append((np.zeros(3), np.eye(3)))
            continue
        pp, pR = X[b["parent"]]
        X.append((pp + pR @ np.array(b["pos"]), pR @ q2R(b["quat"])))
    return X


def mass_matrix_qpos0(model):
    """Dense joint-space inertia at qpos0 from body Jacobians: M = sum_b Jb^T I_b Jb (+armature)."""
    nv = model["nv"]
    X = fk_qpos0(model)
    bodies, joints = model["bodies"], model["joints"]

    def body_jac(bid, point):
        Jp, Jr = np.zeros((3, nv)), np.zeros((3, nv))
        b = bid
        while b > 0:
            j = bodies[b]["joint"]
            if j >= 0:
                jd = joints[j]
                p, R = X[b]
                d = jd["dofadr"]
                if jd["type"] == JT_HINGE:
                    ax = R @ np.array(jd["axis"])
                    Jr[:, d] = ax
                    Jp[:, d] = np.cross(ax, point - p)
                elif jd["type"] == JT_SLIDE:
                    Jp[:, d] = R @ np.array(jd["axis"])
                else:  # free: world-frame linear velocity, body-frame angular velocity
                    Jp[:, d : d + 3] = np.eye(3)
                    for k in range(3):
                        Jr[:, d + 3 + k] = R[:, k]
                        Jp[:, d + 3 + k] = np.cross(R[:, k], point - p)
            b = bodies[b]["parent"]
        return Jp, Jr

    M = np.zeros((nv, nv))
    jacs = {}
    for i, b in enumerate(bodies):
        if i == 0:
            continue
        p, R = X[i]
        com = p + R @ np.array(b["ipos"])
        Jp, Jr = body_jac(i, com)
        jacs[i] = (Jp, Jr)
        if b["mass"] > 0:
            Iw = R @ np.asarray(b["inertia"]) @ R.T
            M += b["mass"] * Jp.T @ Jp + Jr.T @ Iw @ Jr
    for j in joints:
        if j["type"] != JT_FREE:
            M[j["dofadr"], j["dofadr"]] += j["armature"]
    return M, jacs


def derive_constants(model):
    nv = model["nv"]
    M, jacs = mass_matrix_qpos0(model)
    Minv = np.linalg.inv(M)
    dof_inv = np.diag(Minv).copy()
    for j in model["joints"]:
        if j["type"] == JT_FREE:
            d = j["dofadr"]
            dof_inv[d : d + 3] = dof_inv[d : d + 3].mean()
            dof_inv[d + 3 : d + 6] = dof_inv[d + 3 : d + 6].mean()
    body_inv = np.zeros((len(model["bodies"]), 2))
    for i, b in enumerate(model["bodies"]):
        if i == 0 or b["weld"] == 0:
            continue
        Jp, Jr = jacs[i]
        body_inv[i, 0] = np.trace(Jp @ Minv @ Jp.T) / 3
        body_inv[i, 1] = np.trace(Jr @ Minv @ Jr.T) / 3
    model["dof_invweight0"] = dof_inv.tolist()
    model["body_invweight0"] = body_inv.tolist()
    model["meaninertia"] = float(np.trace(M) / nv)
    model["M0"] = M.tolist()


# ---------------------------------------------------------------------------------------------
# emit
# ---------------------------------------------------------------------------------------------
def carr(name, ctype, arr, fmt="%.17g"):
    a = np.asarray(arr)
    dims = "".join(f"[{d}]" for d in a.shape)
    flat = a.reshape(-1)
    if ctype == "int":
        body = ", ".join(str(int(x)) for x in flat)
    else:
        body = ", ".join(fmt % float(x) for x in flat)
    return f"MM_CONST {ctype} {name}{dims} = {{{body}}};\n"


def emit_header(model, path):
    B, J, G = model["bodies"], model["joints"], model["geoms"]
    nb, nj, ng, nu = len(B), len(J), len(G), len(model["actuators"])
    out = []
    out.append("// GENERATED by tools/modelc.py from the reference's MJCF scene - do not edit.\n")
    out.append("// Model data only (numbers derived from pick_and_place_scene.xml + panda.xml + hulls).\n")
    out.append("#pragma once\n#ifndef MM_CONST\n#define MM_CONST static const\n#endif\n")
    out.append(f"#define MM_NBODY {nb}\n#define MM_NJNT {nj}\n#define MM_NGEOM {ng}\n#define MM_NQ {model['nq']}\n"
               f"#define MM_NV {model['nv']}\n#define MM_NU {nu}\n#define MM_NPAIR {len(model['pairs'])}\n"
               f"#define MM_NHULLV {len(model['hull'])}\n")
    out.append(f"#define MM_TIMESTEP {model['opt']['timestep']!r}\n")
    out.append(f"#define MM_MEANINERTIA {model['meaninertia']!r}\n")
    out.append(carr("mm_gravity", "double", model["opt"]["gravity"]))
    out.append(carr("mm_body_parent", "int", [b["parent"] for b in B]))
    out.append(carr("mm_body_weld", "int", [b["weld"] for b in B]))
    out.append(carr("mm_body_jnt", "int", [b["joint"] for b in B]))
    out.append(carr("mm_body_pos", "double", [b["pos"] for b in B]))
    out.append(carr("mm_body_quat", "double", [b["quat"] for b in B]))
    out.append(carr("mm_body_mass", "double", [b["mass"] for b in B]))
    out.append(carr("mm_body_ipos", "double", [b["ipos"] for b in B]))
    out.append(carr("mm_body_inertia", "double", [np.asarray(b["inertia"]).reshape(9) for b in B]))
    out.append(carr("mm_body_invweight0", "double", model["body_invweight0"]))
    out.append(carr("mm_jnt_type", "int", [j["type"] for j in J]))
    out.append(carr("mm_jnt_body", "int", [j["body"] for j in J]))
    out.append(carr("mm_jnt_qposadr", "int", [j["qposadr"] for j in J]))
    out.append(carr("mm_jnt_dofadr", "int", [j["dofadr"] for j in J]))
    out.append(carr("mm_jnt_limited", "int", [j["limited"] for j in J]))
    out.append(carr("mm_jnt_axis", "double", [j["axis"] for j in J]))
    out.append(carr("mm_jnt_range", "double", [j["range"] for j in J]))
    out.append(carr("mm_jnt_armature", "double", [j["armature"] for j in J]))
    out.append(carr("mm_jnt_damping", "double", [j["damping"] for j in J]))
    out.append(carr("mm_dof_invweight0", "double", model["dof_invweight0"]))
    out.append(carr("mm_geom_type", "int", [g["type"] for g in G]))
    out.append(carr("mm_geom_body", "int", [g["body"] for g in G]))
    out.append(carr("mm_geom_condim", "int", [g["condim"] for g in G]))
    out.append(carr("mm_geom_vadr", "int", [g["vadr"] for g in G]))
    out.append(carr("mm_geom_vnum", "int", [g["vnum"] for g in G]))
    out.append(carr("mm_geom_pos", "double", [g["pos"] for g in G]))
    out.append(carr("mm_geom_quat", "double", [g["quat"] for g in G]))
    out.append(carr("mm_geom_size", "double", [g["size"] for g in G]))
    out.append(carr("mm_geom_friction", "double", [g["friction"] for g in G]))
    out.append(carr("mm_geom_bcenter", "double", [g["bcenter"] for g in G]))
    out.append(carr("mm_geom_rbound", "double", [g["rbound"] for g in G]))
    out.append(carr("mm_hull", "double", model["hull"]))
    out.append(carr("mm_pair", "int", model["pairs"]))
    A = model["actuators"]
    out.append(carr("mm_act_gain", "double", [a["gain"] for a in A]))
    out.append(carr("mm_act_bias", "double", [a["bias"] for a in A]))
    out.append(carr("mm_act_ctrlrange", "double", [a["ctrlrange"] for a in A]))
    out.append(carr("mm_act_forcerange", "double", [a["forcerange"] for a in A]))
    out.append(carr("mm_act_trntype", "int", [a["trntype"] for a in A]))
    out.append(carr("mm_act_trnid", "int", [a["trnid"] for a in A]))
    out.append(carr("mm_tendon_dof", "int", [t[0] for t in model["tendon"]]))
    out.append(carr("mm_tendon_coef", "double", [t[1] for t in model["tendon"]]))
    e = model["equality"]
    out.append(f"#define MM_EQ_DOF1 {e['dof1']}\n#define MM_EQ_DOF2 {e['dof2']}\n")
    out.append(carr("mm_eq_solref", "double", e["solref"]))
    out.append(carr("mm_eq_solimp", "double", e["solimp"]))
    out.append(carr("mm_key_qpos", "double", model["key"]["qpos"]))
    out.append(carr("mm_key_ctrl", "double", model["key"]["ctrl"]))
    names = ", ".join('"%s"' % b["name"] for b in B)
    out.append(f"#ifdef MM_WANT_NAMES\nstatic const char* const mm_body_name[{nb}] = {{{names}}};\n")
    names = ", ".join('"%s"' % (g["name"] or g["mesh"] or "pad") for g in G)
    out.append(f"static const char* const mm_geom_name[{ng}] = {{{names}}};\n")
    names = ", ".join('"%s"' % j["name"] for j in J)
    out.append(f"static const char* const mm_jnt_name[{nj}] = {{{names}}};\n#endif\n")
    with open(path, "w") as f:
        f.write("".join(out))



# ---------------------------------------------------------------------------------------------
# specialised tables for the CUDA kernels (structure-exploiting layout)
# ---------------------------------------------------------------------------------------------
def emit_dev_header(model, path):
    """Tables in the kernels' own numbering.

    dynamic bodies db: 0..6 link1..link7, 7 hand, 8 left finger, 9 right finger, 10..12 cubes.
    inertial bodies ib = joints 0..8: link1..6, link7+hand (merged rigidly), left, right finger.
    chain classes: 0 static, 1..7 link1..7 (hand == 7), 8 left finger, 9 right finger, 10..12 cubes.
    boxes (oracle geom order so that geom1/geom2 roles agree): 0..9 pads, 10 tabletop, 11..25 bins,
    26..28 cubes; id 29 = floor plane.
    """
    B, J, G = model["bodies"], model["joints"], model["geoms"]
    name2b = {b["name"]: i for i, b in enumerate(B)}
    chain = ["link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand", "left_finger", "right_finger"]
    link_pos = [B[name2b[n]]["pos"] for n in chain]
    link_R = [q2R(B[name2b[n]]["quat"]).reshape(9) for n in chain]
    # merged link7 + hand
    b7, bh = B[name2b["link7"]], B[name2b["hand"]]
    Rh = q2R(bh["quat"])
    ch = np.array(bh["pos"]) + Rh @ np.array(bh["ipos"])
    c7 = np.array(b7["ipos"])
    m = b7["mass"] + bh["mass"]
    c = (b7["mass"] * c7 + bh["mass"] * ch) / m

    def pax(mass, d):
        return mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))

    I = np.asarray(b7["inertia"]) + pax(b7["mass"], c7 - c) + Rh @ np.asarray(bh["inertia"]) @ Rh.T + pax(bh["mass"], ch - c)
    ib_names = ["link1", "link2", "link3", "link4", "link5", "link6", "link7", "left_finger", "right_finger"]
    ib_mass, ib_com, ib_I = [], [], []
    for n in ib_names:
        b = B[name2b[n]]
        if n == "link7":
            mm_, cc, II = m, c, I
        else:
            mm_, cc, II = b["mass"], np.array(b["ipos"]), np.asarray(b["inertia"])
        ib_mass.append(mm_)
        ib_com.append(cc)
        ib_I.append([II[0, 0], II[1, 1], II[2, 2], II[0, 1], II[0, 2], II[1, 2]])
    # boxes
    db_of_body = {name2b[n]: i for i, n in enumerate(chain)}
    for k, n in enumerate(["obj_red", "obj_green", "obj_blue"]):
        db_of_body[name2b[n]] = 10 + k
    cls_of_body = {name2b[n]: min(i + 1, 7) for i, n in enumerate(chain[:8])}
    cls_of_body[name2b["left_finger"]] = 8
    cls_of_body[name2b["right_finger"]] = 9
    for k, n in enumerate(["obj_red", "obj_green", "obj_blue"]):
        cls_of_body[name2b[n]] = 10 + k
    X = fk_qpos0(model)
    box_ids = [i for i, g in enumerate(G) if g["type"] == GT_BOX]
    assert len(box_ids) == 29
    gid2box = {g: k for k, g in enumerate(box_ids)}
    plane_gid = [i for i, g in enumerate(G) if g["type"] == GT_PLANE][0]
    gid2box[plane_gid] = 29
    box_size, box_pos, box_body, box_class, box_invw, box_cube = [], [], [], [], [], []
    for gi in box_ids:
        g = G[gi]
        b = g["body"]
        assert np.allclose(g["quat"], [1, 0, 0, 0])
        box_size.append(g["size"])
        if B[b]["weld"] == 0:  # static: world position (bodies are axis aligned)
            p, R = X[b]
            assert np.allclose(R, np.eye(3))
            box_pos.append((p + np.array(g["pos"])).tolist())
            box_body.append(-1)
            box_class.append(0)
        else:
            box_pos.append(g["pos"])
            box_body.append(db_of_body[b])
            box_class.append(cls_of_body[b])
        box_invw.append(model["body_invweight0"][b][0])
        box_cube.append(int(g["condim"] == 4))
    # candidate list restricted to plane-box / box-box, sorted by (classA, classB) so that contacts
    # of one body pair are contiguous (the kernels reduce per body pair)
    cand = []
    for g1, g2 in model["pairs"]:
        t1, t2 = G[g1]["type"], G[g2]["type"]
        if (t1, t2) not in ((GT_PLANE, GT_BOX), (GT_BOX, GT_BOX)):
            continue
        a, b_ = gid2box[g1], gid2box[g2]
        ca = 0 if a == 29 else box_class[a]
        cb = box_class[b_]
        cand.append((ca, cb, a, b_))
    cand.sort()
    # ---- unified geom table (all 47 collision geoms, oracle numbering) and the full candidate list ----
    g_type, g_body, g_class, g_cube, g_obst, g_size, g_pos, g_bc, g_rb, g_invw, g_vadr, g_vnum = ([] for _ in range(12))
    robot_names = set(chain) | {"link0"}
    for gi, g in enumerate(G):
        b = g["body"]
        assert np.allclose(g["quat"], [1, 0, 0, 0]), "kernels assume geom frames aligned with their body"
        static = B[b]["weld"] == 0
        if static:
            p, R = X[b]
            assert np.allclose(R, np.eye(3))
            g_pos.append((p + np.array(g["pos"])).tolist())
            g_bc.append((p + np.array(g["pos"]) + np.array(g["bcenter"])).tolist())
            g_body.append(-1)
            g_class.append(0)
        else:
            g_pos.append(list(g["pos"]))
            g_bc.append((np.array(g["pos"]) + np.array(g["bcenter"])).tolist())
            g_body.append(db_of_body[b])
            g_class.append(cls_of_body[b])
        g_type.append(g["type"])
        g_cube.append(int(g["condim"] == 4))
        nm = B[b]["name"]
        g_obst.append(int(static and nm not in ("world", "link0")))  # gym_env.py:137-152
        g_size.append(list(g["size"]))
        g_rb.append(g["rbound"])
        g_invw.append(model["body_invweight0"][b][0])
        g_vadr.append(g["vadr"])
        g_vnum.append(g["vnum"])
    pairs = sorted((g_class[a], g_class[b_], a, b_) for a, b_ in model["pairs"])
    out = []
    out.append("// GENERATED by tools/modelc.py - specialised tables for the CUDA kernels. Do not edit.\n#pragma once\n")
    out.append(f"#define MMD_NBOX 29\n#define MMD_PLANE 29\n#define MMD_NCAND {len(cand)}\n")
    out.append(f"#define MMD_NGEOM {len(G)}\n#define MMD_NPAIR {len(pairs)}\n#define MMD_NHULLV {len(model['hull'])}\n")
    out.append(carr("mmd_link_pos", "double", link_pos))
    out.append(carr("mmd_link_R", "double", link_R))
    out.append(carr("mmd_ib_mass", "double", ib_mass))
    out.append(carr("mmd_ib_com", "double", ib_com))
    out.append(carr("mmd_ib_inertia", "double", ib_I))
    out.append(carr("mmd_g_type", "int", g_type))
    out.append(carr("mmd_g_body", "int", g_body))
    out.append(carr("mmd_g_class", "int", g_class))
    out.append(carr("mmd_g_cube", "int", g_cube))
    out.append(carr("mmd_g_obst", "int", g_obst))
    out.append(carr("mmd_g_vadr", "int", g_vadr))
    out.append(carr("mmd_g_vnum", "int", g_vnum))
    out.append(carr("mmd_g_size", "double", g_size))
    out.append(carr("mmd_g_pos", "double", g_pos))
    out.append(carr("mmd_g_bc", "double", g_bc))
    out.append(carr("mmd_g_rbound", "double", g_rb))
    out.append(carr("mmd_g_invw", "double", g_invw))
    out.append(carr("mmd_pair", "int", [[c[2], c[3]] for c in pairs]))
    with open(path, "w") as f:
        f.write("".join(out))
    return pairs

def main():
    data_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/mujoco_manip/data"
    model = compile_model(data_dir)
    emit_header(model, os.path.join(REPO, "mujoco_manip_b200", "csrc", "model_gen.h"))
    cand = emit_dev_header(model, os.path.join(REPO, "mujoco_manip_b200", "csrc", "model_dev_gen.h"))
    print("device candidate pairs:", len(cand))

    def clean(o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, dict):
            return {k: clean(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [clean(v) for v in o]
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        return o

    slim = clean({k: v for k, v in model.items() if k not in ("hull", "M0")})
    with open(os.path.join(REPO, "mujoco_manip_b200", "model.json"), "w") as f:
        json.dump(slim, f, indent=0)
    print(f"nbody={len(model['bodies'])} njnt={len(model['joints'])} ngeom={len(model['geoms'])} "
          f"nq={model['nq']} nv={model['nv']} npair={len(model['pairs'])} hullv={len(model['hull'])}")
    print("meaninertia", model["meaninertia"])
    print("dof_invweight0", np.round(model["dof_invweight0"], 4))
    for i, b in enumerate(model["bodies"]):
        print(i, b["name"], "weld", b["weld"], "invw", np.round(model["body_invweight0"][i], 4))
    for i, g in enumerate(model["geoms"]):
        print(i, g["name"] or g["mesh"], "body", g["body"], "type", g["type"], "nv", g["vnum"],
              "rbound %.4f" % g["rbound"])


if __name__ == "__main__":
    main()
