"""Golden vectors for the robot/v1 model tables from the reference's OWN second description of the robot.

The reference ships robot/v1 twice, exported independently from the same CAD: the URDF that TSID loads
(ref:robot/v1/urdf/robot_mod.urdf, ref:ctrl/conf.py:16) and the MuJoCo model that main.py simulates
(ref:robot/v1/mujoco/robot.xml, ref:main.py:12).  This script reads the MuJoCo file only (plain XML, no mujoco
package needed), evaluates the kinematic tree for seeded joint configurations and stores frame-independent
whole-body quantities in the torso frame: total mass, centre of mass, rotational inertia about the centre of
mass.  tests/test_oracle.py compares the oracle's model (compiled from the URDF) with them — a check of
joint order/sign, joint placements, masses, levers and inertia tensors against reference-held data.

Run in the build container (reads /root/reference):  python tests/golden/make_mjcf_golden.py
"""
import json
import os
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/robot/v1/mujoco/robot.xml"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mjcf_v1.json")


def quat_R(w, x, y, z):
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_axis(axis, a):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


def floats(s, n, default):
    return np.array([float(t) for t in s.split()]) if s is not None else np.array(default, float)


def read_tree(path):
    """Bodies in document order: (name, parent index, pos, R, hinge name | None, axis, mass, com, I_com)."""
    root = ET.parse(path).getroot()
    bodies = []

    def visit(elem, parent):
        pos = floats(elem.get("pos"), 3, [0, 0, 0])
        R = quat_R(*floats(elem.get("quat"), 4, [1, 0, 0, 0]))
        j = elem.find("joint")
        ine = elem.find("inertial")
        fi = floats(ine.get("fullinertia"), 6, [0] * 6)  # xx yy zz xy xz yz, about the inertial pos, body axes
        I = np.array([[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]])
        assert ine.get("quat") is None
        bodies.append(dict(name=elem.get("name"), parent=parent, pos=pos, R=R,
                           joint=j.get("name") if j is not None else None,
                           axis=floats(j.get("axis"), 3, [0, 0, 1]) if j is not None else None,
                           mass=float(ine.get("mass")), com=floats(ine.get("pos"), 3, [0, 0, 0]), I=I))
        me = len(bodies) - 1
        for c in elem.findall("body"):
            visit(c, me)

    tops = root.find("worldbody").findall("body")
    assert len(tops) == 1 and tops[0].find("freejoint") is not None
    visit(tops[0], -1)
    return bodies


def whole_body(bodies, qj):
    """Total mass, CoM and rotational inertia about the CoM in the frame of the first (free-flying) body."""
    Rw, pw = [], []
    for b in bodies:
        if b["parent"] < 0:
            R, p = np.eye(3), np.zeros(3)  # the torso frame itself (its pos/quat place it in the world)
        else:
            Rp, pp = Rw[b["parent"]], pw[b["parent"]]
            R = Rp @ b["R"]
            p = pp + Rp @ b["pos"]
            if b["joint"] is not None:
                R = R @ rot_axis(b["axis"], qj[b["joint"]])
        Rw.append(R)
        pw.append(p)
    m = sum(b["mass"] for b in bodies)
    c = sum(b["mass"] * (pw[i] + Rw[i] @ b["com"]) for i, b in enumerate(bodies)) / m
    Ic = np.zeros((3, 3))
    for i, b in enumerate(bodies):
        d = pw[i] + Rw[i] @ b["com"] - c
        Ic += Rw[i] @ b["I"] @ Rw[i].T + b["mass"] * (d @ d * np.eye(3) - np.outer(d, d))
    return m, c, Ic, Rw, pw


def main():
    bodies = read_tree(REF)
    names = [b["joint"] for b in bodies if b["joint"] is not None]
    rng = np.random.default_rng(20261019)
    cases = []
    for k in range(12):
        q = {n: (0.0 if k == 0 else float(rng.uniform(-0.9, 0.9))) for n in names}
        m, c, Ic, Rw, pw = whole_body(bodies, q)
        feet = {b["joint"]: dict(R=Rw[i].tolist(), p=pw[i].tolist()) for i, b in enumerate(bodies)
                if b["joint"] in ("left_ankle_roll", "right_ankle_roll")}
        cases.append(dict(q=q, mass=m, com=c.tolist(), inertia_com=Ic.tolist(), foot_bodies=feet))
    out = dict(source="robot/v1/mujoco/robot.xml (reference), evaluated by tests/golden/make_mjcf_golden.py",
               joint_names=names,
               bodies=[dict(name=b["name"], joint=b["joint"], parent_joint=bodies[b["parent"]]["joint"] if b["parent"] >= 0 else None,
                            pos=b["pos"].tolist(), R=b["R"].tolist(), mass=b["mass"], com=b["com"].tolist(), inertia=b["I"].tolist())
                       for b in bodies],
               cases=cases)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(OUT, "bodies", len(bodies), "mass", cases[0]["mass"])


if __name__ == "__main__":
    main()
