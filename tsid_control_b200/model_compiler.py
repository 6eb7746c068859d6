"""Offline model compiler: URDF (+SRDF) -> flat kinematic/inertial tables.

The reference builds its model at run time through
``tsid.RobotWrapper(urdf, [root], pin.JointModelFreeFlyer(), False)``
(ref:ctrl/WalkController.py:13-18, ref:legacy/biped.py:10-15) and reads the
"standing" configuration with ``pin.loadReferenceConfigurations``
(ref:ctrl/WalkController.py:22-23).  Neither Pinocchio nor urdfdom exist on the
GPU, so this module restates what those parsers produce [UPSTREAM pinocchio
urdf parser, SURVEY.md A1] and emits plain tables that are frozen into
constant memory by the CUDA library:

* body 0 is the free-flyer (``root_joint``) carrying the root link;
* bodies 1..na are the revolute joints in Pinocchio order: depth-first, a
  link's child joints visited in joint-name order (urdfdom keeps joints in a
  name-sorted map);
* a link behind a *fixed* joint is merged into the body of its parent joint
  (``Y_parent += placement.act(Y_child)``) and the fixed joint becomes an
  operational frame of that name (the sole frames of conf.lf_fixed_joint /
  conf.rf_fixed_joint, ref:ctrl/conf.py:17-18, ref:legacy/op3_conf.py:56-57);
* ``<origin rpy>`` goes through urdfdom's half-angle quaternion (normalised)
  and Eigen's quaternion->matrix formula, so the truncated literals in the
  files (1.5708, 3.14159) are kept as they are, not snapped.

Nothing in here runs on the hot path.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

MAX_BODIES = 24  # must match TSIDB_MAX_BODIES in include/tsidb.h


# ----------------------------------------------------------------------------
# small SE3 / inertia helpers (numpy, init-time only)
# ----------------------------------------------------------------------------
def rpy_to_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """urdfdom Rotation::setFromRPY -> quaternion -> Eigen matrix."""
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / s, y / s, z / s, w / s
    return quat_to_matrix(x, y, z, w)


def quat_to_matrix(x: float, y: float, z: float, w: float) -> np.ndarray:
    """Eigen::Quaternion::toRotationMatrix (no renormalisation)."""
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [1.0 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1.0 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
        ]
    )


def skew(p: np.ndarray) -> np.ndarray:
    return np.array([[0.0, -p[2], p[1]], [p[2], 0.0, -p[0]], [-p[1], p[0], 0.0]])


@dataclass
class SE3:
    R: np.ndarray = field(default_factory=lambda: np.eye(3))
    p: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def __mul__(self, o: "SE3") -> "SE3":
        return SE3(self.R @ o.R, self.p + self.R @ o.p)


@dataclass
class Inertia:
    """Spatial inertia as Pinocchio stores it: mass, lever, 3x3 about the CoM."""

    m: float = 0.0
    c: np.ndarray = field(default_factory=lambda: np.zeros(3))
    I: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))

    def se3_action(self, M: SE3) -> "Inertia":
        return Inertia(self.m, M.p + M.R @ self.c, M.R @ self.I @ M.R.T)

    def __add__(self, o: "Inertia") -> "Inertia":
        # pinocchio InertiaTpl::__plus__
        mab = self.m + o.m
        if mab == 0.0:
            return Inertia()
        ab = self.c - o.c
        sk = skew(ab)
        return Inertia(
            mab,
            (self.m * self.c + o.m * o.c) / mab,
            self.I + o.I - (self.m * o.m / mab) * (sk @ sk),
        )


# ----------------------------------------------------------------------------
# URDF reading
# ----------------------------------------------------------------------------
def _floats(s: Optional[str], n: int) -> List[float]:
    if s is None:
        return [0.0] * n
    v = [float(t) for t in s.split()]
    assert len(v) == n, s
    return v


def _origin(elem: Optional[ET.Element]) -> SE3:
    if elem is None:
        return SE3()
    o = elem.find("origin")
    if o is None:
        return SE3()
    xyz = _floats(o.get("xyz"), 3)
    rpy = _floats(o.get("rpy"), 3)
    return SE3(rpy_to_matrix(*rpy), np.array(xyz))


def _link_inertia(link: ET.Element) -> Inertia:
    ine = link.find("inertial")
    if ine is None:
        return Inertia()
    M = _origin(ine)
    m = float(ine.find("mass").get("value"))
    e = ine.find("inertia")
    ixx, ixy, ixz = (float(e.get(k)) for k in ("ixx", "ixy", "ixz"))
    iyy, iyz, izz = (float(e.get(k)) for k in ("iyy", "iyz", "izz"))
    I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
    return Inertia(m, M.p.copy(), M.R @ I @ M.R.T)


@dataclass
class CompiledModel:
    name: str
    joint_names: List[str]  # revolute joints, Pinocchio order (body 1..na)
    parent: List[int]  # per body, -1 for the floating base
    jR: np.ndarray  # [nb,3,3] joint placement in the parent body frame
    jp: np.ndarray  # [nb,3]
    mass: np.ndarray  # [nb]
    com: np.ndarray  # [nb,3]
    inertia: np.ndarray  # [nb,3,3] about the CoM, body frame
    frames: Dict[str, dict]  # fixed-joint frames: {"body": b, "R": 3x3, "p": 3}
    effort: np.ndarray  # [na]
    velocity: np.ndarray  # [na]
    lower: np.ndarray
    upper: np.ndarray
    q_ref: Dict[str, np.ndarray]  # SRDF group_state name -> q (nq)

    @property
    def nb(self) -> int:
        return len(self.parent)

    @property
    def na(self) -> int:
        return self.nb - 1

    @property
    def nv(self) -> int:
        return self.na + 6

    @property
    def nq(self) -> int:
        return self.na + 7

    @property
    def total_mass(self) -> float:
        return float(self.mass.sum())

    # -- (de)serialisation ---------------------------------------------------
    def to_json(self) -> str:
        d = {
            "name": self.name,
            "joint_names": self.joint_names,
            "parent": self.parent,
            "jR": self.jR.tolist(),
            "jp": self.jp.tolist(),
            "mass": self.mass.tolist(),
            "com": self.com.tolist(),
            "inertia": self.inertia.tolist(),
            "frames": {
                k: {"body": v["body"], "R": np.asarray(v["R"]).tolist(), "p": np.asarray(v["p"]).tolist()}
                for k, v in self.frames.items()
            },
            "effort": self.effort.tolist(),
            "velocity": self.velocity.tolist(),
            "lower": self.lower.tolist(),
            "upper": self.upper.tolist(),
            "q_ref": {k: v.tolist() for k, v in self.q_ref.items()},
        }
        return json.dumps(d, indent=1)

    @staticmethod
    def from_json(text: str) -> "CompiledModel":
        d = json.loads(text)
        return CompiledModel(
            name=d["name"],
            joint_names=list(d["joint_names"]),
            parent=[int(p) for p in d["parent"]],
            jR=np.array(d["jR"], dtype=np.float64),
            jp=np.array(d["jp"], dtype=np.float64),
            mass=np.array(d["mass"], dtype=np.float64),
            com=np.array(d["com"], dtype=np.float64),
            inertia=np.array(d["inertia"], dtype=np.float64),
            frames={
                k: {"body": int(v["body"]), "R": np.array(v["R"]), "p": np.array(v["p"])}
                for k, v in d["frames"].items()
            },
            effort=np.array(d["effort"], dtype=np.float64),
            velocity=np.array(d["velocity"], dtype=np.float64),
            lower=np.array(d["lower"], dtype=np.float64),
            upper=np.array(d["upper"], dtype=np.float64),
            q_ref={k: np.array(v, dtype=np.float64) for k, v in d["q_ref"].items()},
        )

    # support (ancestor) mask of each body, used by tests and the host mirror
    def supports(self) -> np.ndarray:
        sup = np.zeros((self.nb, self.nb), dtype=bool)
        for b in range(self.nb):
            a = b
            while a >= 0:
                sup[b, a] = True
                a = self.parent[a]
        return sup


def compile_urdf(urdf_path: str, srdf_path: Optional[str] = None, name: Optional[str] = None) -> CompiledModel:
    """Restates pinocchio::urdf::buildModel(..., JointModelFreeFlyer()) + SRDF reference configs."""
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {}
    for j in root.findall("joint"):
        joints[j.get("name")] = j
    child_links = {j.find("child").get("link") for j in joints.values()}
    roots = [l for l in links if l not in child_links]
    assert len(roots) == 1, f"URDF must have exactly one root link, got {roots}"
    root_link = roots[0]

    # urdfdom: child_joints of a link in joint-name order (std::map iteration)
    children: Dict[str, List[str]] = {l: [] for l in links}
    for jn in sorted(joints):
        children[joints[jn].find("parent").get("link")].append(jn)

    parent: List[int] = [-1]
    jplace: List[SE3] = [SE3()]
    inert: List[Inertia] = [_link_inertia(links[root_link])]
    jnames: List[str] = []
    eff: List[float] = []
    vel: List[float] = []
    lo: List[float] = []
    up: List[float] = []
    frames: Dict[str, dict] = {}

    def visit(link: str, body: int, link_placement: SE3) -> None:
        # link_placement: pose of `link`'s frame in the frame of joint `body`
        for jn in children[link]:
            j = joints[jn]
            jt = j.get("type")
            child = j.find("child").get("link")
            M = link_placement * _origin(j)
            if jt == "fixed":
                inert[body] = inert[body] + _link_inertia(links[child]).se3_action(M)
                frames[jn] = {"body": body, "R": M.R.copy(), "p": M.p.copy()}
                visit(child, body, M)
            elif jt in ("revolute", "continuous"):
                axis = _floats(j.find("axis").get("xyz") if j.find("axis") is not None else "1 0 0", 3)
                if axis != [0.0, 0.0, 1.0]:
                    raise NotImplementedError(
                        f"joint {jn}: axis {axis}; the kernels implement JointModelRZ only "
                        "(every revolute joint of robot/v0 and robot/v1 has axis 0 0 1, SURVEY.md A1)"
                    )
                parent.append(body)
                jplace.append(M)
                inert.append(_link_inertia(links[child]))
                jnames.append(jn)
                lim = j.find("limit")
                eff.append(float(lim.get("effort", "0")) if lim is not None else 0.0)
                vel.append(float(lim.get("velocity", "0")) if lim is not None else 0.0)
                lo.append(float(lim.get("lower", "-inf")) if lim is not None and lim.get("lower") else -math.inf)
                up.append(float(lim.get("upper", "inf")) if lim is not None and lim.get("upper") else math.inf)
                visit(child, len(parent) - 1, SE3())
            else:
                raise NotImplementedError(f"joint {jn}: type {jt}")

    visit(root_link, 0, SE3())
    nb = len(parent)
    assert nb <= MAX_BODIES, nb

    q_ref: Dict[str, np.ndarray] = {}
    if srdf_path is not None and os.path.exists(srdf_path):
        sroot = ET.parse(srdf_path).getroot()
        for gs in sroot.findall("group_state"):
            q = np.zeros(nb - 1 + 7)
            q[6] = 1.0
            for sj in gs.findall("joint"):
                vals = [float(t) for t in sj.get("value").split()]
                jn = sj.get("name")
                if jn == "root_joint":
                    q[:7] = vals
                elif jn in jnames:
                    q[7 + jnames.index(jn)] = vals[0]
                # joints the URDF does not have are ignored, as Pinocchio does
            q_ref[gs.get("name")] = q

    return CompiledModel(
        name=name or os.path.basename(os.path.dirname(urdf_path)) or "robot",
        joint_names=jnames,
        parent=parent,
        jR=np.array([M.R for M in jplace]),
        jp=np.array([M.p for M in jplace]),
        mass=np.array([Y.m for Y in inert]),
        com=np.array([Y.c for Y in inert]),
        inertia=np.array([Y.I for Y in inert]),
        frames=frames,
        effort=np.array(eff),
        velocity=np.array(vel),
        lower=np.array(lo),
        upper=np.array(up),
        q_ref=q_ref,
    )


_MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")

# URDF path suffix (as conf.urdf names it) -> compiled table shipped with the package
_KNOWN = {
    "robot/v1/urdf/robot_mod.urdf": "robot_v1.json",
    "robot/v0/robot.urdf": "robot_v0.json",
    # legacy/op3_conf.py:59 still points at the pre-move location of the v0 model
    "robot/robot.urdf": "robot_v0.json",
}


def load_compiled(name: str) -> CompiledModel:
    with open(os.path.join(_MODELS_DIR, name)) as f:
        return CompiledModel.from_json(f.read())


def load_model(urdf: str, srdf: Optional[str] = None) -> CompiledModel:
    """Resolve ``conf.urdf``: compile the file when it exists, else fall back to the
    table precompiled from the reference's own model of that path."""
    if os.path.exists(urdf):
        return compile_urdf(urdf, srdf)
    norm = urdf.replace("\\", "/").lstrip("./")
    for suffix, blob in _KNOWN.items():
        if norm.endswith(suffix):
            return load_compiled(blob)
    raise FileNotFoundError(f"{urdf}: no such URDF and no precompiled table for it")


def main(argv: Sequence[str] = None) -> None:
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("urdf")
    ap.add_argument("--srdf")
    ap.add_argument("--name")
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args(argv)
    m = compile_urdf(a.urdf, a.srdf, a.name)
    with open(a.out, "w") as f:
        f.write(m.to_json())
    print(f"{a.out}: nb={m.nb} na={m.na} mass={m.total_mass:.6f} frames={sorted(m.frames)}")


if __name__ == "__main__":
    main()
