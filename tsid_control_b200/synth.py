"""Synthetic tick inputs — shared by bench.py and the parity tests.

Everything is fp64 numpy from ``numpy.random.Generator(PCG64(seed))``; seed 0 is the
performance seed, seeds >= 1 are parity seeds.  The states are perturbations of the
controller's standing configuration (SRDF "standing" with the z-shift of
ref:ctrl/WalkController.py:74); the references are the values at that configuration, so
PD errors are non-zero and asymmetric.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(x,y,z,w) quaternion product, batched on the leading axis."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        axis=-1,
    )


def quat_exp(w: np.ndarray) -> np.ndarray:
    t = np.linalg.norm(w, axis=-1, keepdims=True)
    half = 0.5 * t
    k = np.where(t > 1e-12, np.sin(half) / np.where(t > 1e-12, t, 1.0), 0.5)
    return np.concatenate([k * w, np.cos(half)], axis=-1)


def random_states(q_stand: np.ndarray, n: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """q [n,nq], v [n,nv]: base xyz +-0.02 m, base rotation exp(U(-0.1,0.1)^3), joints +-0.2 rad;
    v: base linear +-0.2 m/s, base angular +-0.5 rad/s, joints +-1 rad/s."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nq = q_stand.shape[0]
    na = nq - 7
    q = np.tile(q_stand, (n, 1))
    q[:, :3] += rng.uniform(-0.02, 0.02, (n, 3))
    dq = quat_exp(rng.uniform(-0.1, 0.1, (n, 3)))
    q[:, 3:7] = quat_mul(q[:, 3:7], dq)
    q[:, 7:] += rng.uniform(-0.2, 0.2, (n, na))
    v = np.concatenate(
        [rng.uniform(-0.2, 0.2, (n, 3)), rng.uniform(-0.5, 0.5, (n, 3)), rng.uniform(-1.0, 1.0, (n, na))], axis=1
    )
    return np.ascontiguousarray(q), np.ascontiguousarray(v)


def se3_vec12(p: np.ndarray, R: np.ndarray) -> np.ndarray:
    """tsid SE3ToVector: translation then rotation column-major."""
    return np.concatenate([np.asarray(p, dtype=np.float64), np.asarray(R, dtype=np.float64).T.ravel()])


def standing_refs(com9: np.ndarray, foot_lf12: np.ndarray, foot_rf12: np.ndarray, q_stand: np.ndarray) -> Dict[str, np.ndarray]:
    """References of a freshly constructed controller (ref:ctrl/WalkController.py:81,98,122,137,151-152,164-165):
    everything is the value at the standing configuration, zero velocity/acceleration."""
    com = np.zeros(9)
    com[:3] = com9[:3]
    z12 = np.zeros(12)
    return {
        "com": com,
        "foot_lf": np.concatenate([foot_lf12, z12]),
        "foot_rf": np.concatenate([foot_rf12, z12]),
        "contact_lf": foot_lf12.copy(),
        "contact_rf": foot_rf12.copy(),
        "posture": q_stand[7:].copy(),
    }


def walking_batch(refs0: Dict[str, np.ndarray], n: int, seed: int, step_length: float, step_width: float,
                  step_height: float, step_duration: float, com_height: float) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """Config-3 style walking inputs: per-env gait phase phi ~ U[0,1): 20 % double
    support, 40 % left-support (right foot swinging), 40 % right-support; the swing-foot reference
    follows the FootTrajectory semantics (x, y linear in t; z a parabola through (0,0), (T/2, h), (T,0),
    ref:ctrl/Foot_Trajectory.py:8-19 with rise_ratio 0.5) and the CoM reference one LIPM Euler step
    (ref:ctrl/LIPM.py:44-47) from a per-env velocity command vx ~ U(-0.3,0.3), vy ~ U(-0.1,0.1)."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    phi = rng.uniform(0.0, 1.0, n)
    mask = np.where(phi < 0.2, 3, np.where(phi < 0.6, 1, 2)).astype(np.uint8)  # bit0 LF, bit1 RF in contact
    vcmd = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], axis=1)
    s = np.where(phi < 0.2, 0.0, np.where(phi < 0.6, (phi - 0.2) / 0.4, (phi - 0.6) / 0.4))  # swing progress
    T = step_duration
    refs = {k: np.tile(v, (n, 1)) for k, v in refs0.items()}
    # swing foot: start at the standing placement - L/2, target + L/2, scaled by the command direction
    L = step_length * np.sign(vcmd[:, 0] + 1e-300)
    x = -0.5 * L + L * s
    xd = L / T
    z = 4.0 * step_height * s * (1.0 - s)
    zd = 4.0 * step_height * (1.0 - 2.0 * s) / T
    zdd = -8.0 * step_height / (T * T) * np.ones(n)
    for foot, key, bit in ((0, "foot_lf", 1), (1, "foot_rf", 2)):
        sw = (mask & bit) == 0
        refs[key][sw, 0] += x[sw]
        refs[key][sw, 2] += z[sw]
        refs[key][sw, 12] = xd[sw]
        refs[key][sw, 14] = zd[sw]
        refs[key][sw, 20] = zdd[sw]
    # CoM: one semi-implicit LIPM step from the standing CoM with the commanded velocity
    w2 = 9.80665 / com_height
    dt = 0.002
    zmp = refs0["com"][:2][None, :] + 0.0 * vcmd
    pos = np.tile(refs0["com"][:2], (n, 1))
    acc = (zmp - pos) * w2
    vel = vcmd + acc * dt
    pos = pos + vel * dt
    refs["com"][:, 0:2] = pos
    refs["com"][:, 3:5] = vel
    refs["com"][:, 6:8] = acc
    return mask, refs
