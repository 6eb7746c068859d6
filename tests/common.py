"""Shared fixtures: compiled models, C structs, oracle instances, standing references."""
from __future__ import annotations

import functools
import importlib

import numpy as np

from tsid_control_b200 import synth
from tsid_control_b200._capi import conf_to_c, model_to_c
from tsid_control_b200.model_compiler import load_compiled

from oracle_py import Oracle


@functools.lru_cache(maxsize=None)
def setup(kind: str, variant: str = "liboracle.so"):
    """kind: 'v1' (WalkController + ctrl/conf.py) or 'v0' (legacy Biped + op3_conf)."""
    if kind == "v1":
        from tsid_control_b200.ctrl.conf import RobotConfig

        conf = RobotConfig()
        m = load_compiled("robot_v1.json")
        cm = model_to_c(m, conf.lf_fixed_joint, conf.rf_fixed_joint)
        cc = conf_to_c(conf, m, legacy=False)
    else:
        conf = importlib.import_module("tsid_control_b200.legacy.op3_conf")
        m = load_compiled("robot_v0.json")
        cm = model_to_c(m, conf.lf_frame_name, conf.rf_frame_name)
        cc = conf_to_c(conf, m, legacy=True)
    orc = Oracle(cm, cc, variant)
    q0 = m.q_ref["standing"].copy()
    v0 = np.zeros(m.nv)
    ident = np.r_[np.zeros(3), np.eye(3).ravel()]
    dummy = {"com": np.zeros(9), "foot_lf": np.r_[ident, np.zeros(12)], "foot_rf": np.r_[ident, np.zeros(12)],
             "contact_lf": ident, "contact_rf": ident, "posture": np.zeros(m.na)}
    if kind == "v1":
        # ref:ctrl/WalkController.py:72-76: put the left sole on z = 0
        r = orc.tick(q0, v0, 3, dummy)
        q0[2] -= r["foot"][0][2]
    r = orc.tick(q0, v0, 3, dummy)
    refs = synth.standing_refs(r["com"], r["foot"][0], r["foot"][1], q0)
    return {"conf": conf, "model": m, "cm": cm, "cc": cc, "oracle": orc, "q0": q0, "refs": refs}


def canonical_active(rows):
    """Active set as a set of (block, side, i) with degenerate contact corners canonicalised.

    When a corner's normal force is zero all four pyramid rows of that corner are tight but only
    three are linearly independent; which three end in the working set is decided by last-bit
    rounding (the fp64 and 80-bit builds of the SAME oracle already differ there, see
    test_fp64_against_long_double_truth).  Such a corner is reported as all four rows."""
    out = set(rows)
    for blk in (0, 1):
        for c in range(4):
            grp = [(blk, 1, 4 * c + k) for k in range(4)]
            if sum(g in out for g in grp) >= 3:
                out.update(grp)
    return out


def ci_bit(na: int, nv: int, block: int, side: int, i: int) -> int:
    """Row numbering of tsidb_ci_row(): blocks LF(17), RF(17), actuation(na), joint bounds(nv), each
    stacked [lower rows; upper rows]."""
    rows = (17, 17, na, nv)[block]
    off = (0, 17, 34, 34 + na)[block]
    return 2 * off + side * rows + i


def bits_to_rows(na: int, nv: int, bits):
    table = {}
    for blk, rows in enumerate((17, 17, na, nv)):
        for side in (0, 1):
            for i in range(rows):
                table[ci_bit(na, nv, blk, side, i)] = (blk, side, i)
    return [table[b] for b in bits]
