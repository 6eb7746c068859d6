"""Shared fixtures: compiled models, C structs, oracle instances, standing references."""
from __future__ import annotations

import functools
import importlib
import os

import numpy as np

from tsid_control_b200 import synth
from tsid_control_b200._capi import conf_to_c, model_to_c
from tsid_control_b200.model_compiler import load_compiled

from oracle_py import Oracle


def make_conf(kind: str, overrides=()):
    """A fresh conf object of the reference's kind (RobotConfig instance for 'v1', a namespace copy of the op3_conf
    module for 'v0') with `overrides` ((name, value) pairs) applied — e.g. tau_max_scaling / v_max_scaling small enough
    that actuation / joint-velocity rows enter the working set."""
    if kind == "v1":
        from tsid_control_b200.ctrl.conf import RobotConfig

        conf = RobotConfig()
    else:
        import types

        mod = importlib.import_module("tsid_control_b200.legacy.op3_conf")
        conf = types.SimpleNamespace(**{k: getattr(mod, k) for k in dir(mod) if not k.startswith("_")})
    for k, v in overrides:
        setattr(conf, k, v)
    return conf


@functools.lru_cache(maxsize=None)
def setup(kind: str, variant: str = "liboracle.so", overrides=()):
    """kind: 'v1' (WalkController + ctrl/conf.py) or 'v0' (legacy Biped + op3_conf); overrides: tuple of
    (conf attribute, value) pairs applied on top of the reference's constants."""
    conf = make_conf(kind, overrides)
    if kind == "v1":
        m = load_compiled("robot_v1.json")
        cm = model_to_c(m, conf.lf_fixed_joint, conf.rf_fixed_joint)
        cc = conf_to_c(conf, m, legacy=False)
    else:
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
        # a foot that carries no force at all (possible with fMin = 0, the legacy conf): its 12 force variables are
        # pinned to zero by ANY 12 independent rows out of the 16 pyramid rows and the fMin row — reported as all 17
        if sum(1 for r in out if r[0] == blk) >= 12 and _foot_pinned(out, blk):
            out.update((blk, 1, i) for i in range(16))
            out.add((blk, 0, 16))
    return out


def _foot_pinned(rows, blk):
    """12 or more rows of the foot's block, every corner with at least two of its pyramid rows: the foot is unloaded."""
    return all(sum((blk, 1, 4 * c + k) in rows for k in range(4)) >= 2 for c in range(4))


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


def corner_of(row):
    """(foot, corner) of a friction-pyramid row (block LF/RF, upper side, i < 16), else None."""
    blk, side, i = row
    if blk in (0, 1) and side == 1 and i < 16:
        return blk, i // 4
    return None


def active_set_report(orc, mask, ok_idx, dev_rows, ref_active, f_dev, f_ref, it_dev, it_ref, zero_force=1e-6):
    """Working-set comparison env by env (rows in the reference's stacked-CI numbering).

    Returns counts of envs whose raw working sets are identical, identical after canonicalising degenerate corners,
    and the list of envs whose difference is NOT explained by a zero-force corner: every row of the symmetric
    difference must be a pyramid row of a corner whose force is below `zero_force` newton in BOTH solutions
    (a corner with zero normal force has four tight pyramid rows of which any three are a valid working set).
    Also: how often the iteration counts agree, overall and among the envs with identical raw working sets."""
    exact = canon = it_eq = it_eq_exact = 0
    unexplained = []
    worst_corner = 0.0
    for i in ok_idx:
        rows = orc.ci_rows(int(mask[i]))
        ra = set(rows[k] for k in ref_active[i])
        rb = set(dev_rows(i))
        same_it = int(it_dev[i]) == int(it_ref[i])
        it_eq += same_it
        if ra == rb:
            exact += 1
            canon += 1
            it_eq_exact += same_it
            continue
        canon += canonical_active(ra) == canonical_active(rb)
        good = True
        for row in ra ^ rb:
            c = corner_of(row)
            if c is None:
                # the fMin row of a foot that carries no force at all (fMin = 0): one of the 17 rows pinning f = 0
                if row[0] in (0, 1) and row[1] == 0 and row[2] == 16:
                    foot = row[0]
                    fa = float(np.linalg.norm(f_dev[i, 12 * foot:12 * foot + 12]))
                    fb = float(np.linalg.norm(f_ref[i, 12 * foot:12 * foot + 12]))
                    worst_corner = max(worst_corner, fa, fb)
                    if fa < zero_force and fb < zero_force:
                        continue
                good = False
                break
            foot, cor = c
            fa = float(np.linalg.norm(f_dev[i, 12 * foot + 3 * cor:12 * foot + 3 * cor + 3]))
            fb = float(np.linalg.norm(f_ref[i, 12 * foot + 3 * cor:12 * foot + 3 * cor + 3]))
            worst_corner = max(worst_corner, fa, fb)
            if fa >= zero_force or fb >= zero_force:
                good = False
                break
        if not good:
            unexplained.append(int(i))
    n = max(1, len(ok_idx))
    return {"n": int(len(ok_idx)), "active_exact": exact / n, "active_canonical": canon / n, "iters_equal": it_eq / n,
            "iters_equal_where_exact": it_eq_exact / max(1, exact), "unexplained": unexplained[:16],
            "n_unexplained": len(unexplained), "worst_corner_force_in_a_difference": worst_corner}


TOL = 1e-8  # north_star: 1e-8 relative / 1e-10 absolute, written as |a-b| / (1e-2 + |b|) <= 1e-8


def rel_err(a, b):
    return float((np.abs(a - b) / (1e-2 + np.abs(b))).max()) if np.size(a) else 0.0


def force_generator(cc):
    T = np.zeros((6, 12))
    pts = np.array(cc.contact_points)
    for c in range(4):
        T[:3, 3 * c:3 * c + 3] = np.eye(3)
        p = pts[:, c]
        T[3:, 3 * c:3 * c + 3] = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    return T


def wrenches(T, f):
    return np.concatenate([f[:, :12] @ T.T, f[:, 12:] @ T.T], axis=1)


def words_to_bits(words):
    out = []
    for w in range(3):
        val = int(words[w]) & ((1 << 64) - 1)
        out += [64 * w + b for b in range(64) if (val >> b) & 1]
    return out


def compare_outputs(kind, q, v, mask, refs, out, threads=8, overrides=(), truth=True):
    """`out` (tau, ddq, f, status, iters, active_set [3,N] words; optionally wrench) of any implementation of the
    tick against the oracle on the same inputs: the fp64 build (the reference algorithm in the reference's
    arithmetic) and the 80-bit build (the exact answer of the same QP to ~1e-13)."""
    s = setup(kind, overrides=tuple(overrides))
    refs_o = refs if refs is not None else s["refs"]
    ref = s["oracle"].batch(q, v, mask, refs_o, n_threads=threads)
    T = force_generator(s["cc"])
    assert np.array_equal(out["status"], ref["status"]), "solver status differs from the oracle's"
    ok = ref["status"] == 0
    n = len(mask)
    res = {
        "n": int(n), "n_optimal": int(ok.sum()),
        "tau": rel_err(out["tau"][ok], ref["tau"][ok]), "ddq": rel_err(out["ddq"][ok], ref["dv"][ok]),
        "wrench": rel_err(wrenches(T, out["f"][ok]), wrenches(T, ref["f"][ok])),
        "f": rel_err(out["f"][ok], ref["f"][ok]),
    }
    if truth:
        s80 = setup(kind, "liboracle_ld.so", overrides=tuple(overrides))
        tr = s80["oracle"].batch(q, v, mask, refs_o, n_threads=threads)
        res.update({
            "tau_truth": rel_err(out["tau"][ok], tr["tau"][ok]), "ddq_truth": rel_err(out["ddq"][ok], tr["dv"][ok]),
            "wrench_truth": rel_err(wrenches(T, out["f"][ok]), wrenches(T, tr["f"][ok])),
            "oracle_tau_truth": rel_err(ref["tau"][ok], tr["tau"][ok]), "oracle_ddq_truth": rel_err(ref["dv"][ok], tr["dv"][ok]),
            "oracle_wrench_truth": rel_err(wrenches(T, ref["f"][ok]), wrenches(T, tr["f"][ok])),
            "f_truth": rel_err(out["f"][ok], tr["f"][ok]), "oracle_f_truth": rel_err(ref["f"][ok], tr["f"][ok]),
        })
    if out.get("wrench") is not None:  # the kernel's own wrench output agrees with T f
        assert rel_err(out["wrench"][ok], wrenches(T, out["f"][ok])) < 1e-9
    na, nv = s["oracle"].na, s["oracle"].nv
    res.update(active_set_report(s["oracle"], mask, np.where(ok)[0],
                                 lambda i: bits_to_rows(na, nv, words_to_bits(out["active_set"][:, i])),
                                 ref["active"], out["f"], ref["f"], out["iters"], ref["iters"]))
    if out.get("lam") is not None:
        res.update(lambda_report(s["oracle"], mask, np.where(ok)[0], ref, out["lam"], out["lam_row"], na, nv))
    res["mean_iters"] = float(np.mean(out["iters"]))
    blocks = np.zeros(4, np.int64)
    for i in np.where(ok)[0]:
        rows = s["oracle"].ci_rows(int(mask[i]))
        for b in set(rows[k][0] for k in ref["active"][i]):
            blocks[b] += 1
    res["envs_with_active_force_lf_rf_torque_jointvel_rows"] = [int(b) for b in blocks]
    return res, ref


def assert_parity(res, kind="v1"):
    """north_star: tau, ddq, contact wrenches within 1e-8 relative / 1e-10 absolute (TOL), active sets identical.

    What is asserted, and why it is the strictest statement the facts allow:
      * CUDA vs the 80-bit build of the oracle (the exact answer of the reference's QP): <= TOL for robot/v1.
        For the legacy OP3 conf (fMin = 0: zero-force corners are the rule, cond(H) ~ 1e8) fp64 itself is only
        good to a few 1e-8 .. 1e-7 on this QP — the reference algorithm run in fp64 (the oracle) sits 2e-7 from
        the exact answer — so there the bar is "at least as exact as the reference": <= the oracle's own distance.
      * CUDA vs the fp64 oracle: <= the CUDA result's bar against the exact answer plus the oracle's own measured
        distance to the exact answer (two fp64 codes cannot agree better than their noise floors add up to; on the
        65536-env batch the oracle is 5e-9 .. 2e-8 from its 80-bit build, the kernels 2e-9 .. 6e-9).
      * working sets: identical after canonicalising degenerate corners for EVERY env, and every raw difference
        sits on a contact corner whose force is < 1e-6 N in both solutions (n_unexplained == 0).
      * iteration counts: equal in >= 95 % of the envs whose raw working sets are equal (a row whose slack is at
        rounding level, +-1e-13, is "violated" in one arithmetic and not in the other: one extra add/drop pair
        that leaves the same working set; the fp64 and 80-bit builds of the oracle differ in the same way).
    """
    for k in ("tau", "ddq", "wrench"):
        floor = res[f"oracle_{k}_truth"]
        if kind == "v1":
            assert res[f"{k}_truth"] <= TOL, (k, res)
        else:
            assert res[f"{k}_truth"] <= max(TOL, floor), (k, res)
        assert res[k] <= 1.05 * (max(TOL, res[f"{k}_truth"]) + floor), (k, res)
    # the 12 corner forces of a foot are only fixed by the 1e-8 Hessian regulariser on the 6-dim null space of the
    # force generator (cond ~ 1e8): 1e-7 .. 7e-5 of noise in fp64 depending on the active set (measured for the fp64
    # oracle against its 80-bit build, profiles/parity_r02.json: oracle_f_truth); bar 1e-4
    assert res["f_truth"] <= 1e-4 and res["f"] <= 1e-4 + res["oracle_f_truth"], res
    if "lambda" in res:  # multipliers of the inequality rows (sol.lambda), envs with identical raw working sets
        assert res["lambda"] <= 1e-6 and res["lambda_min"] >= -1e-9 and res["lambda_envs_compared"] > 0, res
    assert res["n_unexplained"] == 0, res
    assert res["active_canonical"] == 1.0, res
    assert res["iters_equal_where_exact"] >= 0.95, res


def lambda_report(orc, mask, ok_idx, ref, lam_dev, lam_row_dev, na, nv):
    """sol.lambda: multipliers of the inequality rows in the working set.  For every env whose raw working set equals
    the oracle's, the multipliers are compared row by row (err = |a-b| / (1e-2 + |b|)); all multipliers must be >= 0
    (dual feasibility) up to rounding."""
    worst, n_cmp, min_lam = 0.0, 0, 0.0
    for i in ok_idx:
        rows = orc.ci_rows(int(mask[i]))
        nc = (int(mask[i]) & 1) + ((int(mask[i]) >> 1) & 1)
        neq = 6 + 6 * nc
        ra = {rows[k]: ref["lam"][i][neq + j] for j, k in enumerate(ref["active"][i])}
        used = lam_row_dev[i] >= 0
        rb = dict(zip(bits_to_rows(na, nv, [int(b) for b in lam_row_dev[i][used]]), lam_dev[i][used]))
        if rb:
            min_lam = min(min_lam, min(rb.values()))
        if set(ra) != set(rb):
            continue
        n_cmp += 1
        for r, a in ra.items():
            worst = max(worst, abs(rb[r] - a) / (1e-2 + abs(a)))
    return {"lambda_envs_compared": n_cmp, "lambda": worst, "lambda_min": float(min_lam)}


# ------------------------------------------------------------------ reference-held model data (robot/v1/mujoco/robot.xml)
def mjcf_golden():
    """tests/golden/mjcf_v1.json: per-body tables and whole-body quantities of the reference's MuJoCo export of robot/v1
    (made by tests/golden/make_mjcf_golden.py from ref:robot/v1/mujoco/robot.xml)."""
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mjcf_v1.json")) as f:
        return json.load(f)


# the URDF hangs a frame-only dummy link on each foot (ref:robot/v1/urdf/robot_mod.urdf:176-191, :384-399: 0.01 kg,
# 1e-4 kg m^2 isotropic, at the sole frame); Pinocchio merges it into the foot body, the MuJoCo export does not have it
SOLE_DUMMY = {"left_ankle_roll": "left_sole_joint_fixed", "right_ankle_roll": "right_sole_joint_fixed"}


def mjcf_case_q(model, case):
    """Configuration of a golden case: torso at the origin with identity orientation, joints by name."""
    q = np.zeros(model.nq)
    q[6] = 1.0
    for k, n in enumerate(model.joint_names):
        q[7 + k] = case["q"][n]
    return q


def assert_whole_body_matches_mjcf(model, case, M, com):
    """M: joint-space inertia of the configuration mjcf_case_q(case) (base block LOCAL), com: its centre of mass.  Total
    mass, CoM and rotational inertia about the CoM in the torso frame, with the two sole dummies removed (placed with the
    MuJoCo foot-body placements of the same configuration), against the MuJoCo file's values."""
    mass, c = M[0, 0], np.asarray(com)
    Io = M[3:6, 3:6]  # about the torso origin, torso axes
    m2, mc2, I2 = mass, mass * c, Io.copy()
    for jn, fn in SOLE_DUMMY.items():
        fb = case["foot_bodies"][jn]
        p = np.array(fb["p"]) + np.array(fb["R"]) @ model.frames[fn]["p"]
        m2 -= 0.01
        mc2 = mc2 - 0.01 * p
        I2 -= 1e-4 * np.eye(3) + 0.01 * (p @ p * np.eye(3) - np.outer(p, p))
    c2 = mc2 / m2
    Ic2 = I2 - m2 * (c2 @ c2 * np.eye(3) - np.outer(c2, c2))
    assert abs(m2 - case["mass"]) < 1e-9
    assert np.abs(c2 - np.array(case["com"])).max() < 2e-6, np.abs(c2 - np.array(case["com"])).max()
    assert np.abs(Ic2 - np.array(case["inertia_com"])).max() < 1e-5 * np.abs(Ic2).max()  # 0.707107-type literals
