"""The device gait phase machine (tsid_control_b200/csrc/tsidb_gait.cuh): its host build (tests/emu) against the numpy
restatement tests/gait_ref.py, and the swing-foot shape against this repo's FootTrajectory, which is pinned to the
reference's own scipy splines by tests/golden/planners.npz."""
import numpy as np

from common import setup
from emu_py import EmuGait
from gait_ref import GaitRef
from tsid_control_b200.ctrl.Foot_Trajectory import FootTrajectory

GAIT = dict(dt=0.002, step_duration=0.5, step_length=0.3, step_height=0.2)


def _fake_feet(rng, n, k):
    """placements that move from step to step (so that lift-off / touch-down latch different values)"""
    eye = np.eye(3).T.ravel()
    out = []
    for f in range(2):
        p = np.c_[0.01 * k + rng.uniform(-0.01, 0.01, n), (0.1 if f == 0 else -0.1) + rng.uniform(-0.01, 0.01, n), rng.uniform(0, 0.02, n)]
        out.append(np.c_[p, np.tile(eye, (n, 1))])
    return out


def test_emulated_gait_matches_numpy_restatement_over_two_cycles():
    s = setup("v1")
    n = 40
    rng = np.random.Generator(np.random.PCG64(5))
    phase0 = rng.uniform(0, 1, n)
    vcmd = np.c_[rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)]
    h0 = float(s["refs"]["com"][2])
    a = EmuGait(n, com_height=h0, defaults=s["refs"], phase0=phase0, vcmd=vcmd, **GAIT)
    b = GaitRef(n, com_height=h0, defaults=s["refs"], phase0=phase0, vcmd=vcmd, **GAIT)
    ticks = int(2 * 0.5 / 0.4 / 0.002) + 7  # two gait cycles
    seen = set()
    for k in range(ticks):
        fl, fr = _fake_feet(rng, n, k)
        st = (rng.uniform(0, 1, n) < 0.01).astype(np.int32)
        a.step(fl, fr, st)
        b.step(fl, fr, st)
        assert np.array_equal(a.mask, b.mask)
        seen.update(int(m) for m in a.mask)
        for x, y in ((a.phi, b.phi), (a.com, b.com), (a.foot[0], b.foot[0]), (a.foot[1], b.foot[1]),
                     (a.contact[0], b.contact[0]), (a.contact[1], b.contact[1])):
            assert np.abs(x - y).max() < 1e-12
    assert seen == {1, 2, 3} and np.array_equal(a.fails, b.fails) and a.fails.sum() > 0


def test_swing_reference_is_the_reference_foot_trajectory():
    """one env, one swing of the left foot: positions, first and second derivatives of the device rule equal
    FootTrajectory's polynomials (x, y linear; z the 3-knot parabola, ref:ctrl/Foot_Trajectory.py:13-19)"""
    s = setup("v1")
    h0 = float(s["refs"]["com"][2])
    g = GaitRef(1, com_height=h0, defaults=s["refs"], phase0=[0.6 - 1e-9], vcmd=[[0.2, 0.0]], **GAIT)
    now = [s["refs"]["foot_lf"][None, :12].copy(), s["refs"]["foot_rf"][None, :12].copy()]
    start = now[0][0, :3].copy()
    target = start + np.array([GAIT["step_length"], 0.0, 0.0])
    ft = FootTrajectory([0.0, GAIT["step_duration"]], np.r_[start, 0.0], np.r_[target, 0.0], GAIT["step_height"], 0.5)
    for k in range(1, 240):
        g.step(now[0], now[1])
        if g.mask[0] != 2:
            break
        t = (g.phi[0] - 0.6) / 0.4 * GAIT["step_duration"]
        ref = g.foot[0][0]
        assert np.abs(ref[:3] - ft.get_position(t)[:3]).max() < 1e-12
        assert np.abs(ref[12:15] - ft.velocity(t)[:3]).max() < 1e-10
        assert np.abs(ref[18:21] - ft.acceleration(t)[:3]).max() < 1e-9
    assert k > 200


def _plans(n, rng):
    """Per-env footstep plans through the HOST planner (pinned to the reference by tests/golden/planners.npz): a curved
    path per env, the two initial supports at the standing soles."""
    from tsid_control_b200.ctrl.Footstep_Planner import Footstep, FootstepPlanner

    S = 40
    steps, ns = np.zeros((n, S, 4)), np.zeros(n, np.int32)
    for e in range(n):
        w = rng.uniform(-0.4, 0.4)
        x = y = th = 0.0
        path = []
        for _ in range(50):
            x += 0.05 * np.cos(th); y += 0.05 * np.sin(th); th += w * 0.1
            path.append(np.array([x, y]))
        init = [Footstep(np.array([0, 0.1]), np.array([0, 0, 0]), 0), Footstep(np.array([0, -0.1]), np.array([0, 0, 0]), 1)]
        fs = FootstepPlanner(step_width=0.2, step_length=0.3).plan(path, init)
        ns[e] = len(fs)
        for k, s_ in enumerate(fs):
            steps[e, k] = [s_.position[0], s_.position[1], s_.orientation[2], int(s_.side)]
    return steps, ns


def test_emulated_planned_gait_follows_the_footstep_plan_with_yaw_and_four_knot_swing():
    """Planned mode of the device gait (tsidb_gait_set_plan): every swing goes from the lift-off placement to the env's
    next planned footstep of that side along FootTrajectory with rise_ratio 0.3 (4-knot z spline), x, y and YAW linear,
    first/second derivatives as velocity/acceleration references — device code (host build) against the numpy
    restatement built on this repo's FootTrajectory / FootstepPlanner classes (both pinned to the reference's Python)."""
    s = setup("v1")
    n = 24
    rng = np.random.Generator(np.random.PCG64(11))
    phase0 = rng.uniform(0, 1, n)
    h0 = float(s["refs"]["com"][2])
    steps, ns = _plans(n, rng)
    a = EmuGait(n, com_height=h0, defaults=s["refs"], phase0=phase0, steps=steps, n_steps=ns, rise_ratio=0.3, **GAIT)
    b = GaitRef(n, com_height=h0, defaults=s["refs"], phase0=phase0, steps=steps, n_steps=ns, rise_ratio=0.3, **GAIT)
    ticks = int(3 * 0.5 / 0.4 / 0.002)  # three gait cycles: six swings per env
    yawed = 0.0
    for k in range(ticks):
        # the feet "track" their references of the previous tick (what a converged tick would measure)
        fl, fr = b.foot[0][:, :12].copy(), b.foot[1][:, :12].copy()
        a.step(fl, fr)
        b.step(fl, fr)
        assert np.array_equal(a.mask, b.mask) and np.array_equal(a.step_idx, b.step_idx)
        for x, y in ((a.foot[0], b.foot[0]), (a.foot[1], b.foot[1]), (a.contact[0], b.contact[0]), (a.contact[1], b.contact[1]), (a.com, b.com)):
            assert np.abs(x - y).max() < 1e-11
        yawed = max(yawed, float(np.abs(a.foot[0][:, 17]).max()))
    assert (a.step_idx > 4).all() and yawed > 0.05  # plans were consumed and the feet turned (yaw-rate references)
    # at the end of a swing the foot reference sits on the planned footstep
    e = 0
    done = steps[e, :a.step_idx[e]]
    last = {int(sd): done[done[:, 3] == sd][-1] for sd in (0, 1) if (done[2:, 3] == sd).any()}
    for f, st in last.items():
        if a.mask[e] & (1 << f):  # foot f is down: its contact reference is where the swing ended
            assert np.abs(a.contact[f][e, :2] - st[:2]).max() < 0.02
