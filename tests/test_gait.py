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
