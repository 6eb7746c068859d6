"""Host-side reference generators against golden vectors produced by the REFERENCE'S OWN PYTHON
(tests/golden/make_planner_golden.py imports ref:ctrl/Footstep_Planner.py, Foot_Trajectory.py and LIPM.py)."""
import os

import numpy as np
import pytest

from tsid_control_b200.ctrl.Foot_Trajectory import FootTrajectory
from tsid_control_b200.ctrl.Footstep_Planner import Footstep, FootstepPlanner, Support
from tsid_control_b200.ctrl.LIPM import LIPM
from tsid_control_b200.ctrl.Walk_Planner import WalkPlanner

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "planners.npz"))


def _init():
    return [Footstep(np.array([0, 0.1]), np.array([0, 0, 0]), 0), Footstep(np.array([0, -0.1]), np.array([0, 0, 0]), 1)]


def test_footstep_planner_reference_demo_path():
    steps = FootstepPlanner(step_width=0.2, step_length=0.3).plan(list(G["fs_demo_path"]), _init())
    assert np.allclose([s.position for s in steps], G["fs_demo_pos"], rtol=0, atol=1e-14)
    assert np.allclose([s.orientation[2] for s in steps], G["fs_demo_yaw"], rtol=0, atol=1e-14)
    assert [int(s.side) for s in steps] == G["fs_demo_side"].tolist()


def test_footstep_planner_curved_path_and_support_polygon():
    steps = FootstepPlanner(step_width=0.2, step_length=0.3).plan(list(G["fs2_path"]), _init())
    assert len(steps) == len(G["fs2_pos"])
    assert np.allclose([s.position for s in steps], G["fs2_pos"], rtol=0, atol=1e-14)
    assert [int(s.side) for s in steps] == G["fs2_side"].tolist()
    poly = Support([steps[2], steps[3]], foot_width=0.1, foot_length=0.25).get_support_polygon()
    assert np.allclose(poly, G["support_poly"], rtol=0, atol=1e-14)
    assert np.allclose(Support([steps[3]], 0.1, 0.25).get_support_polygon(), G["support_single"], rtol=0, atol=1e-14)
    assert Support([steps[2], steps[3]], 0.1, 0.25).is_double_support


@pytest.mark.parametrize("tag,rr", [("r50", 0.5), ("r10", 0.1)])
def test_foot_trajectory_matches_scipy_splines_of_the_reference(tag, rr):
    tj = FootTrajectory([0.0, 0.5], np.array([0.1, 0.05, 0.0, 0.2]), np.array([0.4, 0.07, 0.02, -0.1]), 0.2, rr)
    ts = G["ft_ts"]
    assert np.allclose(tj.get_position(ts), G[f"ft_{tag}_pos"], rtol=0, atol=1e-13)
    assert np.allclose(tj.get_velocity(ts), G[f"ft_{tag}_vel"], rtol=0, atol=1e-10)  # reference: derivative order 2
    assert np.allclose(tj.get_acceleration(ts), G[f"ft_{tag}_acc"], rtol=0, atol=1e-8)  # reference: order 3
    assert np.allclose(tj.yaw(ts), G[f"ft_{tag}_yaw"], rtol=0, atol=1e-14)
    assert np.allclose(tj.velocity(ts), G[f"ft_{tag}_d1"], rtol=0, atol=1e-12)
    # scalar call form of the reference
    assert np.allclose(tj.get_position(0.25), G[f"ft_{tag}_pos"][5], rtol=0, atol=1e-13)


def test_foot_trajectory_degenerate_cases_of_survey_kat():
    # 3 knots -> one parabola, z(1/4) = 0.15 for h = 0.2 on [0, 1]
    tj = FootTrajectory([0, 1], np.array([0.0, 0.0, 0.0]), np.array([1.0, 1.0, 0.0]), 0.2)
    assert abs(float(tj.z(0.25)) - 0.15) < 1e-15 and abs(float(tj.z(0.5)) - 0.2) < 1e-15
    assert abs(float(tj.x(0.3)) - 0.3) < 1e-15


def test_lipm_matches_reference_euler_integration():
    pin = G["lipm_in"]
    lip = LIPM(0.2417, dt=0.002)
    pos0, vel0 = pin[0:2].copy(), pin[2:4].copy()
    lip.make_trajectory([0.0, 0.3], 0.002, pos0, vel0, pin[4:6].copy(), pin[6:8])
    assert abs(lip.w - float(G["lipm_w"])) == 0.0
    assert np.array_equal(np.array(lip.x.traj), G["lipm_x"]) and np.array_equal(np.array(lip.y.traj), G["lipm_y"])
    # in-place aliasing of the reference (ref:ctrl/LIPM.py:41-47): the caller's pos0/vel0 hold the final state
    assert pos0[0] == G["lipm_x"][-1, 0] and vel0[1] == G["lipm_y"][-1, 1]
    # accessors (the reference's Trajectory.get_frame cannot run; same intent)
    assert np.allclose(lip.pos(0.1), [G["lipm_x"][50, 0], G["lipm_y"][50, 0]])
    assert np.allclose(lip.dcm(0.1), lip.pos(0.1) + lip.vel(0.1) / lip.w)
    assert np.allclose(lip.zmp(0.1), lip.pos(0.1) - lip.acc(0.1) / lip.w**2)
    with pytest.raises(IndexError):
        lip.pos(10.0)
    # batched single step == first sample
    p, v, a = LIPM.step(pin[0:2], pin[2:4], pin[6:8], lip.w, 0.002)
    assert np.allclose([p[0], v[0], a[0]], G["lipm_x"][0], rtol=0, atol=1e-15)


def test_walk_planner_builds_swing_trajectories():
    steps = FootstepPlanner(step_width=0.2, step_length=0.3).plan(list(G["fs2_path"]), _init())
    plan = WalkPlanner().plan(steps)
    assert len(plan) == len(steps) - 2
    first = plan[0]
    assert first["side"] == 0 and first["t0"] == 0.0
    p0 = first["trajectory"].get_position(0.0)
    p1 = first["trajectory"].get_position(0.5)
    assert np.allclose(p0[:2], steps[0].position) and np.allclose(p1[:2], steps[2].position)
    assert abs(float(first["trajectory"].z(0.25)) - 0.2) < 1e-15  # conf.step_height at mid-swing
