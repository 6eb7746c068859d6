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


# ---- the same planners as DEVICE code (tsidb_gait.cuh), here through the host build of that header (tests/emu);
# ---- tests/test_gpu_parity.py runs the CUDA kernels against the same golden vectors
def _emu_lib():
    import ctypes as C

    from emu_py import build_emu

    return C.CDLL(build_emu()), C


def _device_foot_trajectory(run, tag, rr):
    ts = G["ft_ts"]
    n = len(ts)
    start = np.tile(np.array([0.1, 0.05, 0.0, 0.2]), (n, 1))
    target = np.tile(np.array([0.4, 0.07, 0.02, -0.1]), (n, 1))
    out = run(0.0, 0.5, start, target, 0.2, rr, np.ascontiguousarray(ts)).reshape(n, 4, 4)
    assert np.abs(out[:, 0, :3] - G[f"ft_{tag}_pos"]).max() < 1e-13
    assert np.abs(out[:, 0, 3] - G[f"ft_{tag}_yaw"]).max() < 1e-14
    assert np.abs(out[:, 1, :3] - G[f"ft_{tag}_d1"]).max() < 1e-12
    assert np.abs(out[:, 2, :3] - G[f"ft_{tag}_vel"]).max() < 1e-10   # the reference's "velocity" is the 2nd derivative (:35)
    assert np.abs(out[:, 3, :3] - G[f"ft_{tag}_acc"]).max() < 1e-8    # and its "acceleration" the 3rd (:43)


def _device_footstep_plan(run):
    paths = [G["fs_demo_path"], G["fs2_path"]]
    P = max(len(p) for p in paths)
    path = np.zeros((2, P, 2))
    for e, p in enumerate(paths):
        path[e, :len(p)] = p
    n_pts = np.array([len(p) for p in paths], np.int32)
    init = np.tile(np.array([[0, 0.1, 0, 0], [0, -0.1, 0, 1]], np.float64), (2, 1, 1))
    steps, ns = run(path, n_pts, init, 0.3, 0.2, 64)
    for e, tag in enumerate(("fs_demo", "fs2")):
        k = int(ns[e])
        assert k == len(G[f"{tag}_pos"])
        assert np.abs(steps[e, :k, :2] - G[f"{tag}_pos"]).max() < 1e-14
        assert np.abs(steps[e, :k, 2] - G[f"{tag}_yaw"]).max() < 1e-14
        assert steps[e, :k, 3].astype(int).tolist() == G[f"{tag}_side"].tolist()


@pytest.mark.parametrize("tag,rr", [("r50", 0.5), ("r10", 0.1)])
def test_device_foot_trajectory_code_matches_the_reference_splines(tag, rr):
    lib, C = _emu_lib()

    def run(t0, t1, start, target, h, r, t):
        out = np.zeros((len(t), 16))
        lib.emu_foot_trajectory.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        assert lib.emu_foot_trajectory(len(t), t0, t1, start.ctypes.data, target.ctypes.data, h, r, t.ctypes.data, out.ctypes.data) == 0
        return out

    _device_foot_trajectory(run, tag, rr)


def test_device_footstep_planner_code_matches_the_reference_plan():
    lib, C = _emu_lib()

    def run(path, n_pts, init, L, W, max_steps):
        n = path.shape[0]
        steps, ns = np.zeros((n, max_steps, 4)), np.zeros(n, np.int32)
        lib.emu_footstep_plan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        assert lib.emu_footstep_plan(n, path.ctypes.data, n_pts.ctypes.data, path.shape[1], init.ctypes.data, L, W, steps.ctypes.data,
                                     ns.ctypes.data, max_steps) == 0
        return steps, ns

    _device_footstep_plan(run)
