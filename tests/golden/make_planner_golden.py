#!/usr/bin/env python
"""Golden vectors for the host-side reference generators, produced by IMPORTING THE REFERENCE'S OWN PYTHON
(ref:ctrl/Footstep_Planner.py, ref:ctrl/Foot_Trajectory.py, ref:ctrl/LIPM.py) from /root/reference in the build
container.  matplotlib (absent here) is stubbed so that the import-time demo plots are no-ops; LIPM's broken
`from Trajectory import Trajectory` (ref:ctrl/LIPM.py:3, ref:ctrl/Trajectory.py:1) is satisfied with a stand-in
list container — the arithmetic of make_trajectory is the reference's.

    python tests/golden/make_planner_golden.py        # writes tests/golden/planners.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class _Anything:
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda k: _Anything()
    sys.modules[name] = m
    return m


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    tr = types.ModuleType("Trajectory")

    class Trajectory:
        def __init__(self):
            self.traj = []

    tr.Trajectory = Trajectory
    sys.modules["Trajectory"] = tr

    fp = _load(os.path.join(REF, "ctrl", "Footstep_Planner.py"), "ref_footstep_planner")
    ft = _load(os.path.join(REF, "ctrl", "Foot_Trajectory.py"), "ref_foot_trajectory")
    lp = _load(os.path.join(REF, "ctrl", "LIPM.py"), "ref_lipm")
    out = {}

    # --- FootstepPlanner: the reference's own demo path (module level) and a second, straight path
    steps = fp.footsteps
    out["fs_demo_pos"] = np.array([s.position for s in steps], dtype=np.float64)
    out["fs_demo_yaw"] = np.array([s.orientation[2] for s in steps], dtype=np.float64)
    out["fs_demo_side"] = np.array([int(s.side) for s in steps], dtype=np.int64)
    out["fs_demo_path"] = np.array(fp.path, dtype=np.float64)
    planner = fp.FootstepPlanner(step_width=0.2, step_length=0.3)
    path = [np.array([0.05 * i, 0.01 * i * i]) for i in range(40)]
    init = [fp.Footstep(np.array([0, 0.1]), np.array([0, 0, 0]), 0), fp.Footstep(np.array([0, -0.1]), np.array([0, 0, 0]), 1)]
    steps2 = planner.plan(path, init)
    out["fs2_path"] = np.array(path)
    out["fs2_pos"] = np.array([s.position for s in steps2], dtype=np.float64)
    out["fs2_yaw"] = np.array([s.orientation[2] for s in steps2], dtype=np.float64)
    out["fs2_side"] = np.array([int(s.side) for s in steps2], dtype=np.int64)
    sup = fp.Support([steps2[2], steps2[3]], foot_width=0.1, foot_length=0.25)
    out["support_poly"] = np.array(sup.get_support_polygon(), dtype=np.float64)
    out["support_single"] = np.array(fp.Support([steps2[3]], 0.1, 0.25).get_support_polygon(), dtype=np.float64)

    # --- FootTrajectory: rise_ratio 0.5 (3 knots) and 0.1 (4 knots); reference accessor semantics
    ts = np.linspace(0.0, 0.5, 11)
    for tag, rr in (("r50", 0.5), ("r10", 0.1)):
        tj = ft.FootTrajectory([0.0, 0.5], np.array([0.1, 0.05, 0.0, 0.2]), np.array([0.4, 0.07, 0.02, -0.1]), 0.2, rr)
        out[f"ft_{tag}_pos"] = np.array([tj.get_position(t) for t in ts], dtype=np.float64)
        out[f"ft_{tag}_vel"] = np.array([tj.get_velocity(t) for t in ts], dtype=np.float64)  # order 2 in the reference
        out[f"ft_{tag}_acc"] = np.array([tj.get_acceleration(t) for t in ts], dtype=np.float64)  # order 3
        out[f"ft_{tag}_yaw"] = np.array([tj.yaw(t) for t in ts], dtype=np.float64)
        out[f"ft_{tag}_d1"] = np.array([[tj.x(t, 1), tj.y(t, 1), tj.z(t, 1)] for t in ts], dtype=np.float64)
    out["ft_ts"] = ts

    # --- LIPM: semi-implicit Euler (ref:ctrl/LIPM.py:34-49)
    lip = lp.LIPM(0.2417)
    pos0, vel0, acc0 = np.array([0.0, 0.05]), np.array([0.1, 0.0]), np.array([0.0, 0.0])
    zmp = np.array([0.02, 0.0])
    lip.make_trajectory([0.0, 0.3], 0.002, pos0.copy(), vel0.copy(), acc0.copy(), zmp)
    out["lipm_w"] = np.array(lip.w)
    out["lipm_x"] = np.array(lip.x.traj, dtype=np.float64)
    out["lipm_y"] = np.array(lip.y.traj, dtype=np.float64)
    out["lipm_in"] = np.concatenate([pos0, vel0, acc0, zmp])
    np.savez_compressed(os.path.join(HERE, "planners.npz"), **out)
    print("wrote planners.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
