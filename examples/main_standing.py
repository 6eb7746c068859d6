#!/usr/bin/env python
"""The reference's main loop (ref:main.py:46-47,110-144) against this repository's drop-in classes: one robot,
double-support standing balance, no viewer and no MuJoCo (those stay on the host and are out of scope).

    python examples/main_standing.py [n_ticks]

Every call below has the name and meaning it has in the reference script; the TSID objects are the mirrors of
tsid_control_b200/tsid_mirror.py and the arithmetic runs in libtsidb.so on cuda:0.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from tsid_control_b200.ctrl.conf import RobotConfig  # noqa: E402
from tsid_control_b200.ctrl.WalkController import WalkController  # noqa: E402

conf = RobotConfig()                                    # ref:main.py:46
controller = WalkController(conf)                       # ref:main.py:47
n_ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 500
q, v = controller.q.copy(), controller.v.copy()
q[7:] += 0.05                                           # start away from the posture reference
t = 0.0
for i in range(n_ticks):                                # ref:main.py:110
    HQPData = controller.formulation.computeProblemData(t, q, v)            # :119
    sol = controller.solver.solve(HQPData)                                  # :121
    if sol.status != 0:                                                     # :122
        print(f"QP problem could not be solved! Error code: {sol.status}")
        break
    tau = controller.formulation.getActuatorForces(sol)                     # :126
    dv = controller.formulation.getAccelerations(sol)                       # :127
    q, v = controller.integrate_dv(q, v, dv, conf.dt)                       # :128
    t += conf.dt
    if i % 100 == 0:
        com = controller.robot.com(controller.formulation.data())           # :135
        cop = controller.get_cop(sol)                                       # :132
        print(f"t={t:6.3f}  com={np.round(com, 4)}  cop={np.round(cop, 4)}  |tau|max={np.abs(tau).max():.3f}  iters={sol.iterations}")
print("final joint error to the posture reference:", float(np.abs(q[7:] - controller.q0[7:]).max()))
