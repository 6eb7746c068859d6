#!/usr/bin/env python
"""Closed-loop batched rollout on the device: N robots, per-env gait phase and velocity command, tick -> integrate
-> gait phase machine replayed as one CUDA graph, diagnostics at the end.

    python examples/batched_walk.py [n_envs] [n_ticks]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from tsid_control_b200 import synth  # noqa: E402
from tsid_control_b200.ctrl.conf import RobotConfig  # noqa: E402
from tsid_control_b200.ctrl.WalkController import WalkController  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
conf = RobotConfig()
conf.max_envs = n
ctrl = WalkController(conf, n_envs=n)
dev = ctrl.device
rng = np.random.default_rng(0)
q, v = synth.random_states(ctrl.q, n, 0)
qd, vd = torch.as_tensor(q, device=dev), torch.as_tensor(0.1 * v, device=dev)
phase0 = torch.as_tensor(rng.uniform(0, 1, n), device=dev)
vcmd = torch.as_tensor(np.c_[rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
tau, ddq, f = ctrl.rollout(qd, vd, ticks, phase0=phase0, vcmd=vcmd)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gs = ctrl.engine.gait_state()
print(f"{n} envs x {ticks} ticks in {dt * 1e3:.1f} ms  ({n * ticks / dt / 1e6:.2f} M ticks/s closed loop)")
print("contact classes now:", {int(k): int((gs['mask'] == k).sum()) for k in (1, 2, 3)},
      " envs with a failed QP so far:", int((gs["fails"] > 0).sum()))
out = ctrl.engine.compute(qd, vd, gs["mask"].clone(), {k: gs[k] for k in ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf")}, aux=True)
d = ctrl.engine.diagnostics(out, gs["mask"], float(np.sqrt(9.80665 / ctrl.default_refs["com"][2])))
print("mean CoP:", d["cop"].mean(0).cpu().numpy().round(4), " mean capture point:", d["capture_point"].mean(0).cpu().numpy().round(4))
