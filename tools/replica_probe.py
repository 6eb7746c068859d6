"""Does a single env's tick get faster when the same launch carries copies of it on the other SMs?  (B300_MICROARCH.md reports an
issue throttle for low-grid kernels with large bodies that vanishes at grid >= 148.)  GPU time of one tick of n identical envs."""
import sys, os, statistics, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from tsid_control_b200.ctrl.conf import RobotConfig
from tsid_control_b200.ctrl.WalkController import WalkController
from tsid_control_b200 import synth

out = {}
for n in (1, 2, 8, 37, 148, 296):
    conf = RobotConfig(); conf.max_envs = n
    c = WalkController(conf, n_envs=n); e = c.engine
    q, v = synth.random_states(c.q, 8, 3)
    mask, refs = synth.walking_batch(c.default_refs, 8, 5, 0.3, 0.2, 0.2, 0.5, float(c.default_refs["com"][2]))
    res = {}
    for pick in range(3):
        qd = torch.as_tensor(np.repeat(q[pick:pick + 1], n, 0), device=c.device)
        vd = torch.as_tensor(np.repeat(v[pick:pick + 1], n, 0), device=c.device)
        m = torch.as_tensor(np.repeat(mask[pick:pick + 1], n, 0), device=c.device)
        r = {k: torch.as_tensor(np.ascontiguousarray(np.repeat(a[pick:pick + 1], n, 0)), device=c.device) for k, a in refs.items()}
        for _ in range(20): e.compute(qd, vd, m, r)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gs = []
        for _ in range(100):
            ev0.record(); e.compute(qd, vd, m, r); ev1.record(); torch.cuda.synchronize(); gs.append(ev0.elapsed_time(ev1) * 1e3)
        res[f"env{pick}_mask{int(mask[pick])}"] = round(statistics.median(gs), 1)
    out[n] = res
    e.close()
print(json.dumps(out, indent=1))
