import sys, os, time, statistics
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from tsid_control_b200.ctrl.conf import RobotConfig
from tsid_control_b200.ctrl.WalkController import WalkController
from tsid_control_b200 import synth
for n in (1, 32, 1024):
    conf = RobotConfig(); conf.max_envs = n
    c = WalkController(conf, n_envs=n); e = c.engine
    q, v = synth.random_states(c.q, n, 3)
    qd, vd = torch.as_tensor(q, device=c.device), torch.as_tensor(v, device=c.device)
    m = torch.full((n,), 3, dtype=torch.uint8, device=c.device)
    for _ in range(20): e.compute(qd, vd, m, c.refs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); e.compute(qd, vd, m, c.refs); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    e.set_timing(True)
    ks = []
    for _ in range(20):
        e.compute(qd, vd, m, c.refs); ks.append(e.last_tick_ms())
    e.set_timing(False)
    med = {k: statistics.median(x[k] for x in ks) * 1e3 for k in ks[0]}
    print(n, "wall p50 us %.1f" % (statistics.median(ts) * 1e6), "gpu us", {k: round(v, 1) for k, v in med.items()}, "sum %.1f" % sum(med.values()))
    e.close()
