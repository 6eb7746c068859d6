import sys, os, time, statistics
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from tsid_control_b200.ctrl.conf import RobotConfig
from tsid_control_b200.ctrl.WalkController import WalkController
from tsid_control_b200 import synth
n=1
conf = RobotConfig(); conf.max_envs = 65536
c = WalkController(conf, n_envs=65536); e = c.engine
q, v = synth.random_states(c.q, n, 3)
qd, vd = torch.as_tensor(q, device=c.device), torch.as_tensor(v, device=c.device)
mask, refs = synth.walking_batch(c.default_refs, n, 5, 0.3, 0.2, 0.2, 0.5, float(c.default_refs["com"][2]))
m = torch.as_tensor(mask, device=c.device)
rd = {k: torch.as_tensor(np.ascontiguousarray(a), device=c.device) for k, a in refs.items()}
for _ in range(50): e.compute(qd, vd, m, rd)
torch.cuda.synchronize()
host=[]; wall=[]
for _ in range(300):
    torch.cuda.synchronize()
    t0=time.perf_counter(); e.compute(qd, vd, m, rd); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    host.append(t1-t0); wall.append(t2-t0)
print("host-side call p50 %.1f us, wall p50 %.1f us" % (statistics.median(host)*1e6, statistics.median(wall)*1e6))
wall=[]
for _ in range(300):
    torch.cuda.synchronize()
    t0=time.perf_counter(); e.compute(qd, vd, m, None); torch.cuda.synchronize(); t2=time.perf_counter()
    wall.append(t2-t0)
print("default refs: wall p50 %.1f us" % (statistics.median(wall)*1e6))
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
for _ in range(2000): e.compute(qd, vd, m, rd)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
