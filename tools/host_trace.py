"""Print the per-chunk time line of tsidb_compute_host (TSIDB_HOST_TRACE=1): when the inputs of a chunk are on the device,
when its kernels end and when its outputs are back on the host, for the default split and two equal splits.

usage (GPU box): python tools/host_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import __graft_entry__ as ge
ge.build()
from tsid_control_b200.ctrl.conf import RobotConfig
from tsid_control_b200.ctrl.WalkController import WalkController
n = 65536
conf = RobotConfig(); conf.device, conf.max_envs = 0, n
ctrl = WalkController(conf, n_envs=n); eng = ctrl.engine
q, v, mask, refs = bench.make_workload("v1", "walking", n, 0, ctrl.q, ctrl.default_refs)
hq, hv, hmask = eng.pin(q), eng.pin(v), eng.pin(mask)
hrefs = {k: eng.pin(a) for k, a in refs.items()}
hout = eng.host_buffers(n, pinned=True)
for _ in range(3): eng.compute_host(hq, hv, hmask, hrefs, out=hout)
for env in ({}, {"TSIDB_HOST_CHUNKS": "8"}, {"TSIDB_HOST_CHUNKS": "2"}):
    for k in ("TSIDB_HOST_CHUNKS",): os.environ.pop(k, None)
    os.environ.update(env)
    eng.compute_host(hq, hv, hmask, hrefs, out=hout)
    os.environ["TSIDB_HOST_TRACE"] = "1"
    print("----", env, file=sys.stderr, flush=True)
    eng.compute_host(hq, hv, hmask, hrefs, out=hout)
    os.environ.pop("TSIDB_HOST_TRACE")
