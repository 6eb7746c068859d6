#!/usr/bin/env python
"""Sweep the chunking knobs of tsidb_compute_host (TSIDB_HOST_CHUNKS: equal chunks, TSIDB_HOST_TAPER: denominator of
the small first/last chunk of the tapered 4-chunk split) on the bench workload and print end-to-end ticks/s.

usage (GPU box): python tools/e2e_sweep.py [batch]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def main():
    import torch

    ge.build()
    from tsid_control_b200.ctrl.conf import RobotConfig
    from tsid_control_b200.ctrl.WalkController import WalkController

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    conf = RobotConfig()
    conf.device, conf.max_envs = 0, n
    ctrl = WalkController(conf, n_envs=n)
    eng = ctrl.engine
    q, v, mask, refs = bench.make_workload("v1", "walking", n, 0, ctrl.q, ctrl.default_refs)
    hq, hv, hmask = eng.pin(q), eng.pin(v), eng.pin(mask)
    hrefs = {k: eng.pin(a) for k, a in refs.items()}
    hout = eng.host_buffers(n, pinned=True)

    def run(tag, env):
        for k in ("TSIDB_HOST_CHUNKS", "TSIDB_HOST_TAPER", "TSIDB_HOST_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            eng.compute_host(hq, hv, hmask, hrefs, out=hout)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(8):
                eng.compute_host(hq, hv, hmask, hrefs, out=hout)
            ts.append((time.perf_counter() - t0) / 8)
        best = min(ts)
        print(f"{tag:28s} {best * 1e3:7.3f} ms  {n / best / 1e6:7.3f} M ticks/s", flush=True)

    run("default", {})
    for sp in os.environ.get("SWEEP_SPLITS", "8,24,24,8 8,20,28,8 8,18,30,8 6,18,32,8 8,20,26,10 6,16,34,8 6,14,22,16,6 8,18,26,12 4,12,24,18,6 8,16,24,16").split():
        run(f"split {sp} /64", {"TSIDB_HOST_SPLIT": sp})
    for den in (5, 6, 8, 10, 12, 16, 24, 32):
        run(f"taper 1/{den}", {"TSIDB_HOST_TAPER": str(den)})
    for c in (1, 2, 3, 4, 5, 6, 8):
        run(f"equal chunks {c}", {"TSIDB_HOST_CHUNKS": str(c)})
    qd, vd = torch.as_tensor(q, device=ctrl.device), torch.as_tensor(v, device=ctrl.device)
    ctrl.contact_mask = torch.as_tensor(mask, device=ctrl.device)
    ctrl.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=ctrl.device) for k, a in refs.items()}
    for _ in range(3):
        ctrl._tick(qd, vd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ctrl._tick(qd, vd)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{'device-resident':28s} {dt * 1e3:7.3f} ms  {n / dt / 1e6:7.3f} M ticks/s")


if __name__ == "__main__":
    main()
