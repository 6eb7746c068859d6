"""Wall-clock and GPU latency of one tick for small batches, with the single-launch small-batch kernel
(TSIDB_SMALL_N, default 1024) and with the batched pipeline (TSIDB_SMALL_N=0): python tools/latency_probe.py"""
import sys, os, time, statistics, subprocess, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))


def probe():
    import numpy as np, torch
    import __graft_entry__ as ge
    ge.build()
    from tsid_control_b200.ctrl.conf import RobotConfig
    from tsid_control_b200.ctrl.WalkController import WalkController
    from tsid_control_b200 import synth
    out = {}
    for n in [int(x) for x in os.environ.get("PROBE_N", "1,32,256,1024,4096").split(",")]:
        conf = RobotConfig(); conf.max_envs = n
        c = WalkController(conf, n_envs=n); e = c.engine
        q, v = synth.random_states(c.q, n, 3)
        qd, vd = torch.as_tensor(q, device=c.device), torch.as_tensor(v, device=c.device)
        mask, refs = synth.walking_batch(c.default_refs, n, 5, 0.3, 0.2, 0.2, 0.5, float(c.default_refs["com"][2]))
        m = torch.as_tensor(mask, device=c.device)
        for _ in range(20): e.compute(qd, vd, m, c.refs)
        torch.cuda.synchronize()
        ts = []
        for _ in range(300):
            t0 = time.perf_counter(); e.compute(qd, vd, m, c.refs); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gs = []
        for _ in range(100):
            ev0.record(); e.compute(qd, vd, m, c.refs); ev1.record(); torch.cuda.synchronize(); gs.append(ev0.elapsed_time(ev1) * 1e3)
        out[n] = {"wall_p50_us": round(statistics.median(ts) * 1e6, 1), "gpu_p50_us": round(statistics.median(gs), 1)}
        e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        probe()
    else:
        for small in os.environ.get("PROBE_SMALL", "1024,0").split(","):
            r = subprocess.run([sys.executable, __file__, "child"], env={k: v for k, v in dict(os.environ, TSIDB_SMALL_N=small).items() if v != ""}, capture_output=True, text=True)
            print("TSIDB_SMALL_N=" + small, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:])
