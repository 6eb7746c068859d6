"""Where the time of a single env goes inside the single-launch kernel: builds libtsidb.so with TSIDB_SMALL_PROFILE=1 (the
kernel then leaves clock64 stamps of its stages for env 0), ticks single envs of the walking workload and prints the
median cycles per stage.  Restores the normal build afterwards.  GPU box only:  python tools/small_profile.py"""
import os, subprocess, sys, statistics, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
CSRC = os.path.join(os.getcwd(), "tsid_control_b200", "csrc")
NAMES = ["stage_model", "dynamics+handoff", "Lfinv staging", "eliminate+handoff", "lane_const_init", "active set + decode"]


def build(profile):
    import __graft_entry__ as ge
    flags = ge.NVCC_FLAGS + ([f"-DTSIDB_SMALL_PROFILE=1"] if profile else [])
    subprocess.run(["nvcc"] + flags + ["-o", os.path.join(CSRC, "libtsidb.so"), os.path.join(CSRC, "tsidb.cu")], check=True, cwd=CSRC)


def child():
    import numpy as np, torch
    from tsid_control_b200.ctrl.conf import RobotConfig
    from tsid_control_b200.ctrl.WalkController import WalkController
    from tsid_control_b200 import synth
    conf = RobotConfig(); conf.max_envs = 1
    c = WalkController(conf, n_envs=1); e = c.engine
    N = 48
    q, v = synth.random_states(c.q, N, 3)
    mask, refs = synth.walking_batch(c.default_refs, N, 5, 0.3, 0.2, 0.2, 0.5, float(c.default_refs["com"][2]))
    rows = {1: [], 2: []}
    for i in range(N):
        qd, vd = torch.as_tensor(q[i:i + 1], device=c.device), torch.as_tensor(v[i:i + 1], device=c.device)
        m = torch.as_tensor(mask[i:i + 1], device=c.device)
        r = {k: torch.as_tensor(np.ascontiguousarray(a[i:i + 1]), device=c.device) for k, a in refs.items()}
        for _ in range(3): e.compute(qd, vd, m, r)
        torch.cuda.synchronize()
        nc = (int(mask[i]) & 1) + (int(mask[i]) >> 1)
        st = e.debug_terms(0, nc)["H"][0, :8]
        rows[nc].append(np.diff(st[:7]))
    out = {}
    for nc, rr in rows.items():
        if rr:
            med = np.median(np.array(rr), axis=0)
            out[f"{nc} contacts ({len(rr)} envs)"] = {n: int(x) for n, x in zip(NAMES, med)} | {"total": int(med.sum())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        try:
            build(True)
            for loc in (None, "0"):
                env = dict(os.environ)
                if loc is not None: env["TSIDB_SMALL_LOCAL_N"] = loc
                print("TSIDB_SMALL_LOCAL_N =", loc or "default", flush=True)
                subprocess.run([sys.executable, __file__, "child"], env=env)
        finally:
            build(False)
