#!/usr/bin/env python
"""bench.py — TSID QP ticks/s (robot/v1, batch 65536 per GPU) and tick latency.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One "step" = one TSID tick (computeProblemData + solve + decode, ref:main.py:119-127) of every env of the
batch.  Workload = BASELINE.json configs[2]: robot/v1 LIPM walking, 65536 envs per GPU with per-env contact
phases (20 % double support, 40 % / 40 % single support), swing-foot and CoM references and random
perturbed states (seed 0).  Weak scaling: every rank owns its own 65536 envs; the only
collective is the all-gather of the per-tick diagnostics.

Rank 0 prints ONE JSON line.  `value` is measured with inputs resident in HBM, CUDA events around every
step on the launching stream, an L2 flush between steps (outside the events), max over ranks.  `e2e` is the
same tick through the host-buffer entry point tsidb_compute_host (pinned staging, H2D, kernels, D2H inside
the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BATCH = 65536
METRIC = "tsid_qp_ticks_per_sec"
UNIT = "ticks/s"
WORKLOAD = "robot/v1 LIPM walking (BASELINE configs[2]): 65536 envs/GPU, 20% DS / 40% SS-L / 40% SS-R, seed 0"


# ----------------------------------------------------------------------------------------------
def algorithmic_flops(mask: np.ndarray, iters: np.ndarray) -> dict:
    """F/tick = F_dyn + F_asm + F_fact(n) + equality phase + per_iter * it (1 FMA = 2 flop);
    it = active-set iterations that attempted a constraint change = iterations - 1 (the last pass only checks
    feasibility).  DS 0.50 M + 0.020 M it, SS 0.22 M + 0.012 M it, flight from the same formula (n = 26,
    m_e = 6).  Split by the kernel that does the work: dynamics+assembly 28 k (F_dyn 20 k + F_asm 8 k), the
    factorisation and equality phase (elimination + null-space-basis kernels), the iterations (active set)."""
    nc = (mask & 1) + ((mask >> 1) & 1)
    base = np.where(nc == 2, 0.50e6, np.where(nc == 1, 0.22e6, 0.075e6))
    per = np.where(nc == 2, 0.020e6, np.where(nc == 1, 0.012e6, 0.0075e6))
    it = np.maximum(iters.astype(np.float64) - 1.0, 0.0)
    dyn = 28e3 * len(mask)
    return {"dynamics": float(dyn), "eliminate+j2": float(base.sum() - dyn), "activeset": float((per * it).sum()),
            "tick": float((base + per * it).sum())}


HBM_BYTES_PER_TICK = 1832.0  # 1240 B in + 592 B out (v1, double support)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def make_workload(n: int, seed: int, q_stand: np.ndarray, refs0: dict):
    from tsid_control_b200 import synth

    q, v = synth.random_states(q_stand, n, seed)
    mask, refs = synth.walking_batch(refs0, n, seed, 0.3, 0.2, 0.2, 0.5, float(refs0["com"][2]))
    return q, v, mask, refs


def run_reference(args) -> None:
    """The reference algorithm (restated CPU port: oracle/, -O3 AVX2 build) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from common import setup

    s = setup("v1", "liboracle_fast.so")
    cores = os.cpu_count() or 1
    sample = max(256, min(8192, 512 * cores))
    q, v, mask, refs = make_workload(sample, 0, s["q0"], s["refs"])
    run = s["oracle"].timed_batch(q, v, mask, refs, cores)
    for _ in range(max(1, args.warmup)):
        run()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    val = sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_envs_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} envs of the workload per step, {cores} threads, restated CPU port of "
                                   "pinocchio+tsid+eiquadprog-fast (oracle/, -O3 x86-64-v3); the reference's own binaries "
                                   "cannot be installed in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="envs per GPU (the metric is quoted at 65536)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end and latency legs (the line then has no e2e)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the TSID tick has no CPU fallback (use --impl reference for the CPU port)")
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from tsid_control_b200.ctrl.conf import RobotConfig
    from tsid_control_b200.ctrl.WalkController import WalkController
    from tsid_control_b200.engine import fp64_peak_tflops
    from tsid_control_b200.sharding import gather_diagnostics

    n = args.batch
    conf = RobotConfig()
    conf.device, conf.max_envs = local, n
    ctrl = WalkController(conf, n_envs=n)
    eng = ctrl.engine
    dev = ctrl.device
    # every rank owns a different shard of the global env range (seed offset by rank)
    q, v, mask, refs = make_workload(n, 0 + 1000 * rank, ctrl.q, ctrl.default_refs)
    qd, vd = torch.as_tensor(q, device=dev), torch.as_tensor(v, device=dev)
    ctrl.contact_mask = torch.as_tensor(mask, device=dev)
    ctrl.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in refs.items()}
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MiB > 126 MB L2
    diag_out = torch.empty((world * n, 2), dtype=torch.int32, device=dev) if world > 1 else None

    def step():
        out = ctrl._tick(qd, vd)
        if world > 1:
            gather_diagnostics(out.status, out.iters, diag_out)
        return out

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    fp64_peak = fp64_peak_tflops(local)

    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = eng.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.zero_()  # evict the inputs from L2 (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    iters_np = out.iters.cpu().numpy()
    status_np = out.status.cpu().numpy()
    flops = algorithmic_flops(mask, iters_np)
    step_ms = statistics.mean(ms)
    # per-kernel durations: CUDA events recorded by the library between its launches, on the launching stream,
    # in separate (untimed) steps with the same L2 flush so the events do not perturb `value`
    eng.set_timing(True)
    per_kernel = {}
    reps = max(3, min(args.steps, 10))
    for _ in range(reps):
        flush.zero_()
        step()
        for k, t in eng.last_tick_ms().items():
            per_kernel.setdefault(k, []).append(t)
    eng.set_timing(False)
    kms = {k: statistics.mean(v) for k, v in per_kernel.items()}
    groups = {"dynamics": kms["dynamics"], "eliminate+j2": kms["eliminate"] + kms["j2"], "activeset": kms["activeset"]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = HBM_BYTES_PER_TICK * n / (step_ms * 1e-3) / 1e9
    kernels = []
    for g, t in groups.items():
        tf = flops[g] / (t * 1e-3) / 1e12
        kernels.append({"kernel": g, "ms": t, "algorithmic_flops_per_launch": flops[g], "achieved": tf,
                        "frac": tf / fp64_peak if fp64_peak else None, "traffic": traffic.get(g)})
    dom = max(kernels, key=lambda k: k["ms"])
    tick_tf = flops["tick"] / (step_ms * 1e-3) / 1e12

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the call) ----
    # inputs live in pinned host memory (the contract's "from pinned host memory"), results land in pinned
    # host memory; the library cuts the batch into chunks so the copies run under the kernels
    e2e_steps = max(3, min(args.steps, 10)) if not args.no_e2e else 1
    hq, hv, hmask = eng.pin(q), eng.pin(v), eng.pin(mask)
    hrefs = {k: eng.pin(a) for k, a in refs.items()}
    hout = eng.host_buffers(n, pinned=True)
    for _ in range(2):
        eng.compute_host(hq, hv, hmask, hrefs, out=hout)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ho = eng.compute_host(hq, hv, hmask, hrefs, out=hout)
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(ho["status"], status_np) and np.array_equal(ho["iters"], iters_np), "e2e path disagrees with the device path"
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    # the sampler ran from the start of the timed region to here: the GPU was under the same tick load throughout
    # (timed steps, per-kernel timing steps, end-to-end steps), which gives nvidia-smi time for several samples
    clocks = sampler.stop()
    clocks["window"] = "timed steps + per-kernel timing steps + end-to-end steps"
    e2e_val = world * n * e2e_steps / t_e2e
    na, nv, nq = eng.na, eng.nv, eng.nq
    h2d = n * (8 * (nq + nv + 9 + 24 + 24 + 12 + 12 + na) + 1)
    d2h = n * (8 * (na + nv + 24) + 4 + 4 + 24)

    # ---- single-env tick latency (the reference's own operating point: one robot per call) ----
    lat = None
    if rank == 0 and not args.no_e2e:
        q1, v1 = qd[:1].contiguous(), vd[:1].contiguous()
        ts = []
        m1 = ctrl.contact_mask[:1].contiguous()
        r1 = {k: t[:1].contiguous() for k, t in ctrl.refs.items()}
        for i in range(220):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o1 = eng.compute(q1, v1, m1, r1)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        lat = statistics.median(ts[20:]) * 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "p50_ms_per_step": statistics.median(ms), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n, "global_envs": world * n, "l2_flush_between_steps": True,
                   "timing": "CUDA events per step on the launching stream, flush outside the events, max over ranks",
                   "collective": "all_gather of int32[N_local,2] diagnostics per step" if world > 1 else "none (1 GPU)"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "tsidb_compute_host via TsidEngine.compute_host: pinned host buffers in and out, 4 chunks over 3 streams (copies overlap the kernels)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": dom["achieved"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": dom["frac"],
                     "traffic": dom["traffic"], "kernel": dom["kernel"], "kernel_ms": dom["ms"],
                     "algorithmic_flops_per_launch": dom["algorithmic_flops_per_launch"],
                     "peak_source": "measured on this GPU by tsidb_fp64_peak (dependent-free DFMA chains); MEASURED_PEAKS.json has no FP64 "
                                    "entry; the path is FP64-bound, not HBM- or tensor-bound (SURVEY.md §8d)",
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture in profiles/ "
                                       "(profiles/dram_traffic.json)" if traffic else None,
                     "tick": {"achieved": tick_tf, "frac": tick_tf / fp64_peak if fp64_peak else None, "ms": step_ms,
                              "algorithmic_flops_per_step": flops["tick"],
                              "launches": "class sort (2) + dynamics + eliminate + j2 + activeset"},
                     "kernels": kernels, "kernel_ms_all": kms,
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "algorithmic_bytes_per_tick": HBM_BYTES_PER_TICK,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "solver": {"mean_iters": float(iters_np.mean()), "max_iters": int(iters_np.max()), "status_optimal_frac": float((status_np == 0).mean())},
        "tick_latency_1env_us_p50": lat,
    }
    if not args.no_cpu_baseline:
        from common import setup

        s = setup("v1", "liboracle_fast.so")
        cores = os.cpu_count() or 1
        sample = max(256, min(8192, 512 * cores))
        run = s["oracle"].timed_batch(q[:sample], v[:sample], mask[:sample], {k: a[:sample] for k, a in refs.items()}, cores)
        run()
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 10.0 and reps < 50:
            run()
            reps += 1
        dt = time.perf_counter() - t0
        # single-thread latency of one tick
        run1 = s["oracle"].timed_batch(q[:64], v[:64], mask[:64], {k: a[:64] for k, a in refs.items()}, 1)
        run1()
        t1 = time.perf_counter()
        run1()
        lat_cpu = (time.perf_counter() - t1) / 64 * 1e6
        line["cpu_baseline"] = {"value": sample * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {sample} envs of the workload x {reps} passes on {cores} threads; restated CPU port "
                                          "(oracle/, -O3 x86-64-v3), not the reference binaries", "tick_latency_1thread_us": lat_cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
