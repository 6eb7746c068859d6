#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (gpurun_out/bench_<tag>.log) and the matching pytest log."""
import json, sys, os
tag = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
p = os.path.join(root, f"pytest_{tag}.log")
if os.path.exists(p):
    lines = open(p).read().splitlines()
    print("pytest:", lines[-1] if lines else "?")
    for l in lines:
        if l.startswith("FAILED") or l.startswith("ERROR"):
            print("  ", l[:200])
l = [x for x in open(os.path.join(root, f"bench_{tag}.log")) if x.startswith("{")]
if not l:
    print("no bench line; stderr tail:"); print(open(os.path.join(root, f"bench_{tag}.err")).read()[-1500:]); sys.exit()
d = json.loads(l[-1])
r = d["roofline"]
print(f"value {d['value']/1e6:.2f} M ticks/s  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']/1e6:.2f} M  lat1 {d.get('tick_latency_1env_us_p50')}")
for k in ("e2e_device_refs",):
    if k in d: print(k, f"{d[k]['value']/1e6:.2f} M", {a: b for a, b in d[k].items() if a.endswith('bytes_per_step')})
print("kernel ms:", {k: round(v, 4) for k, v in r["kernel_ms_all"].items()})
print(f"roofline: {r['kernel']} frac {r['frac']:.4f}  tick frac {r['tick']['frac']:.4f}  peak {r['peak']:.2f} TF")
if "cpu_baseline" in d: print("cpu:", round(d["cpu_baseline"]["value"]), "ticks/s on", d["cpu_baseline"]["cores"], "threads")
print("clocks:", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
