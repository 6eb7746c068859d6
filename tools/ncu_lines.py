#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples of one kernel of an ncu report, in line order.
usage: NCU_KERNEL_ID=k tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel> [min_pct]"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, so, kern = sys.argv[1:4]
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines_of, ops, cur, inside = [], [], (None, 0), False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = l.startswith(f"\t.section\t.text.{kern},")
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m:
        lines_of.append(cur); ops.append(m.group(1))
kid = os.environ.get("NCU_KERNEL_ID")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[0].startswith("0x")]
if len(lines_of) and len(body) >= 2 * len(lines_of):
    body = body[: len(lines_of)]
assert len(body) == len(lines_of), (len(body), len(lines_of))
per = defaultdict(lambda: [0, 0, defaultdict(int)])
ti = ts = 0
for k, r in enumerate(body):
    smp = int(r[ix["# Samples"]] or 0); ins = int(r[ix["Instructions Executed"]] or 0)
    p = per[lines_of[k]]; p[0] += smp; p[1] += ins; p[2][ops[k]] += ins
    ti += ins; ts += smp
src = {}
for (f, ln) in per:
    if f not in src:
        for root in (os.path.dirname(so), "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                src[f] = open(pth, errors="replace").read().splitlines(); break
        else:
            src[f] = []
print(f"total instructions {ti}, samples {ts}")
for (f, ln), (smp, ins, opc) in sorted(per.items()):
    if 100 * ins / ti < minp and 100 * smp / ts < minp:
        continue
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ""
    top = ",".join(f"{o}:{100*c/ti:.1f}" for o, c in sorted(opc.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f[:18]:18s}:{ln:5d} inst {100*ins/ti:5.2f}% smp {100*smp/ts:5.2f}%  [{top}]  {text}")
