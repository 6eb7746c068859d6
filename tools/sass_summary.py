#!/usr/bin/env python
"""Per-kernel SASS evidence of libtsidb.so: instruction count and the opcodes that matter on this path (FP64 math, 1-D TMA
bulk copies, mbarrier waits, L2 prefetch, warp reductions): tools/sass_summary.py > profiles/sass_r02_summary.txt"""
import os, re, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tsid_control_b200", "csrc", "libtsidb.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = defaultdict(Counter); name = None
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and name:
        ops[name][m.group(1)] += 1
KEYS = ["DFMA", "DMUL", "DADD", "MUFU.RCP64H", "MUFU.RSQ64H", "UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "CREDUX", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "ATOM"]
print("SASS of", os.path.relpath(so, ROOT), "(cuobjdump -sass; sm_100a): instructions per kernel and opcode families")
print("UBLKCP = cp.async.bulk (1-D TMA), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier, LDGSTS = cp.async, CREDUX = warp reduce (redux.sync); LDL/STL = spills")
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{k.split('.')[-1][:6]:>6s}" for k in KEYS))
for n in sorted(ops):
    c = ops[n]; tot = sum(c.values())
    fam = [sum(v for k, v in c.items() if k.startswith(key)) for key in KEYS]
    print(f"{n[:64]:64s} {tot:6d} " + " ".join(f"{v:6d}" for v in fam))
