#!/usr/bin/env python
"""SASS instruction count (code size) per source line of one kernel: tools/sass_lines.py <lib.so> <mangled kernel> [top]"""
import os, re, subprocess, sys, tempfile
from collections import defaultdict

so, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt = defaultdict(int); cur = ("?", 0); inside = False; tot = 0; ops = defaultdict(int)
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[cur] += 1; tot += 1; ops[m.group(1).split(".")[0]] += 1
print(f"total {tot} instr {tot*16/1024:.1f} KB")
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{c:6d}  {f}:{ln}")
print("opcodes:", ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]))
