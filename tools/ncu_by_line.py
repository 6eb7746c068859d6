#!/usr/bin/env python
"""Join an ncu SASS-level source page (CSV) with nvdisasm line info of the same cubin and aggregate
stall samples / executed instructions per CUDA source line range (kernel phase).

usage: tools/ncu_by_line.py <report.ncu-rep> <lib.so> [kernel_name]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

PHASES = None


def phase_table(src):
    """(start_line, name) from marker comments / function heads in tsidb_kernels.cuh"""
    marks = []
    pats = [
        (r"^TSIDB_DEV void log6_dev", "log6"), (r"^TSIDB_DEV void k1_dynamics", "K1 dynamics"),
        (r"^TSIDB_DEV void se3_rhs", "K2 se3_rhs"), (r"^TSIDB_DEV void k2_assemble", "K2 assemble"),
        (r"^TSIDB_DEV int fvar0", "K3 row helpers"), (r"^TSIDB_DEV void qp_delete", "K3 delete_constraint"),
        (r"^TSIDB_DEV void backsub_LT", "E backsub_LT"), (r"^TSIDB_DEV void fwdsub_L", "E fwdsub_L"),
        (r"^TSIDB_DEV void reflect", "E reflect"), (r"^TSIDB_DEV int k3_eliminate", "E setup"),
        (r"Cholesky of the dv block: lane i keeps row", "E cholesky"), (r"B = L\^-1 \[CE\^T \| g\]", "E build B + QR"),
        (r"w_hat\[0:neq\] = R1\^-T", "E w_hat"), (r"w0 = Q w_hat", "E w0"), (r"x0 = L\^-T w0", "E x0"),
        (r"J2\[:, c\] = L\^-T Q", "E J2 columns"),
        (r"^TSIDB_DEV int as_solve", "K3 AS setup"), (r"for \(;;\) \{ /\* l1 \*/", "K3 AS l1: s, psi"),
        (r"for \(;;\) \{ /\* l2 \*/", "K3 AS l2: pick"), (r"for \(;;\) \{ /\* l2a \*/", "K3 AS l2a head"),
        (r"/\* s = CI x \+ ci0 for the rows this lane owns, one", "K3 AS l1: candidates s = CI x + ci0"),
        (r"violation sum: the other side", "K3 AS l1: pick, psi"),
        (r"/\* d = J2\^T n_ip", "K3 AS l2a: d = J2^T n"), (r"/\* z = J2\[:, iq:\] d\[iq:\]", "K3 AS l2a: z = J2 d"),
        (r"/\* r = R\^-1 d\[0:iq\]", "K3 AS l2a: r = R^-1 d"), (r"/\* partial step t1 = min", "K3 AS l2a: t1, t2, step"),
        (r"/\* full step: add ip", "K3 AS add (Householder update of J2, R column)"), (r"eiquadprog: exclude ip, restore", "K3 AS degenerate restore"),
        (r"dual-only step or partial step: drop the blocking", "K3 AS drop (shift R, Givens on R and J2)"),
        (r"/\* partial step: recompute s\(ip\)", "K3 AS partial step: s(ip) again"),
        (r"^TSIDB_DEV void dynamics_env", "D io + image stores"), (r"^TSIDB_DEV void eliminate_env", "E io + image stores"),
        (r"^struct G2Pipe", "G pipeline"), (r"^TSIDB_DEV int warp_argmin", "K3 argmin"), (r"^TSIDB_DEV unsigned smem_u32", "TMA helpers"), (r"^TSIDB_DEV void activeset_env", "A load + decode"),
        (r"^TSIDB_DEV void j2_columns", "G columns"), (r"^TSIDB_DEV void j2_env", "G load"),
        (r"^TSIDB_DEV (void|double) wrench_of", "K3 wrench_of"), (r"^TSIDB_DEV double eval_one", "K3 eval_one"), (r"^TSIDB_DEV void actuation_normal", "K3 actuation_normal"),
        (r"w0 = Q w_hat: reflectors in reverse", "E w0"), (r"x0 = L\^-T w0: force rows", "E x0"), (r"^__global__ void tsidb_classify", "kernel loops"),
        (r"^TSIDB_DEV void eval_rows", "K3 eval rows"), (r"^TSIDB_DEV double row_dot_col", "K3 row_dot_col"), (r"^TSIDB_DEVNI void qp_delete", "K3 delete_constraint"),
    ]
    for i, line in enumerate(open(src), 1):
        for p, n in pats:
            if re.search(p, line):
                marks.append((i, n))
                break
    return sorted(marks)


def main():
    rep, so = sys.argv[1], sys.argv[2]
    kern = sys.argv[3] if len(sys.argv) > 3 else "tsidb_tick_kernel"
    src = os.path.join(os.path.dirname(so), "tsidb_kernels.cuh")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # instruction index -> (file, line)
    lines_of = []
    cur = (None, 0)
    inside = False
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
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            lines_of.append(cur)
    filt = os.environ.get("NCU_KERNEL_FILTER")
    kid = os.environ.get("NCU_KERNEL_ID")  # n-th result of the report (1-based): template instantiations share a base name
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", f":::{kid}"] if kid else (["-k", "regex:" + filt] if filt else []))
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[0].startswith("0x")]
    if len(body) >= 2 * len(lines_of) and len(lines_of):
        body = body[: len(body) // (len(body) // len(lines_of))]  # several launches of the kernel in the report: first one
    if len(body) != len(lines_of):
        print(f"warning: {len(body)} SASS rows in the report, {len(lines_of)} in the cubin", file=sys.stderr)
    marks = phase_table(src)
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = tot_i = 0
    per_line = defaultdict(lambda: [0, 0])
    exc_line = defaultdict(lambda: [0, 0])  # excessive / total shared wavefronts per source line
    for k, r in enumerate(body[: len(lines_of)]):
        f, ln = lines_of[k]
        smp = int(r[ix["# Samples"]] or 0)
        ins = int(r[ix["Instructions Executed"]] or 0)
        wf = int(r[ix["L1 Wavefronts Shared"]] or 0) if "L1 Wavefronts Shared" in ix else 0
        name = "other"
        if f == "tsidb_kernels.cuh":
            for s, n in marks:
                if ln >= s:
                    name = n
        a = agg[name]
        a[0] += smp; a[1] += ins; a[2] += wf
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                a[3][c] += v
        tot_s += smp; tot_i += ins
        per_line[(f, ln)][0] += smp; per_line[(f, ln)][1] += ins
        if "L1 Wavefronts Shared Excessive" in ix:
            exc_line[(f, ln)][0] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0); exc_line[(f, ln)][1] += wf
    print(f"{'phase':50s} {'samples%':>8s} {'inst%':>7s} {'smem wf (M)':>11s}  top stalls")
    for name, (smp, ins, wf, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{name:50s} {100*smp/max(1,tot_s):8.1f} {100*ins/max(1,tot_i):7.1f} {wf/1e6:11.1f}  " +
              ", ".join(f"{k[6:]} {100*v/max(1,smp):.0f}%" for k, v in top))
    if os.environ.get("NCU_DUMP_LINES"):
        print("\nall source lines with >= 0.15% of the executed instructions, by line:")
        for (f, ln), (smp, ins) in sorted(per_line.items()):
            if ins >= 0.0015 * tot_i:
                print(f"  {f}:{ln:5d}  inst {ins/1e6:8.2f}M {100*ins/max(1,tot_i):5.1f}%  samples {100*smp/max(1,tot_s):5.1f}%")
        print(f"  total inst {tot_i/1e6:.1f}M")
    if os.environ.get("NCU_DUMP_CONFLICTS"):
        tot_e = sum(v[0] for v in exc_line.values()); tot_w = sum(v[1] for v in exc_line.values())
        print(f"\nshared-memory wavefronts {tot_w/1e6:.1f}M, excessive (bank conflicts) {tot_e/1e6:.2f}M = {100*tot_e/max(1,tot_w):.1f}%; lines with the most excessive wavefronts:")
        for (f, ln), (e, w) in sorted(exc_line.items(), key=lambda kv: -kv[1][0])[:20]:
            if e:
                print(f"  {f}:{ln:5d}  excessive {e/1e6:7.2f}M of {w/1e6:7.2f}M")
    print("\nhottest source lines:")
    for (f, ln), (smp, ins) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"  {f}:{ln:5d}  samples {100*smp/max(1,tot_s):5.1f}%  inst {100*ins/max(1,tot_i):5.1f}%")




def sass_size(so, kern="tsidb_tick_kernel"):
    """SASS instruction count per phase (code size), no report needed: tools/ncu_by_line.py --size <lib.so>"""
    src = os.path.join(os.path.dirname(so), "tsidb_kernels.cuh")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    marks = phase_table(src)
    cnt = defaultdict(int)
    cur, inside, tot = (None, 0), False, 0
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
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            name = "other"
            if cur[0] == "tsidb_kernels.cuh":
                for s, n in marks:
                    if cur[1] >= s:
                        name = n
            cnt[name] += 1
            tot += 1
    for n, c in sorted(cnt.items(), key=lambda kv: -kv[1]):
        print(f"{n:34s} {c:7d} instr  {c*16/1024:7.1f} KB")
    print(f"{'total':34s} {tot:7d} instr  {tot*16/1024:7.1f} KB")


if __name__ == "__main__":
    if sys.argv[1] == "--size":
        sass_size(sys.argv[2])
    else:
        main()
