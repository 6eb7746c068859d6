#!/usr/bin/env python
"""Summarise an ncu report of the tick kernels into profiles/: headline metrics per kernel, the per-phase
stall table (tools/ncu_by_line.py) and the DRAM traffic per launch (profiles/dram_traffic.json, read by bench.py).

usage: tools/summarize_profile.py <report.ncu-rep> <tag> [launches.csv]
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
# short name -> (traffic group of bench.py, regex selecting the launch in the report, mangled name in the cubin)
KERNELS = {
    "dynamics": ("dynamics", r"tsidb_dynamics_kernel<26>", "_Z21tsidb_dynamics_kernelILi26EEv8TickArgs"),
    "eliminate_ds": ("eliminate", r"tsidb_eliminate_kernel<26, 2", "_Z22tsidb_eliminate_kernelILi26ELi2ELi8EEv8TickArgs"),
    "eliminate_ss": ("eliminate", r"tsidb_eliminate_kernel<26, 1", "_Z22tsidb_eliminate_kernelILi26ELi1ELi12EEv8TickArgs"),
    "j2_ds": ("eliminate", r"tsidb_j2_kernel<26, 2>", "_Z15tsidb_j2_kernelILi26ELi2EEv8TickArgs"),
    "j2_ss": ("eliminate", r"tsidb_j2_kernel<26, 1>", "_Z15tsidb_j2_kernelILi26ELi1EEv8TickArgs"),
    "activeset_ds": ("activeset", r"tsidb_activeset_kernel<26, 2", "_Z22tsidb_activeset_kernelILi26ELi2ELi8EEv8TickArgs"),
    "activeset_ss": ("activeset", r"tsidb_activeset_kernel<26, 1", "_Z22tsidb_activeset_kernelILi26ELi1ELi12EEv8TickArgs"),
}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"{tag}: ncu --set full --clock-control none, one launch of each tick kernel inside `python bench.py --steps 2 --warmup 3` "
           "(robot/v1 walking, 65536 envs: 13 107 double support, 52 429 single support); per-launch times under ncu are serialised and cold-cache, use them for shares only\n"]
    traffic = defaultdict(float)
    result_id = {}
    for idx, r in enumerate(rows[2:]):
        for k, (_, rx, _m) in KERNELS.items():
            if rx in r[hdr.index("Kernel Name")] and k not in result_id:
                result_id[k] = idx + 1
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append(f"Kernel {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"  {k:90s} {r[i]} {units[i]}")
        short = next((k for k, (_, rx, _m) in KERNELS.items() if rx in name), None)
        if short:
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic[KERNELS[short][0]] += float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
        out.append("")
    so = os.path.join(ROOT, "tsid_control_b200", "csrc", "libtsidb.so")
    for short, (_g, rx, mangled) in KERNELS.items():
        if short not in result_id:
            continue
        env = dict(os.environ, NCU_KERNEL_ID=str(result_id[short]))
        t = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, so, mangled], capture_output=True, text=True, env=env)
        out.append(f"---- {short}: warp-state samples and executed instructions by source phase (tools/ncu_by_line.py)")
        out.append(t.stdout if t.returncode == 0 else t.stderr[-400:])
    if len(sys.argv) > 3:
        lrows = list(csv.reader(l for l in open(sys.argv[3]) if l.startswith('"')))
        h = lrows[0]
        d = defaultdict(list)
        for r in lrows[1:]:
            d[r[h.index("Kernel Name")]].append(float(r[h.index("Metric Value")]))
        out.append("---- launch list (ncu --metrics gpu__time_duration.sum --clock-control none, same bench command): kernel, launches, median ns")
        tot = sum(sorted(v)[len(v) // 2] for k, v in d.items() if "tsidb_" in k and "dfma" not in k)
        for k, v in d.items():
            med = sorted(v)[len(v) // 2]
            share = f"{100 * med / tot:5.1f}% of the tick" if "tsidb_" in k and "dfma" not in k else ""
            out.append(f"  {k[:70]:70s} {len(v):4d} {med:12.0f} {share}")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_summary.txt"), "w").write("\n".join(out) + "\n")
    json.dump({**{k: v for k, v in traffic.items()}, "source": f"profiles/{tag}_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"},
              open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    print("\n".join(out)[:6000])


if __name__ == "__main__":
    main()
