#!/usr/bin/env python
"""Key per-kernel metrics of an ncu report (one line per captured launch)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("smsp__inst_executed.sum", "inst(M)"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wf(M)"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "conflicts(M)"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__cycles_elapsed.avg", "cycles")]
units = r[1]
print(" | ".join(n for _, n in want))
for row in r[2:]:
    cells = []
    for k, n in want:
        if k not in h: cells.append("-"); continue
        v = row[h.index(k)]; u = units[h.index(k)]
        try:
            f = float(v)
            if n in ("inst(M)", "smem wf(M)", "conflicts(M)"): v = f"{f/1e6:.1f}"
            elif n in ("rd", "wr"): v = f"{f:.3g}{u}"
            else: v = f"{f:.4g}"
        except ValueError:
            v = v.replace("void ", "").replace("(TickArgs)", "")[:34]
        cells.append(v)
    print(" | ".join(cells))
