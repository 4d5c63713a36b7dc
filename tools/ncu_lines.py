#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall mix, and the hottest CUDA source lines
(thread-instructions per source pixel).  Usage: ncu_lines.py REP SOURCE_FILE MPIX [top]"""
import csv, subprocess, sys

rep, src, mpix = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: v for h, v in zip(rows[0], rows[2])}
u = {h: v for h, v in zip(rows[0], rows[1])}
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"]:
    print(f"{k:75s} {d.get(k)} {u.get(k)}")
print("thread-instr per px:", float(d["smsp__inst_executed.sum"]) * 32 / (mpix * 1e6))
st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
tot = sum(st.values())
print("stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
lines = open(src).read().split("\n")
def toi(s):
    try: return int(s)
    except Exception: return 0
per, ts = [], 0
for r in rows[3:]:
    if r and r[0].strip().isdigit() and r[2] == "-":
        per.append((int(r[0]), toi(r[iI]), toi(r[iS]))); ts += toi(r[iS])
for n, inst, samp in sorted(per, key=lambda t: -t[1])[:top]:
    print(f"{n:4d} {inst * 32 / (mpix * 1e6):6.2f}/px {100 * samp / max(ts, 1):5.1f}%s  {lines[n - 1].strip()[:110] if n <= len(lines) else '?'}")
