#!/usr/bin/env python
"""Write the judged summary of an `ncu --set full` report: per kernel the duration, DRAM bytes, issue / pipe
utilisation, shared-memory wavefronts, stall mix and thread-instructions per source pixel.
Usage: ncu_summary.py REP MPIX_PER_LAUNCH OUT.json"""
import csv, json, subprocess, sys

rep, mpix, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"]
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("irp::", "")
    e = {k: {"unit": u.get(k, ""), "value": d.get(k)} for k in keys if k in d}
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v}
    tot = sum(st.values()) or 1.0
    e["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
    e["thread_instructions_per_source_pixel"] = float(d["smsp__inst_executed.sum"]) * 32 / (mpix * 1e6)
    to_b = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    e["dram_bytes_per_launch"] = (float(d["dram__bytes_read.sum"]) * to_b.get(u["dram__bytes_read.sum"], 1) +
                                  float(d["dram__bytes_write.sum"]) * to_b.get(u["dram__bytes_write.sum"], 1))
    res[name] = e
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: {"us": v["gpu__time_duration.sum"]["value"], "instr/px": round(v["thread_instructions_per_source_pixel"], 2),
                      "dram_MB": round(v["dram_bytes_per_launch"] / 1e6, 1)} for k, v in res.items()}))
