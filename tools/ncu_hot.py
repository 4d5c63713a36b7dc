#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel in an .ncu-rep (captured with --import-source on, built with -lineinfo).
Usage: ncu_hot.py REP KERNEL_REGEX MPIX [top]   -> thread-instructions per source pixel and stall-sample share per line"""
import csv, re, subprocess, sys

rep, kern, mpix = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs, cur, path = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]
    elif r[0] == "Function Name":
        cur = {"fn": r[1], "path": path, "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
per, ts = [], 0
for sec in secs:
    if not re.search(kern, sec["fn"]) or not sec["rows"]:
        continue
    hdr = sec["rows"][0]
    if "Instructions Executed" not in hdr:
        continue
    iI, iN = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        lines = open(sec["path"]).read().split("\n")
    except Exception:
        lines = []
    for r in sec["rows"][1:]:
        if r[0].strip().isdigit() and len(r) > iI:
            try:
                n, ins, sm = int(r[0]), int(r[iI] or 0), int(r[iN] or 0)
            except ValueError:
                continue
            per.append((ins, sm, sec["path"].split("/")[-1], n, lines[n - 1].strip()[:100] if n <= len(lines) else r[1][:100]))
            ts += sm
tot = sum(p[0] for p in per)
print(f"total {tot * 32 / (mpix * 1e6):.2f} thread-instr/px over {len(per)} lines")
for ins, sm, f, n, txt in sorted(per, key=lambda t: -t[0])[:top]:
    print(f"{f[:22]:22s}{n:4d} {ins * 32 / (mpix * 1e6):6.2f}/px {100 * sm / max(ts, 1):5.1f}%s  {txt}")
