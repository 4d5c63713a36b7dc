#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line of one kernel in an .ncu-rep.  Usage: ncu_smem.py REP KERNEL_REGEX [top]"""
import csv, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs, cur, path = [], None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1]
    elif r[0] == "Function Name": cur = {"fn": r[1], "path": path, "rows": []}; secs.append(cur)
    elif cur is not None: cur["rows"].append(r)
per = []
for sec in secs:
    if not re.search(kern, sec["fn"]) or not sec["rows"]: continue
    hdr = sec["rows"][0]
    if "L1 Wavefronts Shared" not in hdr: continue
    iW, iD, iI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
    try: lines = open(sec["path"]).read().split("\n")
    except Exception: lines = []
    for r in sec["rows"][1:]:
        if r[0].strip().isdigit() and len(r) > iW:
            try: n, w, d, ins = int(r[0]), int(r[iW] or 0), int(r[iD] or 0), int(r[iI] or 0)
            except ValueError: continue
            if w: per.append((w, d, ins, sec["path"].split("/")[-1], n, lines[n - 1].strip()[:90] if n <= len(lines) else ""))
tot = sum(p[0] for p in per)
print("total shared wavefronts", tot)
for w, d, ins, f, n, txt in sorted(per, key=lambda t: -t[0])[:top]:
    print(f"{f[:20]:20s}{n:4d} {100*w/tot:5.1f}%  ideal {100*d/max(w,1):4.0f}%  {txt}")
