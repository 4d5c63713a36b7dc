#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libirp_b200.so, how often the instructions occur that prove TMA loads
(UTMALDG), bulk copies (UBLKCP), tcgen05 MMAs and tensor-memory traffic (UTCIMMA, UTCBAR, LDTM, STTM), mbarrier
transactions (SYNCS) and the packed integer arithmetic (IDP.4A / IDP.2A), with the first two of each kind verbatim.
    cuobjdump -sass image-restoration-platform_b200/libirp_b200.so | python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import re
import subprocess
import sys

WANT = re.compile(r"UTMALDG|UTMASTG|UBLKCP|UTCIMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS\.ARRIVE\.TRANS64|SYNCS\.PHASECHK|FENCE\.VIEW\.ASYNC|IDP\.4A|IDP\.2A|ATOMS\.POPC")
text = sys.stdin.read()
counts, first, fn = {}, [], None
for line in text.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    k = WANT.search(line)
    if k and fn:
        c = counts.setdefault(fn, {})
        c[k.group(0)] = c.get(k.group(0), 0) + 1
        if c[k.group(0)] <= 2:
            first.append((fn, line.strip()))
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass image-restoration-platform_b200/libirp_b200.so | python tools/sass_excerpts.py")
print(f"# architectures in the fatbin: {sorted(set(re.findall(r'arch = (sm_[0-9a-z]+)', text)))}\n")
for fn, name in zip(counts, names):
    print(f"== {name}\n   counts: {dict(sorted(counts[fn].items()))}")
    for g, line in first:
        if g == fn:
            print("     " + line)
    print()
