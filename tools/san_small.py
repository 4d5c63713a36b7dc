#!/usr/bin/env python
"""A small pass through every kernel for compute-sanitizer (one tool per gpurun call):
streaming + generic classify, streaming + generic resize, tiled orientation, identity copy, fusion canvas."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irp_b200
from oracle import oracle
rng = np.random.default_rng(5)
def img(h, w, c=3): return rng.integers(0, 256, (h, w, c), dtype=np.uint8)
a, b, c4, g1, small = img(2200, 2304), img(2100, 2500), img(2100, 300, 4), img(300, 2100, 1), img(333, 517)
with irp_b200.Engine(0) as eng:
    res, outs = eng.analyze_batch([a, b, c4, g1, small], orientations=[1, 6, 3, 1, 8])
    canv = eng.fusion_prepare_batch([[a, small, g1]])
    d = eng.upload(img(601, 2101), pitch_align=1)      # unaligned rows: generic kernels
    r2, o2 = eng.analyze_batch([d])
ok = all(np.array_equal(o, oracle.preprocess(i, orr)) for o, i, orr in zip(outs, [a, b, c4, g1, small], [1, 6, 3, 1, 8]))
print("sanitizer pass ran; preprocess parity", ok, "scores", [round(v, 3) for v in res[0]["scores"].values()])
