import os, sys, time, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from irp_b200.synth import synth_image
eng = irp_b200.Engine(0)
base = [synth_image(4000, 3000, idx=s) for s in range(4)]
small = eng.preprocess_batch(base)
dev = [eng.upload(small[i % 4], pitch_align=256) for i in range(64)]
for opt in (False, True, False, True):
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter(); files = eng.encode_jpeg_batch(dev, quality=85, optimize=opt); best = min(best, time.perf_counter() - t0)
    print("optimize", opt, f"{best*1e3:.2f} ms wall per 64 x 2048x1536 (pageable outputs)", f"{sum(map(len, files))/1e6:.1f} MB")
