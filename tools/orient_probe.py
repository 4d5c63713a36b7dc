#!/usr/bin/env python
"""Preprocess time per EXIF orientation (device-resident 12 MP batch) and single-image latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irp_b200
from irp_b200.synth import synth_batch

W, H, B = 4000, 3000, 16
imgs = synth_batch(W, H, B, distinct=2)
with irp_b200.Engine(0) as eng:
    for o in (1, 2, 3, 6, 8, 5):
        d_in = [eng.upload(im, orientation=o) for im in imgs]
        ow, oh = eng.preprocess_dims(W, H, o)
        d_out = [eng.alloc_device(ow, oh, 3) for _ in imgs]
        for _ in range(3):
            eng.preprocess_batch(d_in, device_outputs=d_out)
        t = []
        for _ in range(5):
            eng.preprocess_batch(d_in, device_outputs=d_out)
            t.append(eng.timing()["preprocess_ms"])
        print(f"orientation {o}: preprocess {min(t):.3f} ms per {B} x 12 MP  ({B * W * H / min(t) / 1e6:.1f} GPix/s) out {ow}x{oh}")
        for d in d_in + d_out:
            eng.free(d)
    # single image latency, device resident and host resident
    d1 = eng.upload(imgs[0]); o1 = eng.alloc_device(2048, 1536, 3)
    for _ in range(5): eng.analyze_batch([d1], device_outputs=[o1])
    t0 = time.perf_counter()
    for _ in range(20): eng.analyze_batch([d1], device_outputs=[o1])
    dt = (time.perf_counter() - t0) / 20
    tm = eng.timing()
    print(f"single 12 MP image, device resident: {dt * 1e3:.3f} ms wall per call (classify {tm['classify_ms']:.3f} ms, preprocess {tm['preprocess_ms']:.3f} ms)")
    for _ in range(3): eng.analyze_batch([imgs[0]])
    t0 = time.perf_counter()
    for _ in range(10): eng.analyze_batch([imgs[0]])
    dt = (time.perf_counter() - t0) / 10
    print(f"single 12 MP image, pageable host numpy in/out: {dt * 1e3:.3f} ms wall per call")
