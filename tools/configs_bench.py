#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE.json configs on one GPU (bench.py stays on configs[1]):
configs[2] fusion (32 triplets of 12 MP -> 2048x2048 canvases), configs[3] 4K (256 x 3840x2160),
configs[4] mixed-resolution queue (512 images, 0.5-24 MP).  Prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irp_b200
from irp_b200.synth import synth_image, synth_batch, mixed_resolution_sizes


def timed(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


with irp_b200.Engine(0) as eng:
    # configs[3]: 4K
    W, H, B = 3840, 2160, 256
    imgs = synth_batch(W, H, B, distinct=4)
    d_in = [eng.upload(im) for im in imgs]
    ow, oh = eng.preprocess_dims(W, H)
    d_out = [eng.alloc_device(ow, oh, 3) for _ in imgs]
    dt = timed(lambda: eng.analyze_batch(d_in, device_outputs=d_out, raw=True))
    t = eng.timing()
    print(json.dumps({"config": "configs[3]: 256 x 3840x2160 classify+preprocess, 1 GPU, device-resident", "MPix/s": B * W * H / dt / 1e6,
                      "ms": dt * 1e3, "classify_ms": t["classify_ms"], "preprocess_ms": t["preprocess_ms"]}))
    for d in d_in + d_out:
        eng.free(d)
    # configs[2]: fusion triplets (mixed aspect)
    shapes = [(4000, 3000), (3000, 4000), (3840, 2160)]
    base = [synth_image(w, h, i) for i, (w, h) in enumerate(shapes)]
    groups = [[eng.upload(np.roll(base[k], 13 * g, axis=0)) for k in range(3)] for g in range(32)]
    canv = [eng.alloc_device(2048, 2048, 3) for _ in range(96)]
    dt = timed(lambda: eng.fusion_prepare_batch(groups, device_outputs=canv))
    px = 32 * sum(w * h for w, h in shapes)
    print(json.dumps({"config": "configs[2]: 32 fusion triplets (4000x3000, 3000x4000, 3840x2160) -> 2048^2 canvases, device-resident",
                      "MPix/s": px / dt / 1e6, "ms": dt * 1e3, "triplets/s": 32 / dt}))
    for g in groups:
        for d in g:
            eng.free(d)
    for d in canv:
        eng.free(d)
    # configs[4]: mixed-resolution queue
    sizes = mixed_resolution_sizes(512)
    big = synth_image(6600, 6100, 11)   # width x height: covers 24 MP at every aspect of the set
    d_in, d_out, px = [], [], 0
    for i, (w, h) in enumerate(sizes):
        x0, y0 = (37 * i) % max(1, 6600 - w), (91 * i) % max(1, 6100 - h)
        assert x0 + w <= 6600 and y0 + h <= 6100, (w, h)
        d_in.append(eng.upload(np.ascontiguousarray(big[y0:y0 + h, x0:x0 + w])))
        o_w, o_h = eng.preprocess_dims(w, h)
        d_out.append(eng.alloc_device(o_w, o_h, 3))
        px += w * h
    dt = timed(lambda: eng.analyze_batch(d_in, device_outputs=d_out, raw=True), n=3)
    t = eng.timing()
    print(json.dumps({"config": "configs[4]: 512 images, 0.5-24 MP log-uniform, 6 aspects, classify+preprocess, device-resident",
                      "MPix/s": px / dt / 1e6, "ms": dt * 1e3, "total_MP": px / 1e6, "classify_ms": t["classify_ms"],
                      "preprocess_ms": t["preprocess_ms"], "launches": t["kernel_launches"]}))
