"""Tensor-core resize kernel (irp_resize_mma.cuh) against the oracle on a few geometries, with a diff summary."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from irp_b200.synth import synth_image
from oracle import oracle

shapes = [(4000, 3000), (3840, 2160), (2304, 2200), (6000, 4000), (2100, 2050), (3000, 4000), (4096, 2048), (5986, 3991), (8000, 2100), (4597, 4597)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
with irp_b200.Engine(0) as eng:
    for k, (w, h) in enumerate(shapes):
        img = synth_image(w, h, idx=k)
        t0 = time.time()
        got = eng.preprocess_batch([img])[0]
        t1 = time.time()
        ref = oracle.preprocess(img, 1)
        if got.shape != ref.shape:
            print(f"{w}x{h}: shape {got.shape} vs {ref.shape}")
            continue
        d = got.astype(int) - ref.astype(int)
        bad = np.argwhere(d != 0)
        print(f"{w}x{h} -> {got.shape[1]}x{got.shape[0]}: {len(bad)} bytes differ (max |d| {np.abs(d).max()}), call {1e3 * (t1 - t0):.1f} ms", flush=True)
        if len(bad):
            ys, xs = np.unique(bad[:, 0]), np.unique(bad[:, 1])
            print("   rows", ys[:12], "... cols", xs[:12], "... channels", np.unique(bad[:, 2]))
            print("   first:", bad[:6].tolist(), "got", [int(got[tuple(b)]) for b in bad[:6]], "ref", [int(ref[tuple(b)]) for b in bad[:6]])
    # timing on a device-resident batch
    dev = [eng.upload(synth_image(4000, 3000, idx=i)) for i in range(16)]
    outs = [eng.alloc_device(2048, 1536, 3) for _ in dev]
    for _ in range(3):
        eng.preprocess_batch(dev, device_outputs=outs)
        print("timing", eng.timing(), flush=True)
    ref = oracle.preprocess(synth_image(4000, 3000, idx=5), 1)
    print("device-resident image 5 equal:", bool(np.array_equal(eng.download(outs[5]), ref)))
