#!/usr/bin/env python
"""Does running two half-batches on two contexts from two host threads hide the host-side phases (marker scan,
un-stuffing, uploads, read-backs) of the files-in / files-out chain behind the other half's kernels?"""
import io, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from irp_b200.synth import synth_image
from PIL import Image

n = int(os.environ.get("N", "64"))
blobs = []
for s in range(4):
    b = io.BytesIO()
    Image.fromarray(synth_image(4000, 3000, idx=s)).save(b, "JPEG", quality=90, subsampling=2)
    blobs.append(b.getvalue())
jb = [blobs[i % 4] for i in range(n)]
for lanes in (1, 2, 3, 4):
    engs = [irp_b200.Engine(0) for _ in range(lanes)]
    parts = [jb[i::lanes] for i in range(lanes)]
    def work(k):
        engs[k].transcode_jpeg_batch(parts[k], quality=85)
    best = 1e9
    for it in range(5):
        th = [threading.Thread(target=work, args=(k,)) for k in range(lanes)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        if it >= 2: best = min(best, dt)
    print(f"{lanes} lane(s): {best*1e3:.2f} ms per {n} files ({n*12/best/1e3:.1f} GPix/s) [python buffers included]", flush=True)
    for e in engs: e.close()
