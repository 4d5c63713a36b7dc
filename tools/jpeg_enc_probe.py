#!/usr/bin/env python
"""Time the device JPEG encoder on the sizes preprocessImage produces, and the files-in / files-out entry point
on the bench workload (64 x 12 MP baseline JPEGs).  IRP_TRACE=1 prints the encoder's stage times."""
import io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B

def main():
    eng = irp_b200.Engine(0)
    n = int(os.environ.get("N", "64"))
    base = [B.synth_image(3000, 4000, seed=s) for s in range(4)] if hasattr(B, "synth_image") else None
    if base is None:
        rng = np.random.default_rng(0)
        yy, xx = np.mgrid[0:3000, 0:4000].astype(np.float32)
        base = []
        for s in range(4):
            a = 128 + 60 * np.sin(xx / (40 + 9 * s)) * np.cos(yy / (31 + 5 * s)) + 30 * np.sin((xx + yy) / 7.0)
            a = a[:, :, None] + rng.normal(0, 5, (3000, 4000, 3))
            base.append(np.clip(a, 0, 255).astype(np.uint8))
    small = eng.preprocess_batch(base)
    dev = [eng.upload(small[i % 4], pitch_align=256) for i in range(n)]
    for it in range(3):
        t0 = time.perf_counter()
        files = eng.encode_jpeg_batch(dev, quality=85)
        dt = time.perf_counter() - t0
        print(f"encode {n} x {small[0].shape}: {dt*1e3:.2f} ms wall (incl. numpy buffer setup), {sum(map(len, files))/1e6:.1f} MB out", flush=True)
    ref = io.BytesIO(); Image.fromarray(small[0]).save(ref, "JPEG", quality=85, subsampling=0)
    print("identical to Pillow:", files[0] == ref.getvalue())
    t0 = time.perf_counter(); 
    for k in range(4):
        b = io.BytesIO(); Image.fromarray(small[k]).save(b, "JPEG", quality=85, subsampling=0)
    print(f"Pillow (libjpeg-turbo, 1 thread) encode: {(time.perf_counter()-t0)/4*1e3:.1f} ms per image")
    blobs = []
    for im in base:
        b = io.BytesIO(); Image.fromarray(im).save(b, "JPEG", quality=90, subsampling=2); blobs.append(b.getvalue())
    jb = [blobs[i % 4] for i in range(n)]
    for it in range(3):
        t0 = time.perf_counter()
        res, files = eng.transcode_jpeg_batch(jb, quality=85)
        dt = time.perf_counter() - t0
        print(f"transcode {n} x 12 MP files: {dt*1e3:.2f} ms wall, {n*12/dt/1e3:.1f} GPix/s, in {sum(map(len, jb))/1e6:.0f} MB out {sum(map(len, files))/1e6:.0f} MB", flush=True)
    for it in range(2):
        t0 = time.perf_counter()
        res, outs = eng.analyze_jpeg_batch(jb)
        dt = time.perf_counter() - t0
        print(f"analyze_jpeg (pixels out) {n} x 12 MP files: {dt*1e3:.2f} ms wall", flush=True)

main()
