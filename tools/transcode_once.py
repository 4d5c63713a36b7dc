#!/usr/bin/env python
"""Two irp_transcode_jpeg_batch calls on the bench workload's shape (N baseline 4:2:0 12 MP files in, scores and
q85 4:4:4 files out) — the command the ncu launch list / captures of the files-in-files-out chain are taken from."""
import io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from irp_b200.synth import synth_image
from PIL import Image

n = int(os.environ.get("N", "64"))
calls = int(os.environ.get("CALLS", "2"))
blobs = []
for s in range(4):
    b = io.BytesIO()
    Image.fromarray(synth_image(4000, 3000, idx=s)).save(b, "JPEG", quality=90, subsampling=2)
    blobs.append(b.getvalue())
jb = [blobs[i % 4] for i in range(n)]
eng = irp_b200.Engine(0)
for _ in range(calls):
    res, files = eng.transcode_jpeg_batch(jb, quality=85)
print(len(files), sum(map(len, files)), res[0]["scores"])
