#!/usr/bin/env python
"""8 x 12 MP JPEG decode, for an ncu launch list."""
import ctypes as C, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
import irp_b200
from irp_b200 import _ffi
from irp_b200.synth import synth_batch
W, H, B = 4000, 3000, 8
blobs = []
for im in synth_batch(W, H, B, distinct=B):
    b = io.BytesIO(); Image.fromarray(im).save(b, "JPEG", quality=90, subsampling=2); blobs.append(b.getvalue())
keep = [np.frombuffer(b, np.uint8) for b in blobs]
with irp_b200.Engine(0) as eng:
    descs = (_ffi.JpegDesc * B)(*[_ffi.JpegDesc(k.ctypes.data, k.size, 1, 0) for k in keep])
    d_out = [eng.alloc_device(W, H, 3) for _ in range(B)]
    douts = (_ffi.OutDesc * B)(*[_ffi.OutDesc(d.ptr, d.pitch, d.nbytes, 0, 0, 0, 1) for d in d_out])
    for _ in range(3):
        assert eng._lib.irp_decode_jpeg_batch(eng._ctx, descs, B, douts) == 0
print("ok")
