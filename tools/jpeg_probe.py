#!/usr/bin/env python
"""Throughput of the compressed-input path: 64 x 12 MP baseline JPEGs (q90, 4:2:0) in host memory ->
irp_analyze_jpeg_batch (H2D of the compressed bytes, device decode, classify, preprocess, results and resized
pixels back on the host), next to the decode alone and to Pillow (libjpeg-turbo) on the host cores."""
import ctypes as C, io, os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
import irp_b200
from irp_b200 import _ffi
from irp_b200.synth import synth_batch

W, H, B = 4000, 3000, 64
imgs = synth_batch(W, H, 8, distinct=8)
blobs = []
for im in imgs:
    b = io.BytesIO(); Image.fromarray(im).save(b, "JPEG", quality=90, subsampling=2); blobs.append(b.getvalue())
blobs = [blobs[i % 8] for i in range(B)]
print(f"{B} JPEGs, {sum(map(len, blobs)) / B / 1e6:.2f} MB each on average")
keep = [np.frombuffer(b, np.uint8) for b in blobs]
with irp_b200.Engine(0) as eng:
    lib, ctx = eng._lib, eng._ctx
    descs = (_ffi.JpegDesc * B)(*[_ffi.JpegDesc(k.ctypes.data, k.size, 1, 0) for k in keep])
    ow, oh = eng.preprocess_dims(W, H)
    h_out = [eng.pinned_empty((oh, ow, 3)) for _ in range(B)]
    outs = (_ffi.OutDesc * B)(*[_ffi.OutDesc(o.ctypes.data, ow * 3, o.nbytes, 0, 0, 0, 0) for o in h_out])
    res = (_ffi.Result * B)()
    def step():
        rc = lib.irp_analyze_jpeg_batch(ctx, descs, B, res, outs)
        assert rc == 0, eng._lib.irp_last_error(ctx)
    for _ in range(2): step()
    t0 = time.perf_counter()
    for _ in range(5): step()
    dt = (time.perf_counter() - t0) / 5
    tm = eng.timing()
    print(f"analyze from JPEG bytes: {dt * 1e3:.1f} ms per {B} images  {B * W * H / dt / 1e9:.2f} GPix/s  (classify {tm['classify_ms']:.2f} ms, preprocess {tm['preprocess_ms']:.2f} ms, launches {tm['kernel_launches']})")
    # the same pixels, decoded by Pillow and uploaded: is the classify / preprocess time a matter of content?
    d_px = [eng.upload(np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(b))))) for b in blobs[:8]]
    d_px = [d_px[i % 8] for i in range(B)]
    d_o2 = [eng.alloc_device(ow, oh, 3) for _ in range(B)]
    for _ in range(3): eng.analyze_batch(d_px, device_outputs=d_o2, raw=True)
    tm = eng.timing()
    print(f"analyze of the Pillow-decoded pixels (device resident): classify {tm['classify_ms']:.2f} ms, preprocess {tm['preprocess_ms']:.2f} ms")
    # decode only, to device buffers
    d_out = [eng.alloc_device(W, H, 3) for _ in range(B)]
    douts = (_ffi.OutDesc * B)(*[_ffi.OutDesc(d.ptr, d.pitch, d.nbytes, 0, 0, 0, 1) for d in d_out])
    for _ in range(2): assert lib.irp_decode_jpeg_batch(ctx, descs, B, douts) == 0
    t0 = time.perf_counter()
    for _ in range(5): assert lib.irp_decode_jpeg_batch(ctx, descs, B, douts) == 0
    dd = (time.perf_counter() - t0) / 5
    print(f"decode only (to device): {dd * 1e3:.1f} ms per {B} images  {B * W * H / dd / 1e9:.2f} GPix/s")
T = len(os.sched_getaffinity(0))
def pil_decode(b): return np.asarray(Image.open(io.BytesIO(b))).shape
with ThreadPoolExecutor(T) as ex:
    t0 = time.perf_counter(); list(ex.map(pil_decode, blobs[:2 * T])); dp = time.perf_counter() - t0
print(f"Pillow / libjpeg-turbo decode on {T} host threads: {2 * T * W * H / dp / 1e9:.2f} GPix/s")
