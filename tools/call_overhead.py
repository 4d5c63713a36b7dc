#!/usr/bin/env python
"""Where a device-resident irp_analyze_batch call spends its time: wall per call vs the stream's busy span vs the kernels."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irp_b200
from irp_b200 import _ffi
from irp_b200.synth import synth_batch
W, H, B = 4000, 3000, 64
imgs = synth_batch(W, H, B, distinct=4)
with irp_b200.Engine(0) as eng:
    d_in = [eng.upload(im) for im in imgs]
    d_out = [eng.alloc_device(2048, 1536, 3) for _ in imgs]
    descs, _k = eng._descs(d_in, True, None)
    outs = (_ffi.OutDesc * B)(); res = (_ffi.Result * B)()
    for i, d in enumerate(d_out): outs[i] = _ffi.OutDesc(d.ptr, d.pitch, d.nbytes, 0, 0, 0, 1)
    def step(): eng._check(eng._lib.irp_analyze_batch(eng._ctx, descs, B, res, outs))
    for _ in range(3): step()
    n = 20; t0 = time.perf_counter(); tot = cls = pre = 0.0
    for _ in range(n):
        step(); t = eng.timing(); tot += t["total_ms"]; cls += t["classify_ms"]; pre += t["preprocess_ms"]
    wall = (time.perf_counter() - t0) / n * 1e3
    print(f"wall per call {wall:.3f} ms | stream span (first event -> last) {tot / n:.3f} ms | kernels {cls / n:.3f} + {pre / n:.3f} = {(cls + pre) / n:.3f} ms")
