#!/usr/bin/env python
"""Single-image requests from concurrent clients (the reference's calling pattern): synchronous calls that
serialise on the context vs irp_submit / irp_wait, where a dispatcher thread batches whatever is queued.
12 MP RGB images in pinned host memory, classify + preprocess, results back on the host."""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irp_b200
from irp_b200 import _ffi
from irp_b200.synth import synth_batch

W, H, PER = 4000, 3000, 12
imgs = synth_batch(W, H, 4, distinct=4)
with irp_b200.Engine(0) as eng:
    ow, oh = eng.preprocess_dims(W, H)
    lib, ctx = eng._lib, eng._ctx
    for nthreads in (1, 4, 16, 32):
        # per-thread pinned input / output buffers and descriptors, built once
        slots = []
        for t in range(nthreads):
            p = eng.pinned_empty(imgs[0].shape); p[...] = imgs[t % 4]
            o = eng.pinned_empty((oh, ow, 3))
            d = (_ffi.ImageDesc * 1)(_ffi.ImageDesc(p.ctypes.data, W * 3, W, H, 3, 1, 1, 0))
            od = (_ffi.OutDesc * 1)(_ffi.OutDesc(o.ctypes.data, ow * 3, o.nbytes, 0, 0, 0, 0))
            slots.append((p, o, d, od, _ffi.Result()))
        for mode in ("sync", "async"):
            def client(t):
                p, o, d, od, res = slots[t]
                for _ in range(PER):
                    if mode == "sync":
                        rc = lib.irp_analyze_batch(ctx, d, 1, C.byref(res), od)
                    else:
                        tk = C.c_void_p()
                        rc = lib.irp_submit(ctx, d, C.byref(res), od, C.byref(tk)) or lib.irp_wait(ctx, tk, None, 0)
                    assert rc == 0, rc
            for warm in (True, False):
                ts = [threading.Thread(target=client, args=(t,)) for t in range(nthreads)]
                t0 = time.perf_counter()
                for th in ts: th.start()
                for th in ts: th.join()
                dt = time.perf_counter() - t0
            n = nthreads * PER
            print(f"{nthreads:3d} clients {mode:5s}: {n / dt:7.1f} images/s  {n * W * H / dt / 1e9:6.2f} GPix/s  {dt / PER * 1e3:7.2f} ms per request round")

    # the same clients holding JPEG FILES (q90 4:2:0, ~3.5 MB): nothing is decoded on the host
    import io
    from PIL import Image
    blobs = []
    for im in imgs:
        b = io.BytesIO(); Image.fromarray(im).save(b, "JPEG", quality=90, subsampling=2); blobs.append(np.frombuffer(b.getvalue(), np.uint8))
    for nthreads in (1, 16, 64):
        slots = []
        for t in range(nthreads):
            o = eng.pinned_empty((oh, ow, 3))
            k = blobs[t % 4]
            jd = _ffi.JpegDesc(k.ctypes.data, k.size, 1, 0)
            od = (_ffi.OutDesc * 1)(_ffi.OutDesc(o.ctypes.data, ow * 3, o.nbytes, 0, 0, 0, 0))
            slots.append((o, jd, od, _ffi.Result()))
        def jclient(t):
            o, jd, od, res = slots[t]
            for _ in range(PER):
                tk = C.c_void_p()
                rc = lib.irp_submit_jpeg(ctx, C.byref(jd), C.byref(res), od, C.byref(tk)) or lib.irp_wait(ctx, tk, None, 0)
                assert rc == 0, rc
        for warm in (True, False):
            ts = [threading.Thread(target=jclient, args=(t,)) for t in range(nthreads)]
            t0 = time.perf_counter()
            for th in ts: th.start()
            for th in ts: th.join()
            dt = time.perf_counter() - t0
        n = nthreads * PER
        print(f"{nthreads:3d} clients JPEG files: {n / dt:7.1f} images/s  {n * W * H / dt / 1e9:6.2f} GPix/s")
    # files on both sides: one upload per request, scores + the preprocessed FILE back (irp_submit_transcode)
    for nthreads in (1, 16, 64):
        slots = []
        for t in range(nthreads):
            o = eng.pinned_empty((oh, ow, 3))
            k = blobs[t % 4]
            jd = _ffi.JpegDesc(k.ctypes.data, k.size, 1, 0)
            slots.append((o, jd, _ffi.JpegOut(o.ctypes.data, o.nbytes, 0, 0, 0, 0, 0), _ffi.Result()))
        def tclient(t):
            o, jd, enc, res = slots[t]
            for _ in range(PER):
                tk = C.c_void_p()
                rc = lib.irp_submit_transcode(ctx, C.byref(jd), C.byref(res), 85, C.byref(enc), C.byref(tk)) or lib.irp_wait(ctx, tk, None, 0)
                assert rc == 0, rc
        for warm in (True, False):
            ts = [threading.Thread(target=tclient, args=(t,)) for t in range(nthreads)]
            t0 = time.perf_counter()
            for th in ts: th.start()
            for th in ts: th.join()
            dt = time.perf_counter() - t0
        n = nthreads * PER
        print(f"{nthreads:3d} clients files in / files out: {n / dt:7.1f} images/s  {n * W * H / dt / 1e9:6.2f} GPix/s  ({slots[0][2].size / 1e6:.2f} MB per returned file)")
