#!/usr/bin/env python
"""What the host link of this box can do, next to the library's end-to-end step: pinned H2D alone,
D2H alone, both directions at once, and irp_analyze_batch on host buffers for several chunk sizes."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import irp_b200
from irp_b200 import _ffi
from irp_b200.synth import synth_batch

W, H, B = 4000, 3000, 64
nin, nout = B * W * H * 3, B * 2048 * 1536 * 3
hin = torch.empty(nin, dtype=torch.uint8).pin_memory()
hout = torch.empty(nout, dtype=torch.uint8).pin_memory()
din = torch.empty(nin, dtype=torch.uint8, device="cuda")
dout = torch.empty(nout, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


def both():
    h2d()
    d2h()


t = timed(h2d); print(f"H2D alone  {nin / t / 1e9:6.1f} GB/s  ({t * 1e3:.1f} ms for {nin / 1e9:.2f} GB)")
t = timed(d2h); print(f"D2H alone  {nout / t / 1e9:6.1f} GB/s  ({t * 1e3:.1f} ms for {nout / 1e9:.2f} GB)")
t = timed(both); print(f"both       {(nin + nout) / t / 1e9:6.1f} GB/s aggregate ({t * 1e3:.1f} ms) -> floor of an end-to-end step")
del hin, hout, din, dout

imgs = synth_batch(W, H, B, distinct=4)
for chunk_mb in (0, 64, 128, 512):
    lib = _ffi.load()
    opts = _ffi.Opts(C.sizeof(_ffi.Opts), 0, 0, 0, chunk_mb << 20)
    eng = irp_b200.Engine.__new__(irp_b200.Engine)
    eng._lib = lib
    eng._ctx = lib.irp_create(0, C.byref(opts))
    eng.device, eng._keep = 0, []
    ow, oh = eng.preprocess_dims(W, H)
    h_in = []
    for im in imgs:
        p = eng.pinned_empty(im.shape); p[...] = im; h_in.append(p)
    h_out = [eng.pinned_empty((oh, ow, 3)) for _ in imgs]
    descs, keep = eng._descs(h_in, True, None)
    outs = (_ffi.OutDesc * B)(); res = (_ffi.Result * B)()
    def step():
        for i, o in enumerate(h_out):
            outs[i] = _ffi.OutDesc(o.ctypes.data, ow * 3, o.nbytes, 0, 0, 0, 0)
        eng._check(lib.irp_analyze_batch(eng._ctx, descs, B, res, outs))
    step(); step()
    t0 = time.perf_counter()
    for _ in range(4):
        step()
    dt = (time.perf_counter() - t0) / 4
    tm = eng.timing()
    print(f"chunk {chunk_mb or 256:4d} MiB: {dt * 1e3:6.1f} ms/step  {B * W * H / dt / 1e9:5.2f} GPix/s  chunks {tm.get('chunks')}")
    eng.close()
