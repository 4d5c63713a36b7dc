#!/usr/bin/env python
"""Generate image-restoration-platform_b200/csrc/grey_tables.inc.

libvips' colourspace(B_W) on sRGB u8 is, per pixel,
three float LUT reads, a double-promoted weighted sum rounded to float, a float
multiply, and a float lerp through an integer gamma LUT followed by VIPS_RINT.
On the GPU that chain costs ~5 shared-memory reads and ~14 mixed FP64/FP32/
conversion instructions.  The result g is a monotone step function of the
weighted sum Y, so it can be produced with integer arithmetic only:

    I   = lutR[r] + lutG[g] + lutB[b]            (u32, Y scaled by ~2^32)
    e   = inv[I >> 20]                           (4096 bins, <= 1 step per bin)
    out = (e + (I & 0xFFFFF)) >> 24              (carry out of the low 24 bits = "I >= step")

This script restates the float pipeline with numpy float32/float64 scalars,
evaluates it on ALL 2^24 (r,g,b) triples, searches a scale for which the
integer pipeline reproduces it on every triple, and bakes the tables.
tests/test_grey_tables.py re-checks all 2^24 triples against the C oracle.

Usage: python tools/gen_grey_tables.py   (writes the .inc in place)
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "image-restoration-platform_b200", "csrc", "grey_tables.inc")

WEIGHTS = {0: (0.2, 0.7, 0.1), 1: (0.2126, 0.7152, 0.0722)}  # IRP_LUMA_VIPS_USUAL, IRP_LUMA_CIE
NBINS = 4096
BIN_SHIFT = 20


def vips_rint(x: float) -> int:
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def float_tables():
    """calcul_tables() of libvips colour/LabQ2sRGB.c for range 256."""
    y2v = np.zeros(257, np.int32)
    v2y = np.zeros(256, np.float32)
    for i in range(256):
        f = np.float32(i) / np.float32(255)
        if float(f) <= 0.0031308:
            v = np.float32(12.92 * float(f))
        else:
            v = np.float32((1.0 + 0.055) * math.pow(float(f), 1.0 / 2.4) - 0.055)
        y2v[i] = vips_rint(float(np.float32(255) * v))
    y2v[256] = y2v[255]
    for i in range(256):
        f = np.float32(i) / np.float32(255)
        if float(f) <= 0.04045:
            v2y[i] = np.float32(float(f) / 12.92)
        else:
            v2y[i] = np.float32(math.pow((float(f) + 0.055) / (1 + 0.055), 2.4))
    return v2y, y2v


def reference_grey_all(v2y, y2v, w):
    """g for all triples, index (r<<16)|(g<<8)|b, following vips_col_scRGB2BW_8."""
    lin = v2y.astype(np.float64)
    a_r = (w[0] * lin)[:, None, None]
    a_g = (w[1] * lin)[None, :, None]
    a_b = (w[2] * lin)[None, None, :]
    y = ((a_r + a_g) + a_b).astype(np.float32).ravel()  # double sum, stored to float
    yf = y * np.float32(255)
    np.clip(yf, np.float32(0), np.float32(255), out=yf)
    yi = yf.astype(np.int32)
    lo = y2v[yi].astype(np.float32)
    d = (y2v[yi + 1] - y2v[yi]).astype(np.float32)
    v = lo + d * (yf - yi.astype(np.float32))
    g = (v.astype(np.float64) + 0.5).astype(np.int32)  # VIPS_RINT for v >= 0
    return g.astype(np.uint8)


def int_luts(v2y, w, scale):
    lin = v2y.astype(np.float64)
    return [np.rint(w[k] * lin * scale).astype(np.uint64).astype(np.uint32) for k in range(3)]


def int_sum_all(luts):
    s = luts[0].astype(np.uint64)[:, None, None] + luts[1].astype(np.uint64)[None, :, None] + luts[2].astype(np.uint64)[None, None, :]
    assert int(s.max()) < 2 ** 32
    return s.astype(np.uint32).ravel()


def solve(mode):
    v2y, y2v = float_tables()
    w = WEIGHTS[mode]
    g = reference_grey_all(v2y, y2v, w)
    order = np.argsort(g, kind="stable")
    gs = g[order]
    present = np.unique(gs)
    starts = np.searchsorted(gs, present, side="left")
    assert len(present) == 256, "some grey level is never produced"
    for j in range(4000):
        scale = float(2 ** 32 - 64 - 4099 * j)
        luts = int_luts(v2y, w, scale)
        i_all = int_sum_all(luts)[order]
        mins = np.minimum.reduceat(i_all, starts)
        maxs = np.maximum.reduceat(i_all, starts)
        if not np.all(maxs[:-1] < mins[1:]):
            continue
        thr = mins.astype(np.int64)  # thr[k] = smallest I with g >= k (k >= 1)
        bins = thr[1:] >> BIN_SHIFT
        if len(np.unique(bins)) != len(bins):
            continue  # two steps in one bin
        inv = np.zeros(NBINS, np.uint32)
        for b in range(NBINS):
            lo_edge = b << BIN_SHIFT
            base = int(np.searchsorted(thr[1:], lo_edge, side="right"))  # steps at or below the bin start
            inside = thr[1:][(thr[1:] > lo_edge) & (thr[1:] < lo_edge + (1 << BIN_SHIFT))]
            e = base << 24
            if len(inside):
                e |= (1 << 24) - int(inside[0] - lo_edge)
            inv[b] = e
        # verify the exact device formula on every triple
        i_nat = int_sum_all(luts)
        e = inv[i_nat >> BIN_SHIFT].astype(np.uint64)
        got = ((e + (i_nat & 0xFFFFF).astype(np.uint64)) >> 24).astype(np.uint8)
        if np.array_equal(got, g):
            return scale, luts, inv, v2y, y2v
    raise RuntimeError("no scale found")


def emit(f, name, arr, per=8, fmt="0x%08xu"):
    f.write("static const uint32_t %s[%d] = {\n" % (name, len(arr)))
    for i in range(0, len(arr), per):
        f.write("  " + ", ".join(fmt % int(v) for v in arr[i : i + per]) + ",\n")
    f.write("};\n")


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_grey_tables.py — do not edit.\n")
        f.write("// Integer restatement of libvips colourspace(B_W) for sRGB u8, verified on all 2^24 triples.\n")
        f.write("// g = (inv[I >> 20] + (I & 0xFFFFF)) >> 24,  I = lut_r[r] + lut_g[g] + lut_b[b]\n")
        for mode in (0, 1):
            scale, luts, inv, v2y, y2v = solve(mode)
            print("mode", mode, "scale", scale, file=sys.stderr)
            f.write("// luma mode %d: weights %s, scale %.1f\n" % (mode, WEIGHTS[mode], scale))
            for k, ch in enumerate("rgb"):
                emit(f, "kGreyLut%d_%s" % (mode, ch), luts[k])
            emit(f, "kGreyInv%d" % mode, inv)
    print("wrote", OUT, file=sys.stderr)


if __name__ == "__main__":
    main()
