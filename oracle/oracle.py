"""ctypes front end of the CPU oracle (oracle/irp_oracle.c).

TEST INFRASTRUCTURE ONLY — parity unpinned (see the header of irp_oracle.c).
Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libirp_oracle.so")

SCORE_KEYS = ("blur", "noise", "lowLight", "compression", "scratch", "fade", "colorShift")


class Result(C.Structure):
    """Mirror of irp_result (include/irp.h)."""

    _fields_ = [
        ("score", C.c_double * 7),
        ("sum", C.c_uint64 * 4),
        ("sumsq", C.c_uint64 * 4),
        ("e_sum", C.c_uint64 * 2),
        ("e_sumsq", C.c_uint64 * 2),
        ("b_sum", C.c_uint64),
        ("b_sumsq", C.c_uint64),
        ("scratch_v", C.c_uint32),
        ("scratch_h", C.c_uint32),
        ("block_edges", C.c_uint32 * 2),
        ("luma_hist", C.c_uint32 * 256),
        ("status", C.c_int32),
        ("issues", C.c_uint8 * 4),
    ]


class ImageDesc(C.Structure):
    """Mirror of irp_image_desc (include/irp.h)."""

    _fields_ = [
        ("pixels", C.c_void_p),
        ("pitch", C.c_size_t),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("channels", C.c_int32),
        ("is_jpeg", C.c_int32),
        ("exif_orientation", C.c_int32),
        ("on_device", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "irp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_grey_rgb.restype = C.c_uint8
    return _lib


def _img(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    return a, w, h, c


def result_to_dict(r: Result) -> dict:
    return {
        "scores": dict(zip(SCORE_KEYS, list(r.score))),
        "sum": list(r.sum),
        "sumsq": list(r.sumsq),
        "e_sum": list(r.e_sum),
        "e_sumsq": list(r.e_sumsq),
        "b_sum": int(r.b_sum),
        "b_sumsq": int(r.b_sumsq),
        "scratch_v": int(r.scratch_v),
        "scratch_h": int(r.scratch_h),
        "block_edges": list(r.block_edges),
        "luma_hist": list(r.luma_hist),
        "status": int(r.status),
        "issues": [{"type": SCORE_KEYS[int(r.issues[k]) & 15], "confidence": float(r.score[int(r.issues[k]) & 15]),
                    "severity": {1: "low", 2: "medium", 3: "high"}[int(r.issues[k]) >> 4]} for k in range(int(r.issues[3]))],
    }


def classify(a: np.ndarray, is_jpeg: bool = True, luma_mode: int = 0, blur_mode: int = 0) -> dict:
    a, w, h, c = _img(a)
    r = Result()
    rc = lib().orc_classify_m(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), int(is_jpeg), luma_mode, blur_mode, C.byref(r))
    if rc:
        raise ValueError(f"orc_classify rc={rc}")
    return result_to_dict(r)


def grey(a: np.ndarray, luma_mode: int = 0) -> np.ndarray:
    a, w, h, c = _img(a)
    out = np.empty((h, w), np.uint8)
    rc = lib().orc_grey(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), luma_mode, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"orc_grey rc={rc}")
    return out


def grey_all(luma_mode: int = 0) -> np.ndarray:
    """Grey of every (r,g,b) triple, index (r<<16)|(g<<8)|b — for the exhaustive table check."""
    rr, gg, bb = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.stack([rr.ravel(), gg.ravel(), bb.ravel()], axis=1).reshape(4096, 4096, 3)
    return grey(img, luma_mode).ravel()


def stencil(g: np.ndarray, which: int) -> np.ndarray:
    g = np.ascontiguousarray(g, np.uint8)
    h, w = g.shape
    out = np.empty_like(g)
    lib().orc_stencil(g.ctypes.data_as(C.c_void_p), w, h, which, out.ctypes.data_as(C.c_void_p))
    return out


def blur1(a: np.ndarray, blur_mode: int = 0) -> np.ndarray:
    a, w, h, c = _img(a)
    out = np.empty((h, w, c), np.uint8)
    lib().orc_blur1_m(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), blur_mode, out.ctypes.data_as(C.c_void_p))
    return out


def box_factor(in_size: int, out_size: int) -> int:
    return lib().orc_box_factor(in_size, out_size)


def box_shrink(a: np.ndarray, kh: int, kv: int) -> np.ndarray:
    a, w, h, c = _img(a)
    out = np.empty(((h + kv - 1) // kv, (w + kh - 1) // kh, c), np.uint8)
    rc = lib().orc_box_shrink(a.ctypes.data_as(C.c_void_p), w, h, c, kh, kv, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"orc_box_shrink rc={rc}")
    return out


def orient(a: np.ndarray, orientation: int) -> np.ndarray:
    a, w, h, c = _img(a)
    ow, oh = C.c_int(), C.c_int()
    lib().orc_orient_dims(w, h, orientation, C.byref(ow), C.byref(oh))
    out = np.empty((oh.value, ow.value, c), np.uint8)
    lib().orc_orient(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), orientation, out.ctypes.data_as(C.c_void_p))
    return out


def preprocess_dims(w: int, h: int, orientation: int = 1):
    ow, oh, sh = C.c_int(), C.c_int(), C.c_double()
    lib().orc_preprocess_dims(w, h, orientation, C.byref(ow), C.byref(oh), C.byref(sh))
    return ow.value, oh.value, sh.value


def fusion_dims(w: int, h: int, orientation: int = 1):
    ow, oh, ox, oy, sh = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
    lib().orc_fusion_dims(w, h, orientation, C.byref(ow), C.byref(oh), C.byref(ox), C.byref(oy), C.byref(sh))
    return ow.value, oh.value, ox.value, oy.value, sh.value


def reduce_plan(in_size: int, out_size: int, shrink: float, coef_mode: int = 0, reduce_mode: int = 0):
    n = C.c_int()
    start = np.empty(out_size, np.int32)
    phase = np.empty(out_size, np.int32)
    coefs = np.zeros((65, 25), np.int16)
    rc = lib().orc_reduce_plan_m(in_size, out_size, C.c_double(shrink), coef_mode, reduce_mode, C.byref(n), start.ctypes.data_as(C.c_void_p),
                                 phase.ctypes.data_as(C.c_void_p), coefs.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"orc_reduce_plan rc={rc}")
    return n.value, start, phase, coefs


def preprocess(a: np.ndarray, orientation: int = 1, coef_mode: int = 0, reduce_mode: int = 0) -> np.ndarray:
    a, w, h, c = _img(a)
    ow, oh, _ = preprocess_dims(w, h, orientation)
    oc = 3 if c == 4 else c
    out = np.empty((oh, ow, oc), np.uint8)
    rw, rh, rc_ = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_preprocess_m(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), orientation, coef_mode, reduce_mode,
                                out.ctypes.data_as(C.c_void_p), C.byref(rw), C.byref(rh), C.byref(rc_))
    if rc:
        raise ValueError(f"orc_preprocess rc={rc}")
    assert (rw.value, rh.value, rc_.value) == (ow, oh, oc)
    return out


def fusion_canvas(a: np.ndarray, orientation: int = 1, coef_mode: int = 0, reduce_mode: int = 0) -> np.ndarray:
    a, w, h, c = _img(a)
    out = np.empty((2048, 2048, 3), np.uint8)
    rc = lib().orc_fusion_canvas_m(a.ctypes.data_as(C.c_void_p), w, h, c, C.c_size_t(w * c), orientation, coef_mode, reduce_mode,
                                   out.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"orc_fusion_canvas rc={rc}")
    return out


def analyze_batch(images, orientations=None, threads: int = 1, with_preprocess: bool = True, luma_mode: int = 0, coef_mode: int = 0):
    """Classify (+ preprocess) a list of HWC u8 arrays on `threads` host threads. Returns (results, outs)."""
    n = len(images)
    arrs = [_img(a) for a in images]
    descs = (ImageDesc * n)()
    results = (Result * n)()
    outs = []
    outp = (C.c_void_p * n)()
    for i, (a, w, h, c) in enumerate(arrs):
        o = 1 if orientations is None else int(orientations[i])
        descs[i] = ImageDesc(a.ctypes.data, w * c, w, h, c, 1, o, 0)
        if with_preprocess:
            ow, oh, _ = preprocess_dims(w, h, o)
            out = np.empty((oh, ow, 3 if c == 4 else c), np.uint8)
            outs.append(out)
            outp[i] = out.ctypes.data
    rc = lib().orc_analyze_batch(descs, n, luma_mode, coef_mode, results, outp if with_preprocess else None, threads)
    if rc:
        raise ValueError(f"orc_analyze_batch rc={rc}")
    return [result_to_dict(r) for r in results], outs
