"""ctypes front end of oracle/jpeg_oracle.c — the CPU restatement of libjpeg-turbo's baseline decode
(JDCT_ISLOW + fancy upsampling), the decoder in front of every sharp pipeline of the hot path.

TEST INFRASTRUCTURE ONLY, and — unlike the rest of oracle/ — PINNED: tests/test_jpeg_oracle.py holds it
bit-exact against Pillow's decoder, which is libjpeg-turbo itself.  It is the oracle of the next row of
the hot-path table (JPEG decode on the device)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjpeg_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "jpeg_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "libjpeg_oracle.so"], check=True, capture_output=True)
        lib = C.CDLL(_SO)
        lib.irp_jpeg_info.argtypes = [C.c_char_p, C.c_size_t] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_int * 6)]
        lib.irp_jpeg_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        lib.irp_jpeg_decode_scaled.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def info(data: bytes) -> dict:
    lib = _load()
    w, h, n, ri = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    samp = (C.c_int * 6)()
    rc = lib.irp_jpeg_info(data, len(data), C.byref(w), C.byref(h), C.byref(n), C.byref(ri), C.byref(samp))
    if rc:
        raise ValueError(f"jpeg oracle: error {rc}")
    return {"width": w.value, "height": h.value, "components": n.value, "restart_interval": ri.value,
            "sampling": [(samp[2 * i], samp[2 * i + 1]) for i in range(n.value)]}


def decode(data: bytes) -> np.ndarray:
    """HxWx3 RGB (or HxW grey) exactly as libjpeg-turbo's default decode gives it."""
    i = info(data)
    shape = (i["height"], i["width"], 3) if i["components"] == 3 else (i["height"], i["width"])
    out = np.empty(shape, np.uint8)
    rc = _load().irp_jpeg_decode(data, len(data), out.ctypes.data, None, -1)
    if rc:
        raise ValueError(f"jpeg oracle: error {rc}")
    return out


def decode_scaled(data: bytes, denom: int) -> np.ndarray:
    """The picture at libjpeg's scale 1 / denom (2, 4, 8): ceil(H / denom) x ceil(W / denom), reduced-size IDCTs
    (jidctred.c) — what libvips' shrink-on-load asks libjpeg-turbo for."""
    i = info(data)
    h, w = -(-i["height"] // denom), -(-i["width"] // denom)
    out = np.empty((h, w, 3) if i["components"] == 3 else (h, w), np.uint8)
    rc = _load().irp_jpeg_decode_scaled(data, len(data), denom, out.ctypes.data)
    if rc:
        raise ValueError(f"jpeg oracle: error {rc}")
    return out


def coefficients(data: bytes, comp: int) -> np.ndarray:
    """Quantised DCT coefficients of one component: [block rows][blocks per row][64], natural order."""
    i = info(data)
    hmax = max(h for h, _ in i["sampling"])
    vmax = max(v for _, v in i["sampling"])
    h, v = i["sampling"][comp]
    if i["components"] == 1:
        h = v = hmax = vmax = 1
    mcux = -(-i["width"] // (8 * hmax))
    mcuy = -(-i["height"] // (8 * vmax))
    out = np.zeros((mcuy * v, mcux * h, 64), np.int16)
    rc = _load().irp_jpeg_decode(data, len(data), None, out.ctypes.data, comp)
    if rc:
        raise ValueError(f"jpeg oracle: error {rc}")
    return out
