"""preprocessImage — mirror of server-node/src/middleware/imagePreprocess.js.

`preprocess_image(req, _res, next)` keeps the Express-middleware shape: it reads `req.file.buffer`,
rewrites the same `req.file` fields (imagePreprocess.js:70-78), records the same operation strings
(:43,54,64-65) and reports failures through `next(Problem)` with the same status codes (:25-34,81-90).
Everything between the upload and the returned file runs in libirp_b200.so: JPEG uploads (baseline and progressive) are
decoded on the device (other containers by Pillow on the host), then EXIF auto-orient, lanczos3 fit-inside <= 2048, normalise,
and the q85 4:4:4 JPEG encode with optimised Huffman tables and the sRGB profile attached (SURVEY.md §8f rows 1-2;
the file is libjpeg-turbo's optimised sequential file of those pixels, not mozjpeg's trellis / progressive one).
"""
from __future__ import annotations

import io
from types import SimpleNamespace
from typing import Optional

import numpy as np

from .classifier import decode_image
from .engine import Engine

MAX_DIMENSION = 2048  # imagePreprocess.js:4
JPEG_QUALITY = 85  # imagePreprocess.js:5


class Problem(Exception):
    """RFC 7807 problem, server-node/src/utils/problem.js:5-22."""

    def __init__(self, type: str, title: str, status: int, detail: str):
        super().__init__(detail)
        self.type, self.title, self.status, self.detail = type, title, status, detail


def needs_resize(width, height) -> bool:  # imagePreprocess.js:7-10
    if not width or not height:
        return False
    return width > MAX_DIMENSION or height > MAX_DIMENSION


def calculate_resize_dimensions(width, height) -> dict:  # imagePreprocess.js:12-22
    if not width or not height:
        return {}
    scale = MAX_DIMENSION / max(width, height)
    if scale >= 1:
        return {"width": width, "height": height}
    import math

    return {"width": math.floor(width * scale + 0.5), "height": math.floor(height * scale + 0.5)}


_engine: Optional[Engine] = None


def _get_engine(req) -> Engine:
    global _engine
    eng = getattr(req, "engine", None) if req is not None else None
    if eng is not None:
        return eng
    if _engine is None:
        _engine = Engine(0)
    return _engine


def _get(file, key, default=None):
    return file.get(key, default) if isinstance(file, dict) else getattr(file, key, default)


def _set(file, key, value):
    if isinstance(file, dict):
        file[key] = value
    else:
        setattr(file, key, value)


def encode_jpeg(px: np.ndarray, engine: Optional[Engine] = None) -> bytes:
    """.jpeg({quality:85, chromaSubsampling:'4:4:4'}).withMetadata({icc:'sRGB'}) — on the device.  The profile is named
    per call (Engine.SRGB, the library's generated sRGB profile): nothing on the shared engine is toggled."""
    eng = engine if engine is not None else _get_engine(None)
    return eng.encode_jpeg_batch([px[:, :, 0] if px.shape[2] == 1 else px], quality=JPEG_QUALITY, optimize=True, icc=Engine.SRGB)[0]


def _jpeg_orientation(buf) -> int:
    """EXIF orientation from the header only (no pixel decode)."""
    try:
        from PIL import Image

        o = int(Image.open(io.BytesIO(bytes(buf))).getexif().get(0x0112, 1))
        return o if 1 <= o <= 8 else 1
    except Exception:
        return 1


def preprocess_image(req, _res, next):  # imagePreprocess.js:24-91
    file = getattr(req, "file", None) if not isinstance(req, dict) else req.get("file")
    buf = _get(file, "buffer") if file is not None else None
    if not buf:
        return next(Problem("https://docs.image-restoration.ai/problem/image-missing", "Image File Required", 400,
                            "An image file must be provided in the request."))
    try:
        operations = []
        eng = _get_engine(req)
        info = eng.jpeg_info(buf)   # (w, h, channels) when the device decoder takes the file, else None
        if info is not None:
            w, h, c = info
            fmt, orientation, px = "jpeg", _jpeg_orientation(buf), None
        else:
            px, fmt, orientation = decode_image(buf)
            h, w, c = px.shape
        source_metadata = {"width": w, "height": h, "format": fmt, "channels": c, "orientation": orientation}
        operations.append("auto_orient")
        if needs_resize(w, h):
            d = calculate_resize_dimensions(w, h)
            operations.append(f"resize_{d['width']}x{d['height']}")
        operations.append(f"compress_jpeg_q{JPEG_QUALITY}")
        operations.append("attach_sRGB_icc")
        q = JPEG_QUALITY | 0x100   # IRP_JPEG_OPTIMIZE: sharp's mozjpeg preset implies optimised Huffman tables
        if px is None:   # a baseline JPEG upload: ONE call, file in -> file out, no pixel crosses PCIe
            processed = eng.transcode_jpeg_batch([bytes(buf)], orientations=[orientation], quality=q, classify=False, icc=Engine.SRGB)[1][0]
        else:            # other containers were decoded on the host: pixels in -> file out, still one call
            processed = eng.analyze_encode_batch([px[:, :, 0] if c == 1 else px], is_jpeg=False, orientations=[orientation], quality=q,
                                                 classify=False, icc=Engine.SRGB)[1][0]
        ow, oh = eng.preprocess_dims(w, h, orientation)
        _set(file, "originalBuffer", buf)
        _set(file, "originalMetadata", source_metadata)
        _set(file, "buffer", processed)
        _set(file, "processedMetadata", {"width": ow, "height": oh, "format": "jpeg", "channels": 1 if c == 1 else 3})
        _set(file, "mimetype", "image/jpeg")
        _set(file, "detectedMime", "image/jpeg")
        _set(file, "detectedExt", "jpg")
        _set(file, "size", len(processed))
        _set(file, "preprocessOperations", operations)
        return next()
    except Exception as error:
        return next(Problem("https://docs.image-restoration.ai/problem/preprocess-failed", "Image Preprocessing Failed", 422,
                            str(error) or "Unable to preprocess the uploaded image."))


preprocessImage = preprocess_image


def make_request(buffer: Optional[bytes] = None, engine: Optional[Engine] = None):
    """Tiny stand-in for the Express `req` the reference tests build (tests/middleware.test.js:14-30)."""
    req = SimpleNamespace(file=SimpleNamespace(buffer=buffer) if buffer is not None else None)
    if engine is not None:
        req.engine = engine
    return req
