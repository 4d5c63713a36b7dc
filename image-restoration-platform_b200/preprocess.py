"""preprocessImage — mirror of server-node/src/middleware/imagePreprocess.js.

`preprocess_image(req, _res, next)` keeps the Express-middleware shape: it reads `req.file.buffer`,
rewrites the same `req.file` fields (imagePreprocess.js:70-78), records the same operation strings
(:43,54,64-65) and reports failures through `next(Problem)` with the same status codes (:25-34,81-90).
The pixel stages — EXIF auto-orient, lanczos3 fit-inside <= 2048, normalise — run in libirp_b200.so;
container decode and the q85 4:4:4 JPEG entropy coding stay on the host.
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
    eng = getattr(req, "engine", None)
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


def encode_jpeg(px: np.ndarray) -> bytes:
    """.jpeg({quality:85, chromaSubsampling:'4:4:4'}).withMetadata({icc:'sRGB'}) — host side."""
    from PIL import Image, ImageCms

    im = Image.fromarray(px[:, :, 0] if px.shape[2] == 1 else px)
    icc = ImageCms.ImageCmsProfile(ImageCms.createProfile("sRGB")).tobytes()
    out = io.BytesIO()
    im.save(out, format="JPEG", quality=JPEG_QUALITY, subsampling=0, optimize=True, progressive=True, icc_profile=icc)
    return out.getvalue()


def preprocess_image(req, _res, next):  # imagePreprocess.js:24-91
    file = getattr(req, "file", None) if not isinstance(req, dict) else req.get("file")
    buf = _get(file, "buffer") if file is not None else None
    if not buf:
        return next(Problem("https://docs.image-restoration.ai/problem/image-missing", "Image File Required", 400,
                            "An image file must be provided in the request."))
    try:
        operations = []
        px, fmt, orientation = decode_image(buf)
        h, w, c = px.shape
        source_metadata = {"width": w, "height": h, "format": fmt, "channels": c, "orientation": orientation}
        operations.append("auto_orient")
        if needs_resize(w, h):
            d = calculate_resize_dimensions(w, h)
            operations.append(f"resize_{d['width']}x{d['height']}")
        out = _get_engine(req).preprocess_batch([px], orientations=[orientation])[0]
        operations.append(f"compress_jpeg_q{JPEG_QUALITY}")
        operations.append("attach_sRGB_icc")
        processed = encode_jpeg(out)
        _set(file, "originalBuffer", buf)
        _set(file, "originalMetadata", source_metadata)
        _set(file, "buffer", processed)
        _set(file, "processedMetadata", {"width": out.shape[1], "height": out.shape[0], "format": "jpeg", "channels": out.shape[2]})
        _set(file, "processedPixels", out)  # decoded form, so the worker can classify without a re-decode
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
