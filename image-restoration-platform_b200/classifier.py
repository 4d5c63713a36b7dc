"""ClassifierService — drop-in mirror of server-node/src/services/classifier.js.

Same constructor option (`logger`), same `analyze(imageBuffer)` coroutine returning the
seven scores in the reference's key order (classifier.js:62-70), same static
`getDegradationTypes()` and `createClassifierService(options)` factory (classifier.js:342-349).
What changed is underneath: instead of six sharp pipelines and JS reductions, the decoded
pixels make ONE trip through libirp_b200.so (hand-written sm_100a kernels).  JPEG files — baseline and
progressive — are decoded on the device too (bit-exact with libjpeg-turbo, sharp's decoder); PNG / WebP
containers are decoded on the host first.

The reference's per-analysis fallback constants (classifier.js:123-126 ...) have no analogue:
the fused kernel either produces all moments or the whole call fails, which the reference
already models as a rejected promise (classifier.js:91-95).
"""
from __future__ import annotations

import asyncio
import io
import logging
from typing import Optional, Sequence, Union

import numpy as np

from .engine import Engine, SCORE_KEYS

DEGRADATION_TYPES = {  # classifier.js:17-25
    "blur": "Motion blur or out-of-focus areas",
    "noise": "Grain and digital noise",
    "lowLight": "Underexposed or shadow detail loss",
    "compression": "JPEG artifacts and quality loss",
    "scratch": "Physical damage and blemishes",
    "fade": "Color loss and contrast reduction",
    "colorShift": "White balance and color cast issues",
}

_blockiness_warning_logged = False  # classifier.js:27
_scratch_warning_logged = False  # classifier.js:28


def decode_image(image_buffer: Union[bytes, bytearray, memoryview]):
    """sharp(buf).metadata() + raw decode: returns (pixels HxWxC u8, format, exif_orientation).

    No EXIF rotation is applied — the classifier pipelines never call .rotate()."""
    from PIL import Image

    im = Image.open(io.BytesIO(bytes(image_buffer)))
    fmt = (im.format or "").lower()
    orientation = 1
    try:
        orientation = int(im.getexif().get(0x0112, 1))
    except Exception:
        orientation = 1
    if im.mode in ("L", "RGB", "RGBA"):
        pass
    elif im.mode in ("P", "PA"):
        im = im.convert("RGBA" if "transparency" in im.info or im.mode == "PA" else "RGB")
    elif im.mode in ("1", "I;16", "I", "F"):
        im = im.convert("L")
    else:  # CMYK, YCbCr, LA ...
        im = im.convert("RGBA" if im.mode == "LA" else "RGB")
    px = np.asarray(im, dtype=np.uint8)
    if px.ndim == 2:
        px = px[:, :, None]
    return np.ascontiguousarray(px), fmt, (orientation if 1 <= orientation <= 8 else 1)


class ClassifierService:
    def __init__(self, logger=None, engine: Optional[Engine] = None, device: int = 0):
        self.logger = logger if logger is not None else logging.getLogger("classifier")
        self._engine = engine
        self._device = device

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._device)
        return self._engine

    # classifier.js:40-99
    async def analyze(self, image_buffer) -> dict:
        loop = asyncio.get_running_loop()
        # like napi_async_work: never block the event loop with decode / GPU wait
        return await loop.run_in_executor(None, self.analyze_sync, image_buffer)

    def analyze_sync(self, image_buffer) -> dict:
        try:
            if isinstance(image_buffer, np.ndarray):
                px, fmt = image_buffer, "raw"
            else:
                # a baseline JPEG never touches a host decoder: the file bytes go to the device, which decodes
                # them bit-exactly as libjpeg-turbo (sharp's decoder) would; other containers decode on the host
                if self.engine.jpeg_info(image_buffer) is not None:
                    res, _ = self.engine.analyze_jpeg_batch([bytes(image_buffer)], preprocess=False)
                    return self._finish([res[0]], [self.engine.jpeg_info(image_buffer)[:2]], [True])[0]
                px, fmt, _ = decode_image(image_buffer)
            return self.analyze_pixels([px], [fmt == "jpeg"])[0]
        except Exception as error:  # classifier.js:91-95
            self._log("error", "[classifier] Analysis failed", {"error": str(error)})
            raise

    def analyze_pixels(self, images: Sequence, is_jpeg: Sequence[bool]) -> list:
        """Batched entry the queue worker uses: decoded (or device-resident) images -> score dicts."""
        global _blockiness_warning_logged, _scratch_warning_logged
        if any(is_jpeg) and not _blockiness_warning_logged:  # classifier.js:289-292
            self._log("warning", "[classifier] Blockiness detection is using a simplified heuristic. "
                                 "Enhance with DCT-based analysis for production accuracy.")
            _blockiness_warning_logged = True
        if not _scratch_warning_logged:  # classifier.js:311-314
            self._log("warning", "[classifier] Scratch detection is using a simplified heuristic. "
                                 "Integrate Hough transforms for better accuracy.")
            _scratch_warning_logged = True
        results = self.engine.classify_batch(list(images), is_jpeg=list(is_jpeg))
        sizes = [((im.width, im.height) if hasattr(im, "ptr") else (im.shape[1], im.shape[0])) for im in images]
        return self._finish(results, sizes, is_jpeg)

    def _finish(self, results, sizes, is_jpeg) -> list:
        global _blockiness_warning_logged, _scratch_warning_logged
        if any(is_jpeg) and not _blockiness_warning_logged:  # classifier.js:289-292
            self._log("warning", "[classifier] Blockiness detection is using a simplified heuristic. "
                                 "Enhance with DCT-based analysis for production accuracy.")
            _blockiness_warning_logged = True
        if not _scratch_warning_logged:  # classifier.js:311-314
            self._log("warning", "[classifier] Scratch detection is using a simplified heuristic. "
                                 "Integrate Hough transforms for better accuracy.")
            _scratch_warning_logged = True
        out = []
        for (w, h), r in zip(sizes, results):
            analysis = {k: r["scores"][k] for k in SCORE_KEYS}
            # classifier.js:73-76 — the library returns the top three next to the scores (irp_result.issues)
            self._log("debug", "[classifier] Analysis complete",
                      {"topIssues": [{"type": t["type"], "score": f"{t['confidence']:.2f}"} for t in r["issues"]], "imageSize": f"{w}x{h}"})
            self.last_top_issues = r["issues"]   # what PromptEnhancerService._identifyTopIssues would compute again (promptEnhancer.js:121-137)
            out.append(analysis)
        return out

    def _log(self, level: str, msg: str, meta: Optional[dict] = None) -> None:
        fn = getattr(self.logger, level, None) or getattr(self.logger, "warn" if level == "warning" else level, None)
        if fn is None:
            return
        try:
            fn(msg, meta) if meta is not None and not isinstance(self.logger, logging.Logger) else fn(msg if meta is None else f"{msg} {meta}")
        except TypeError:
            fn(msg)

    @staticmethod
    def getDegradationTypes() -> dict:  # classifier.js:342-344
        return dict(DEGRADATION_TYPES)

    get_degradation_types = getDegradationTypes


def createClassifierService(options: Optional[dict] = None, **kw) -> ClassifierService:  # classifier.js:347-349
    opts = dict(options or {})
    opts.update(kw)
    return ClassifierService(**opts)


create_classifier_service = createClassifierService
