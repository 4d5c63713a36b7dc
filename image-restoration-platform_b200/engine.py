"""Engine — one GPU context of libirp_b200.so, numpy in / numpy out.

Thin marshalling over the C ABI (include/irp.h): it builds the descriptor
arrays, owns nothing but the context handle and optional device buffers, and
raises IrpError with irp_last_error() when a call returns non-zero — the same
"whole-call failure = rejected Promise" contract the reference classifier has
(server-node/src/services/classifier.js:91-95).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _ffi

SCORE_KEYS = ("blur", "noise", "lowLight", "compression", "scratch", "fade", "colorShift")  # classifier.js:62-70


class IrpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"irp error {code}: {message}")
        self.code = code
        self.message = message


@dataclass
class DeviceImage:
    """An image (or output buffer) resident in this context's HBM."""

    ptr: int
    width: int
    height: int
    channels: int
    pitch: int
    nbytes: int
    is_jpeg: bool = True
    orientation: int = 1


ImageLike = Union[np.ndarray, DeviceImage]


def _as_hwc(a: np.ndarray) -> np.ndarray:
    if a.dtype != np.uint8:
        raise TypeError("pixels must be uint8")
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3:
        raise ValueError("pixels must be HxW or HxWxC")
    if a.strides[2] != 1 or a.strides[1] != a.shape[2]:
        a = np.ascontiguousarray(a)
    return a


def result_to_dict(r: _ffi.Result) -> dict:
    return {
        "scores": dict(zip(SCORE_KEYS, [float(x) for x in r.score])),
        "sum": list(r.sum), "sumsq": list(r.sumsq), "e_sum": list(r.e_sum), "e_sumsq": list(r.e_sumsq),
        "b_sum": int(r.b_sum), "b_sumsq": int(r.b_sumsq), "scratch_v": int(r.scratch_v), "scratch_h": int(r.scratch_h),
        "block_edges": list(r.block_edges), "luma_hist": list(r.luma_hist), "status": int(r.status),
        "issues": issues_of(r),
    }


SEVERITY = {1: "low", 2: "medium", 3: "high"}


def issues_of(r) -> list:
    """irp_result.issues -> what PromptEnhancerService._identifyTopIssues returns (promptEnhancer.js:121-137):
    [{type, confidence, severity}], highest confidence first, at most three."""
    out = []
    for k in range(int(r.issues[3])):
        v = int(r.issues[k])
        out.append({"type": SCORE_KEYS[v & 15], "confidence": float(r.score[v & 15]), "severity": SEVERITY[v >> 4]})
    return out


class Engine:
    def __init__(self, device: int = 0, luma_mode: int = 0, coef_mode: int = 0, blur_mode: int = 0, reduce_mode: int = 0):
        self._lib = _ffi.load()
        opts = _ffi.Opts(C.sizeof(_ffi.Opts), luma_mode, coef_mode, 0, 0, blur_mode, reduce_mode)
        self._ctx = self._lib.irp_create(device, C.byref(opts))
        if not self._ctx:
            raise IrpError(_ffi.IRP_ERR_NO_DEVICE, (self._lib.irp_last_error(None) or b"irp_create failed").decode())
        self.device = device
        self._keep: list = []

    # -- lifetime ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.irp_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise IrpError(rc, (self._lib.irp_last_error(self._ctx) or b"").decode())

    # -- plumbing -----------------------------------------------------------
    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Launch on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._lib.irp_set_stream(self._ctx, cuda_stream or None))

    def timing(self) -> dict:
        t = _ffi.Timing()
        self._check(self._lib.irp_get_timing(self._ctx, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _ffi.Timing._fields_}

    def synchronize(self) -> None:
        self._check(self._lib.irp_synchronize(self._ctx))

    def upload(self, a: np.ndarray, is_jpeg: bool = True, orientation: int = 1, pitch_align: int = 16) -> DeviceImage:
        """Stage an image in HBM (rows padded to `pitch_align` bytes) for device-resident calls."""
        a = _as_hwc(a)
        h, w, c = a.shape
        pitch = (w * c + pitch_align - 1) // pitch_align * pitch_align
        if pitch != w * c:
            buf = np.zeros((h, pitch), np.uint8)
            buf[:, : w * c] = a.reshape(h, w * c)
        else:
            buf = np.ascontiguousarray(a).reshape(h, pitch)
        nbytes = pitch * h
        ptr = self._lib.irp_dev_alloc(self._ctx, nbytes)
        if not ptr:
            self._check(_ffi.IRP_ERR_NOMEM)
        self._check(self._lib.irp_memcpy_h2d(self._ctx, ptr, buf.ctypes.data, nbytes))
        return DeviceImage(ptr, w, h, c, pitch, nbytes, is_jpeg, orientation)

    def alloc_device(self, width: int, height: int, channels: int) -> DeviceImage:
        nbytes = width * height * channels
        ptr = self._lib.irp_dev_alloc(self._ctx, nbytes)
        if not ptr:
            self._check(_ffi.IRP_ERR_NOMEM)
        return DeviceImage(ptr, width, height, channels, width * channels, nbytes)

    def download(self, d: DeviceImage) -> np.ndarray:
        buf = np.empty((d.height, d.pitch), np.uint8)
        self._check(self._lib.irp_memcpy_d2h(self._ctx, buf.ctypes.data, d.ptr, d.pitch * d.height))
        return np.ascontiguousarray(buf[:, : d.width * d.channels]).reshape(d.height, d.width, d.channels)

    def free(self, d: DeviceImage) -> None:
        self._check(self._lib.irp_dev_free(self._ctx, d.ptr))
        d.ptr = 0

    def pinned_empty(self, shape: Tuple[int, ...]) -> np.ndarray:
        """A page-locked uint8 host array (freed with the context)."""
        n = int(np.prod(shape))
        ptr = self._lib.irp_host_alloc_pinned(self._ctx, max(n, 1))
        if not ptr:
            self._check(_ffi.IRP_ERR_NOMEM)
        arr = np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)).reshape(shape)
        return arr

    # -- descriptor marshalling ----------------------------------------------
    def _descs(self, images: Sequence[Optional[ImageLike]], is_jpeg, orientations):
        n = len(images)
        descs = (_ffi.ImageDesc * n)()
        keep = []
        for i, im in enumerate(images):
            if im is None:
                descs[i] = _ffi.ImageDesc(None, 0, 0, 0, 0, 0, 1, 0)
                continue
            if isinstance(im, DeviceImage):
                descs[i] = _ffi.ImageDesc(im.ptr, im.pitch, im.width, im.height, im.channels, int(im.is_jpeg), im.orientation, 1)
            else:
                a = _as_hwc(im)
                keep.append(a)
                h, w, c = a.shape
                jp = is_jpeg if isinstance(is_jpeg, bool) else bool(is_jpeg[i])
                o = 1 if orientations is None else int(orientations[i])
                descs[i] = _ffi.ImageDesc(a.ctypes.data, a.strides[0], w, h, c, int(jp), o, 0)
        return descs, keep

    @staticmethod
    def _orient_of(im: ImageLike, orientations, i) -> int:
        if isinstance(im, DeviceImage):
            return im.orientation
        return 1 if orientations is None else int(orientations[i])

    def preprocess_dims(self, width: int, height: int, orientation: int = 1) -> Tuple[int, int]:
        ow, oh = C.c_int(), C.c_int()
        self._check_static(self._lib.irp_preprocess_dims(width, height, orientation, C.byref(ow), C.byref(oh)), width, height)
        return ow.value, oh.value

    def fusion_dims(self, width: int, height: int, orientation: int = 1) -> Tuple[int, int, int, int]:
        ow, oh, ox, oy = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check_static(self._lib.irp_fusion_dims(width, height, orientation, C.byref(ow), C.byref(oh), C.byref(ox), C.byref(oy)), width, height)
        return ow.value, oh.value, ox.value, oy.value

    @staticmethod
    def _check_static(rc, w, h):
        if rc:
            raise IrpError(rc, f"unsupported geometry {w}x{h}")

    def _outs(self, images, orientations, device_outputs: Optional[Sequence[DeviceImage]], fusion: bool):
        n = len(images)
        outs = (_ffi.OutDesc * n)()
        arrays: List[Optional[np.ndarray]] = [None] * n
        for i, im in enumerate(images):
            if im is None:
                continue
            w, h, c = (im.width, im.height, im.channels) if isinstance(im, DeviceImage) else (im.shape[1], im.shape[0], 1 if im.ndim == 2 else im.shape[2])
            o = self._orient_of(im, orientations, i)
            if fusion:
                ow = oh = _ffi.FUSION_CANVAS
                oc = 3
            else:
                ow, oh = self.preprocess_dims(w, h, o)
                oc = 1 if c == 1 else 3
            if device_outputs is not None:
                d = device_outputs[i]
                outs[i] = _ffi.OutDesc(d.ptr, d.pitch, d.nbytes, 0, 0, 0, 1)
            else:
                arr = np.empty((oh, ow, oc), np.uint8)
                arrays[i] = arr
                outs[i] = _ffi.OutDesc(arr.ctypes.data, ow * oc, arr.nbytes, 0, 0, 0, 0)
        return outs, arrays

    # -- the hot path ---------------------------------------------------------
    def classify_batch(self, images: Sequence[ImageLike], is_jpeg=True, raw: bool = False):
        """ClassifierService.analyze for a batch (classifier.js:40-99). Returns list of dicts."""
        descs, keep = self._descs(images, is_jpeg, None)
        res = (_ffi.Result * len(images))()
        self._check(self._lib.irp_classify_batch(self._ctx, descs, len(images), res))
        return list(res) if raw else [result_to_dict(r) for r in res]

    def preprocess_batch(self, images: Sequence[ImageLike], orientations=None, device_outputs=None):
        """preprocessImage pixel stages (imagePreprocess.js:42-53). Returns list of HxWxC arrays
        (or fills `device_outputs`)."""
        descs, keep = self._descs(images, True, orientations)
        outs, arrays = self._outs(images, orientations, device_outputs, fusion=False)
        self._check(self._lib.irp_preprocess_batch(self._ctx, descs, len(images), outs))
        if device_outputs is not None:
            for d, o in zip(device_outputs, outs):
                d.width, d.height, d.channels = o.width, o.height, o.channels
                d.pitch = o.pitch or o.width * o.channels
            return device_outputs
        return arrays

    def analyze_batch(self, images: Sequence[ImageLike], is_jpeg=True, orientations=None, device_outputs=None, raw: bool = False):
        """classify + preprocess of the same sources in one submission (BASELINE.json configs[1])."""
        descs, keep = self._descs(images, is_jpeg, orientations)
        outs, arrays = self._outs(images, orientations, device_outputs, fusion=False)
        res = (_ffi.Result * len(images))()
        self._check(self._lib.irp_analyze_batch(self._ctx, descs, len(images), res, outs))
        results = list(res) if raw else [result_to_dict(r) for r in res]
        if device_outputs is not None:
            for d, o in zip(device_outputs, outs):
                d.width, d.height, d.channels = o.width, o.height, o.channels
                d.pitch = o.pitch or o.width * o.channels
            return results, device_outputs
        return results, arrays

    # -- compressed input: baseline JPEG decoded on the device ------------------
    def jpeg_info(self, data) -> Optional[Tuple[int, int, int]]:
        """(width, height, channels) if `data` is a baseline JPEG the device decoder takes, else None."""
        k = np.frombuffer(bytes(data) if not isinstance(data, (bytes, np.ndarray)) else data, np.uint8)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        if self._lib.irp_jpeg_info(k.ctypes.data, k.size, C.byref(w), C.byref(h), C.byref(c)):
            return None
        return w.value, h.value, c.value

    def analyze_jpeg_batch(self, blobs: Sequence[bytes], orientations=None, classify: bool = True, preprocess: bool = True, raw: bool = False):
        """JPEG file bytes -> (score dicts, resized images): decode, classify and preprocess on the device."""
        n = len(blobs)
        keep = [np.frombuffer(b, np.uint8) for b in blobs]
        descs = (_ffi.JpegDesc * n)()
        outs = (_ffi.OutDesc * n)() if preprocess else None
        arrays = []
        for i, k in enumerate(keep):
            o = 1 if orientations is None else int(orientations[i])
            descs[i] = _ffi.JpegDesc(k.ctypes.data, k.size, o, 0)
            if preprocess:
                info = self.jpeg_info(k)
                if info is None:
                    raise IrpError(_ffi.IRP_ERR_UNSUPPORTED, f"blob {i} is not a JPEG the device decoder takes")
                ow, oh = self.preprocess_dims(info[0], info[1], o)
                a = np.empty((oh, ow, 1 if info[2] == 1 else 3), np.uint8)
                arrays.append(a)
                outs[i] = _ffi.OutDesc(a.ctypes.data, 0, a.nbytes, 0, 0, 0, 0)
        res = (_ffi.Result * n)() if classify else None
        self._check(self._lib.irp_analyze_jpeg_batch(self._ctx, descs, n, res, outs))
        results = None if res is None else (list(res) if raw else [result_to_dict(r) for r in res])
        return results, (arrays if preprocess else None)

    def decode_jpeg_batch(self, blobs: Sequence[bytes], scale_denom: int = 1):
        """JPEG file bytes -> decoded u8 images (HxWx3 or HxW), bit-exact with libjpeg-turbo's default decode; with
        scale_denom 2, 4 or 8 at libjpeg's reduced scale (ceil(H / denom) x ceil(W / denom): shrink-on-load)."""
        n = len(blobs)
        keep = [np.frombuffer(b, np.uint8) for b in blobs]
        descs = (_ffi.JpegDesc * n)(*[_ffi.JpegDesc(k.ctypes.data, k.size, 1, scale_denom) for k in keep])
        outs = (_ffi.OutDesc * n)()
        arrays = []
        for i, k in enumerate(keep):
            info = self.jpeg_info(k)
            if info is None:
                raise IrpError(_ffi.IRP_ERR_UNSUPPORTED, f"blob {i} is not a JPEG the device decoder takes")
            d = max(1, scale_denom)
            a = np.empty((-(-info[1] // d), -(-info[0] // d), info[2]), np.uint8)
            arrays.append(a)
            outs[i] = _ffi.OutDesc(a.ctypes.data, 0, a.nbytes, 0, 0, 0, 0)
        self._check(self._lib.irp_decode_jpeg_batch(self._ctx, descs, n, outs))
        return [a[:, :, 0] if a.shape[2] == 1 else a for a in arrays]

    # -- compressed output: baseline JPEG encoded on the device ---------------
    def _encode_call(self, fn, n, caps, icc: int = 0):
        """Run fn(outs) with host buffers of caps[i] bytes (+ the call's profile); one retry with the sizes the library asks for."""
        extra = self._icc_overhead(icc)
        for attempt in range(2):
            bufs = [np.empty(int(c) + extra, np.uint8) for c in caps]
            outs = (_ffi.JpegOut * n)(*[_ffi.JpegOut(b.ctypes.data, b.size, 0, 0, 0, 0, 0) for b in bufs])
            rc = fn(outs)
            if rc == _ffi.IRP_ERR_CAPACITY and attempt == 0:
                caps = [max(int(c), int(o.size)) * 2 for c, o in zip(caps, outs)]
                continue
            self._check(rc)
            return [bytes(memoryview(b)[:o.size]) for b, o in zip(bufs, outs)]

    # -- ICC profiles: chosen per call (IRP_JPEG_ICC), never toggled on the context ---------------
    SRGB = _ffi.ICC_SRGB   # the library's generated sRGB profile, always registered

    def register_icc(self, profile: bytes) -> int:
        """Publish an immutable profile in the context's registry; the returned id goes into `icc=` of the encode calls."""
        b = np.frombuffer(profile, np.uint8)
        rc = self._lib.irp_register_icc(self._ctx, b.ctypes.data, b.size)
        if rc < 0:
            raise IrpError(rc, "irp_register_icc rejected the profile")
        return rc

    def icc_bytes(self, icc: int) -> bytes:
        """The bytes of a registered profile (icc=Engine.SRGB: the generated sRGB one)."""
        size = C.c_size_t()
        self._check(self._lib.irp_get_icc(self._ctx, icc, None, 0, C.byref(size)))
        buf = np.empty(max(size.value, 1), np.uint8)
        self._check(self._lib.irp_get_icc(self._ctx, icc, buf.ctypes.data, buf.size, C.byref(size)))
        return bytes(memoryview(buf)[: size.value])

    def _icc_overhead(self, icc: int) -> int:
        """Bytes the APP2 segments of profile `icc` add to a file (18 per 65519-byte segment)."""
        size = C.c_size_t()
        self._check(self._lib.irp_get_icc(self._ctx, icc, None, 0, C.byref(size)))
        n = size.value
        return n + 18 * ((n + 65518) // 65519) if n else 0

    def set_output_icc(self, profile: Optional[bytes]) -> None:
        """The context DEFAULT profile (used by calls that pass icc=0); None clears it.  Set it once, e.g. at start-up:
        per-request profiles go through register_icc + `icc=`, which concurrent requests cannot disturb."""
        b = np.frombuffer(profile, np.uint8) if profile else None
        self._check(self._lib.irp_set_output_icc(self._ctx, b.ctypes.data if b is not None else None, b.size if b is not None else 0))

    def encode_jpeg_batch(self, images: Sequence[ImageLike], quality: int = 85, optimize: bool = False, icc: int = 0) -> List[bytes]:
        """u8 RGB / grey images (host arrays or DeviceImages) -> baseline 4:4:4 JPEG files, byte-identical to
        libjpeg-turbo's (imagePreprocess.js:50-53 without mozjpeg's trellis / progressive passes)."""
        n = len(images)
        descs, keep = self._descs(images, True, None)
        caps = [d.width * d.height * d.channels + 4096 for d in descs]
        q = quality | (_ffi.JPEG_OPTIMIZE if optimize else 0) | _ffi.jpeg_icc(icc)
        return self._encode_call(lambda outs: self._lib.irp_encode_jpeg_batch(self._ctx, descs, n, q, outs), n, caps, icc)

    def analyze_encode_batch(self, images: Sequence[ImageLike], is_jpeg=True, orientations=None, quality: int = 85, classify: bool = True,
                             raw: bool = False, icc: int = 0):
        """analyze() + preprocessImage() of raw pixels: (score dicts, preprocessed JPEG FILES)."""
        n = len(images)
        descs, keep = self._descs(images, is_jpeg, orientations)
        res = (_ffi.Result * n)() if classify else None
        caps = []
        for d in descs:
            ow, oh = self.preprocess_dims(d.width, d.height, d.exif_orientation)
            caps.append(ow * oh * (1 if d.channels == 1 else 3) // 2 + 4096)
        q = quality | _ffi.jpeg_icc(icc)
        files = self._encode_call(lambda outs: self._lib.irp_analyze_encode_batch(self._ctx, descs, n, res, q, outs), n, caps, icc)
        results = None if res is None else (list(res) if raw else [result_to_dict(r) for r in res])
        return results, files

    def transcode_jpeg_batch(self, blobs: Sequence[bytes], orientations=None, quality: int = 85, classify: bool = True, raw: bool = False,
                             icc: int = 0):
        """JPEG files in -> (score dicts, preprocessed JPEG files): decode, classify, resize and re-encode on the
        device; only file bytes cross PCIe in either direction."""
        n = len(blobs)
        keep = [np.frombuffer(b, np.uint8) for b in blobs]
        descs = (_ffi.JpegDesc * n)()
        caps = []
        for i, k in enumerate(keep):
            o = 1 if orientations is None else int(orientations[i])
            descs[i] = _ffi.JpegDesc(k.ctypes.data, k.size, o, 0)
            info = self.jpeg_info(k)
            if info is None:
                raise IrpError(_ffi.IRP_ERR_UNSUPPORTED, f"blob {i} is not a JPEG the device decoder takes")
            ow, oh = self.preprocess_dims(info[0], info[1], o)
            caps.append(ow * oh * (1 if info[2] == 1 else 3) // 2 + 4096)
        res = (_ffi.Result * n)() if classify else None
        q = quality | _ffi.jpeg_icc(icc)
        files = self._encode_call(lambda outs: self._lib.irp_transcode_jpeg_batch(self._ctx, descs, n, res, q, outs), n, caps, icc)
        results = None if res is None else (list(res) if raw else [result_to_dict(r) for r in res])
        return results, files

    # -- concurrent single-image requests (irp_submit / irp_wait) -------------
    def submit(self, image: ImageLike, is_jpeg: bool = True, orientation: int = 1, classify: bool = True, preprocess: bool = True):
        """Queue ONE image (as the reference's callers do, one analyze() per promise) and return a handle at
        once; concurrent submissions are batched by the context's dispatcher thread. Thread-safe."""
        descs, keep = self._descs([image], is_jpeg, [orientation])
        outs, arrays = (self._outs([image], [orientation], None, fusion=False) if preprocess else (None, [None]))
        res = _ffi.Result() if classify else None
        ticket = C.c_void_p()
        rc = self._lib.irp_submit(self._ctx, descs, C.byref(res) if classify else None, outs if preprocess else None, C.byref(ticket))
        if rc:
            raise IrpError(rc, "irp_submit rejected the request")
        return {"ticket": ticket, "descs": descs, "keep": keep, "outs": outs, "array": arrays[0], "res": res}

    def submit_jpeg(self, blob: bytes, orientation: int = 1, classify: bool = True, preprocess: bool = True):
        """Queue ONE baseline JPEG FILE (what analyze(imageBuffer) receives); decoded on the device with its batch."""
        k = np.frombuffer(blob, np.uint8)
        info = self.jpeg_info(k)
        if info is None:
            raise IrpError(_ffi.IRP_ERR_UNSUPPORTED, "not a JPEG the device decoder takes")
        desc = _ffi.JpegDesc(k.ctypes.data, k.size, orientation, 0)
        outs, arr = None, None
        if preprocess:
            ow, oh = self.preprocess_dims(info[0], info[1], orientation)
            arr = np.empty((oh, ow, 1 if info[2] == 1 else 3), np.uint8)
            outs = (_ffi.OutDesc * 1)(_ffi.OutDesc(arr.ctypes.data, 0, arr.nbytes, 0, 0, 0, 0))
        res = _ffi.Result() if classify else None
        ticket = C.c_void_p()
        rc = self._lib.irp_submit_jpeg(self._ctx, C.byref(desc), C.byref(res) if classify else None, outs, C.byref(ticket))
        if rc:
            raise IrpError(rc, "irp_submit_jpeg rejected the request")
        return {"ticket": ticket, "descs": desc, "keep": k, "outs": outs, "array": arr, "res": res}

    def submit_transcode(self, blob: bytes, orientation: int = 1, quality: int = 85, classify: bool = True, icc: int = 0):
        """Queue ONE baseline JPEG FILE for analyze() + preprocessImage(): wait() returns (scores, preprocessed FILE bytes)."""
        k = np.frombuffer(blob, np.uint8)
        info = self.jpeg_info(k)
        if info is None:
            raise IrpError(_ffi.IRP_ERR_UNSUPPORTED, "not a JPEG the device decoder takes")
        desc = _ffi.JpegDesc(k.ctypes.data, k.size, orientation, 0)
        ow, oh = self.preprocess_dims(info[0], info[1], orientation)
        # a file never exceeds its pixels by more than the header (and its profile) at q <= 95
        buf = np.empty(ow * oh * (1 if info[2] == 1 else 3) + 4096 + self._icc_overhead(icc), np.uint8)
        quality = quality | _ffi.jpeg_icc(icc)
        enc = _ffi.JpegOut(buf.ctypes.data, buf.size, 0, 0, 0, 0, 0)
        res = _ffi.Result() if classify else None
        ticket = C.c_void_p()
        rc = self._lib.irp_submit_transcode(self._ctx, C.byref(desc), C.byref(res) if classify else None, quality, C.byref(enc), C.byref(ticket))
        if rc:
            raise IrpError(rc, "irp_submit_transcode rejected the request")
        return {"ticket": ticket, "descs": desc, "keep": k, "outs": None, "array": None, "res": res, "enc": enc, "file": buf}

    def wait(self, handle, raw: bool = False):
        """Block until the request is done; returns (result dict or None, output array or None)."""
        err = C.create_string_buffer(512)
        rc = self._lib.irp_wait(self._ctx, handle["ticket"], err, len(err))
        if rc:
            raise IrpError(rc, err.value.decode(errors="replace"))
        res = handle["res"]
        if handle.get("enc") is not None:
            return (None if res is None else (res if raw else result_to_dict(res))), bytes(memoryview(handle["file"])[:handle["enc"].size])
        out = handle["array"]
        if out is not None:
            o = handle["outs"][0]
            out = out.reshape(-1)[: o.height * o.width * o.channels].reshape(o.height, o.width, o.channels)
        return (None if res is None else (res if raw else result_to_dict(res))), out

    def fusion_prepare_batch(self, groups: Sequence[Sequence[Optional[ImageLike]]], orientations=None, device_outputs=None):
        """Each group of <= 3 images -> aligned 2048x2048x3 canvases."""
        flat: List[Optional[ImageLike]] = []
        flat_or = []
        for gi, g in enumerate(groups):
            g = list(g)
            if len(g) > _ffi.FUSION_MAX_IMAGES:
                raise ValueError("a fusion group holds at most 3 images")
            for k in range(_ffi.FUSION_MAX_IMAGES):
                flat.append(g[k] if k < len(g) else None)
                flat_or.append(1 if orientations is None or k >= len(g) else int(orientations[gi][k]))
        descs, keep = self._descs(flat, True, flat_or)
        outs, arrays = self._outs(flat, flat_or, device_outputs, fusion=True)
        self._check(self._lib.irp_fusion_prepare_batch(self._ctx, descs, len(groups), outs))
        if device_outputs is not None:
            return device_outputs
        return [[arrays[gi * 3 + k] for k in range(len(list(g)))] for gi, g in enumerate(groups)]
