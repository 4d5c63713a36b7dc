"""ctypes binding of libirp_b200.so — the C ABI declared in include/irp.h.

This is the Python counterpart of the N-API addon (addon/irp_addon.cc): zero
logic, struct mirrors and argtypes only.  Loading fails loudly when the CUDA
library has not been built; there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IRP_LIB_PATH") or os.path.join(_HERE, "libirp_b200.so")  # override: kernel experiments only

IRP_OK = 0
IRP_ERR_BAD_ARG, IRP_ERR_UNSUPPORTED, IRP_ERR_CUDA, IRP_ERR_NOMEM, IRP_ERR_NO_DEVICE, IRP_ERR_CAPACITY = -1, -2, -3, -4, -5, -6
ICC_SRGB = 1              # IRP_ICC_SRGB: the library's generated sRGB profile
JPEG_OPTIMIZE = 0x100     # IRP_JPEG_OPTIMIZE


def jpeg_icc(profile_id: int) -> int:
    """IRP_JPEG_ICC(id): the profile of one encode call, OR-ed into `quality`."""
    return (profile_id & 0xFF) << 16


FUSION_CANVAS = 2048
FUSION_MAX_IMAGES = 3

# every symbol include/irp.h declares (tests/test_abi.py checks the .so exports each one)
SYMBOLS = (
    "irp_abi_version", "irp_device_count", "irp_create", "irp_destroy", "irp_last_error", "irp_set_stream",
    "irp_get_timing", "irp_preprocess_dims", "irp_fusion_dims", "irp_scores_from_moments", "irp_top_issues", "irp_grey_tables",
    "irp_classify_batch", "irp_preprocess_batch", "irp_analyze_batch", "irp_fusion_prepare_batch",
    "irp_submit", "irp_submit_jpeg", "irp_submit_transcode", "irp_wait", "irp_jpeg_info", "irp_decode_jpeg_batch", "irp_analyze_jpeg_batch",
    "irp_set_output_icc", "irp_register_icc", "irp_get_icc", "irp_encode_jpeg_batch", "irp_analyze_encode_batch", "irp_transcode_jpeg_batch",
    "irp_dev_alloc", "irp_dev_free", "irp_host_alloc_pinned", "irp_host_free_pinned", "irp_memcpy_h2d",
    "irp_memcpy_d2h", "irp_synchronize",
)


class Opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("luma_mode", C.c_int32), ("coef_mode", C.c_int32),
                ("reserved0", C.c_int32), ("staging_bytes", C.c_uint64), ("blur_mode", C.c_int32), ("reduce_mode", C.c_int32)]


class ImageDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("pitch", C.c_size_t), ("width", C.c_int32), ("height", C.c_int32),
                ("channels", C.c_int32), ("is_jpeg", C.c_int32), ("exif_orientation", C.c_int32),
                ("on_device", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("score", C.c_double * 7), ("sum", C.c_uint64 * 4), ("sumsq", C.c_uint64 * 4),
                ("e_sum", C.c_uint64 * 2), ("e_sumsq", C.c_uint64 * 2), ("b_sum", C.c_uint64),
                ("b_sumsq", C.c_uint64), ("scratch_v", C.c_uint32), ("scratch_h", C.c_uint32),
                ("block_edges", C.c_uint32 * 2), ("luma_hist", C.c_uint32 * 256), ("status", C.c_int32),
                ("issues", C.c_uint8 * 4)]


class OutDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("pitch", C.c_size_t), ("capacity", C.c_size_t), ("width", C.c_int32),
                ("height", C.c_int32), ("channels", C.c_int32), ("on_device", C.c_int32)]


class JpegDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("exif_orientation", C.c_int32), ("scale_denom", C.c_int32)]


class JpegOut(C.Structure):
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("size", C.c_size_t), ("width", C.c_int32),
                ("height", C.c_int32), ("channels", C.c_int32), ("reserved", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("classify_ms", C.c_float), ("preprocess_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("kernel_launches", C.c_uint32),
                ("chunks", C.c_uint32)]


_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library (built by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). irp_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int
    lib.irp_abi_version.restype = i32
    lib.irp_device_count.restype = i32
    lib.irp_create.restype = vp
    lib.irp_create.argtypes = [i32, C.POINTER(Opts)]
    lib.irp_destroy.restype = None
    lib.irp_destroy.argtypes = [vp]
    lib.irp_last_error.restype = C.c_char_p
    lib.irp_last_error.argtypes = [vp]
    lib.irp_set_stream.argtypes = [vp, vp]
    lib.irp_get_timing.argtypes = [vp, C.POINTER(Timing)]
    ip = C.POINTER(C.c_int)
    lib.irp_preprocess_dims.argtypes = [i32, i32, i32, ip, ip]
    lib.irp_fusion_dims.argtypes = [i32, i32, i32, ip, ip, ip, ip]
    lib.irp_scores_from_moments.argtypes = [C.POINTER(Result), i32, i32, i32, i32]
    lib.irp_grey_tables.argtypes = [i32, vp, vp, vp, vp]
    lib.irp_classify_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, C.POINTER(Result)]
    lib.irp_preprocess_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, C.POINTER(OutDesc)]
    lib.irp_analyze_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, C.POINTER(Result), C.POINTER(OutDesc)]
    lib.irp_fusion_prepare_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, C.POINTER(OutDesc)]
    lib.irp_jpeg_info.argtypes = [vp, C.c_size_t, ip, ip, ip]
    lib.irp_decode_jpeg_batch.argtypes = [vp, C.POINTER(JpegDesc), i32, C.POINTER(OutDesc)]
    lib.irp_analyze_jpeg_batch.argtypes = [vp, C.POINTER(JpegDesc), i32, C.POINTER(Result), C.POINTER(OutDesc)]
    lib.irp_set_output_icc.argtypes = [vp, vp, C.c_size_t]
    lib.irp_register_icc.argtypes = [vp, vp, C.c_size_t]
    lib.irp_get_icc.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.irp_encode_jpeg_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, i32, C.POINTER(JpegOut)]
    lib.irp_analyze_encode_batch.argtypes = [vp, C.POINTER(ImageDesc), i32, C.POINTER(Result), i32, C.POINTER(JpegOut)]
    lib.irp_transcode_jpeg_batch.argtypes = [vp, C.POINTER(JpegDesc), i32, C.POINTER(Result), i32, C.POINTER(JpegOut)]
    lib.irp_submit.argtypes = [vp, C.POINTER(ImageDesc), C.POINTER(Result), C.POINTER(OutDesc), C.POINTER(vp)]
    lib.irp_submit_jpeg.argtypes = [vp, C.POINTER(JpegDesc), C.POINTER(Result), C.POINTER(OutDesc), C.POINTER(vp)]
    lib.irp_submit_transcode.argtypes = [vp, C.POINTER(JpegDesc), C.POINTER(Result), i32, C.POINTER(JpegOut), C.POINTER(vp)]
    lib.irp_wait.argtypes = [vp, vp, C.c_char_p, C.c_size_t]
    lib.irp_dev_alloc.restype = vp
    lib.irp_dev_alloc.argtypes = [vp, C.c_size_t]
    lib.irp_dev_free.argtypes = [vp, vp]
    lib.irp_host_alloc_pinned.restype = vp
    lib.irp_host_alloc_pinned.argtypes = [vp, C.c_size_t]
    lib.irp_host_free_pinned.argtypes = [vp, vp]
    lib.irp_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.irp_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.irp_synchronize.argtypes = [vp]
    _lib = lib
    return lib
