"""The C-ABI library loads on a CPU-only box and exports every symbol include/irp.h declares; the pure
host helpers work without a GPU; anything that needs pixels fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "irp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(irp_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    from irp_b200 import _ffi

    lib = _ffi.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/irp.h but not exported"
    assert sorted(_ffi.SYMBOLS) == declared, "the ctypes binding and the header disagree"
    assert lib.irp_abi_version() == 1


def test_struct_layouts_match_the_oracle_mirror():
    from irp_b200 import _ffi
    from oracle import oracle

    assert C.sizeof(_ffi.Result) == C.sizeof(oracle.Result) == 7 * 8 + 8 * 8 + 4 * 8 + 2 * 8 + 4 * 4 + 256 * 4 + 8
    assert C.sizeof(_ffi.ImageDesc) == C.sizeof(oracle.ImageDesc) == 40
    for f, _ in _ffi.Result._fields_:
        assert getattr(_ffi.Result, f).offset == getattr(oracle.Result, f).offset


@pytest.mark.parametrize("w,h,o", [(4000, 3000, 1), (3840, 2160, 1), (6000, 4000, 1), (4000, 3000, 6), (2048, 2048, 3), (1024, 768, 5),
                                   (2049, 100, 1), (100, 2049, 3), (5000, 5000, 7), (8000, 4000, 2)])
def test_preprocess_and_fusion_dims_match_oracle(oracle, w, h, o):
    from irp_b200 import _ffi

    lib = _ffi.load()
    ow, oh, ox, oy = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.irp_preprocess_dims(w, h, o, C.byref(ow), C.byref(oh)) == 0
    assert (ow.value, oh.value) == oracle.preprocess_dims(w, h, o)[:2]
    assert lib.irp_fusion_dims(w, h, o, C.byref(ow), C.byref(oh), C.byref(ox), C.byref(oy)) == 0
    assert (ow.value, oh.value, ox.value, oy.value) == oracle.fusion_dims(w, h, o)[:4]


def test_large_shrinks_are_sized_by_the_host_helper(oracle):
    """Shrink factors of 4 and more (sides beyond 8192 px) used to be refused; they now take libvips' integer box
    pre-shrink in front of the lanczos passes, and the host helper sizes them like the oracle."""
    from irp_b200 import _ffi

    ow, oh = C.c_int(), C.c_int()
    for w, h, o in [(9000, 16, 1), (12000, 9000, 1), (24000, 300, 6), (100, 2049, 8)]:   # the last: the pre-rotation-dims quirk of SURVEY.md section 8a P3
        assert _ffi.load().irp_preprocess_dims(w, h, o, C.byref(ow), C.byref(oh)) == 0
        assert (ow.value, oh.value) == oracle.preprocess_dims(w, h, o)[:2]
    assert _ffi.load().irp_preprocess_dims(0, 16, 1, C.byref(ow), C.byref(oh)) == _ffi.IRP_ERR_BAD_ARG


@pytest.mark.parametrize("c,jpeg", [(3, True), (3, False), (1, True), (4, True)])
def test_scores_from_moments_match_the_js_literal_oracle(oracle, c, jpeg):
    """The product computes the seven scores from exact integer moments; the oracle walks the buffers
    the way the JS does (two-pass, sequential doubles). They must agree to 1e-4 relative (north_star)."""
    from conftest import rand_image, rel_close
    from irp_b200 import _ffi

    img = rand_image(97, 131, c, seed=c, kind="smooth")
    ref = oracle.classify(img, is_jpeg=jpeg)
    r = _ffi.Result()
    for k in ("sum", "sumsq", "e_sum", "e_sumsq", "block_edges", "luma_hist"):
        for i, v in enumerate(ref[k]):
            getattr(r, k)[i] = v
    r.b_sum, r.b_sumsq, r.scratch_v, r.scratch_h = ref["b_sum"], ref["b_sumsq"], ref["scratch_v"], ref["scratch_h"]
    assert _ffi.load().irp_scores_from_moments(C.byref(r), 131, 97, c, int(jpeg)) == 0
    for i, k in enumerate(oracle.SCORE_KEYS):
        assert rel_close(r.score[i], ref["scores"][k]), (k, r.score[i], ref["scores"][k])


def test_flat_image_scores_are_exact_from_moments():
    from irp_b200 import _ffi

    n = 128 * 128
    r = _ffi.Result()
    for ch in range(3):
        r.sum[ch], r.sumsq[ch] = 180 * n, 180 * 180 * n
    r.e_sum[1], r.e_sumsq[1] = 180 * n, 180 * 180 * n  # Sharp9 of a flat image is the image
    r.b_sum, r.b_sumsq = 3 * 180 * n, 3 * 180 * 180 * n
    _ffi.load().irp_scores_from_moments(C.byref(r), 128, 128, 3, 1)
    assert list(r.score) == [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0]


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU irp_create must fail with IRP_ERR_NO_DEVICE and a message; with a GPU
    this test only checks that a bad device index is refused."""
    import irp_b200
    from irp_b200 import _ffi

    lib = _ffi.load()
    if lib.irp_device_count() == 0:
        with pytest.raises(irp_b200.IrpError) as e:
            irp_b200.Engine(0)
        assert e.value.code == _ffi.IRP_ERR_NO_DEVICE and "no CPU fallback" in e.value.message
    else:
        with pytest.raises(irp_b200.IrpError):
            irp_b200.Engine(10_000)
    assert lib.irp_classify_batch(None, None, 0, None) == _ffi.IRP_ERR_BAD_ARG


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package or include/ may reference it."""
    pkg = os.path.join(ROOT, "image-restoration-platform_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "grey_tables.inc", f"{f} mentions the oracle"


def test_jpeg_header_parse_runs_without_a_gpu():
    """irp_jpeg_info is pure host code (sharp's metadata() for the device-decodable subset): baseline files give
    their stored dims, everything else is refused with IRP_ERR_UNSUPPORTED."""
    import ctypes as C
    import io

    import numpy as np
    from PIL import Image

    from irp_b200 import _ffi

    lib = _ffi.load()

    def info(blob):
        k = np.frombuffer(blob, np.uint8)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        rc = lib.irp_jpeg_info(k.ctypes.data, k.size, C.byref(w), C.byref(h), C.byref(c))
        return rc, (w.value, h.value, c.value)

    def enc(arr, **kw):
        b = io.BytesIO()
        Image.fromarray(arr).save(b, "JPEG", **kw)
        return b.getvalue()

    rgb = np.zeros((37, 53, 3), np.uint8)
    for ss in (0, 1, 2):
        assert info(enc(rgb, quality=80, subsampling=ss)) == (0, (53, 37, 3))
    assert info(enc(rgb[:, :, 0], quality=80)) == (0, (53, 37, 1))
    assert info(enc(rgb, quality=80, restart_marker_rows=1)) == (0, (53, 37, 3))
    assert info(enc(rgb, quality=80, progressive=True)) == (0, (53, 37, 3))
    b4 = io.BytesIO()
    Image.fromarray(rgb).convert("CMYK").save(b4, "JPEG", quality=80)
    assert info(b4.getvalue())[0] == _ffi.IRP_ERR_UNSUPPORTED   # four components
    assert info(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)[0] == _ffi.IRP_ERR_UNSUPPORTED
    assert info(enc(rgb, quality=80)[:40])[0] == _ffi.IRP_ERR_UNSUPPORTED   # truncated inside the header


def test_progressive_scan_headers_are_validated_on_the_host():
    """jdphuff.c's rules for scan parameters are enforced by the parser (irp_jpeg_info walks every scan of a multi-scan
    file, no GPU involved): the device only ever sees scans it can walk."""
    import ctypes as C
    import io

    import numpy as np
    from PIL import Image

    from irp_b200 import _ffi

    lib = _ffi.load()

    def info(blob):
        k = np.frombuffer(bytes(blob), np.uint8)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        return lib.irp_jpeg_info(k.ctypes.data, k.size, C.byref(w), C.byref(h), C.byref(c))

    b = io.BytesIO()
    Image.fromarray(np.zeros((40, 56, 3), np.uint8)).save(b, "JPEG", quality=80, progressive=True)
    good = b.getvalue()
    assert info(good) == 0
    sos = [i for i in range(len(good) - 1) if good[i] == 0xFF and good[i + 1] == 0xDA]
    assert len(sos) >= 6

    def patched(k, off, val):       # byte `off` of the k-th scan header's payload (negative: from its end)
        p = sos[k]
        ln = (good[p + 2] << 8) | good[p + 3]
        out = bytearray(good)
        out[p + 2 + ln + off if off < 0 else p + 4 + off] = val
        return out

    assert info(patched(1, -3, 70)) == _ffi.IRP_ERR_UNSUPPORTED      # Ss > 63
    assert info(patched(1, -2, 0)) == _ffi.IRP_ERR_UNSUPPORTED       # Se < Ss
    assert info(patched(1, -1, 0x0F)) == _ffi.IRP_ERR_UNSUPPORTED    # Al > 13
    assert info(patched(0, -2, 5)) == _ffi.IRP_ERR_UNSUPPORTED       # a DC scan with Se != 0
    assert info(patched(0, 3, good[sos[0] + 5])) == _ffi.IRP_ERR_UNSUPPORTED   # the same component twice in one scan
    assert info(patched(1, 2, 0x09)) == _ffi.IRP_ERR_UNSUPPORTED     # table selector > 3
    assert info(patched(1, 1, 0x7E)) == _ffi.IRP_ERR_UNSUPPORTED     # a component the frame does not have
    eoi = good.rfind(b"\xff\xd9")
    one = good[sos[-1]:eoi]
    assert info(good[:eoi] + one * 70 + b"\xff\xd9") == _ffi.IRP_ERR_UNSUPPORTED   # more scans than any script holds


def test_parser_survives_mutated_headers():
    """irp_jpeg_info on thousands of randomly damaged baseline and progressive files: every call returns (accepted or
    IRP_ERR_UNSUPPORTED), none reads outside the buffer it was given (the copy ends exactly at the file's last byte)."""
    import ctypes as C
    import io

    import numpy as np
    from PIL import Image

    from irp_b200 import _ffi

    lib = _ffi.load()
    rng = np.random.default_rng(3)
    seen = set()
    for prog in (True, False):
        b = io.BytesIO()
        Image.fromarray(rng.integers(0, 255, (48, 64, 3), dtype=np.uint8)).save(b, "JPEG", quality=80, progressive=prog, subsampling=2)
        good = b.getvalue()
        for t in range(4000):
            a = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                a[int(rng.integers(2, len(a)))] = int(rng.integers(0, 256))
            if t % 7 == 0:
                a = a[: int(rng.integers(4, len(a)))]
            k = np.frombuffer(bytes(a), np.uint8)
            w, h, c = C.c_int(), C.c_int(), C.c_int()
            seen.add(lib.irp_jpeg_info(k.ctypes.data, k.size, C.byref(w), C.byref(h), C.byref(c)))
    assert seen <= {0, _ffi.IRP_ERR_UNSUPPORTED}


def test_committed_jpeg_tables_are_what_libjpeg_turbo_writes(tmp_path):
    """csrc/jpeg_std_tables.inc (the encoder's Annex K quantisation / Huffman tables) is generated from files
    written by libjpeg-turbo; regenerating it must reproduce the committed file byte for byte."""
    pytest.importorskip("PIL")
    import shutil
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "image-restoration-platform_b200", "csrc", "jpeg_std_tables.inc")
    committed = open(inc).read()
    work = tmp_path / "repo"
    (work / "tools").mkdir(parents=True)
    (work / "image-restoration-platform_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(root, "tools", "gen_jpeg_std_tables.py"), work / "tools" / "gen_jpeg_std_tables.py")
    subprocess.run([sys.executable, str(work / "tools" / "gen_jpeg_std_tables.py")], check=True, capture_output=True)
    assert open(work / "image-restoration-platform_b200" / "csrc" / "jpeg_std_tables.inc").read() == committed


def test_generated_srgb_profile_is_a_valid_srgb_profile():
    """`.withMetadata({icc:'sRGB'})` (imagePreprocess.js:63): the library generates its own sRGB profile (libvips' file
    is not redistributed).  LittleCMS must parse it and it must be colorimetrically LittleCMS' own sRGB: converting
    through either profile to Lab gives the same numbers; the bytes are reproducible (fixed date)."""
    import ctypes as C
    import io

    import numpy as np
    from PIL import Image, ImageCms

    from irp_b200 import _ffi

    lib = _ffi.load()
    size = C.c_size_t()
    assert lib.irp_get_icc(None, _ffi.ICC_SRGB, None, 0, C.byref(size)) == 0 and 400 < size.value < 1024
    buf = (C.c_uint8 * size.value)()
    assert lib.irp_get_icc(None, _ffi.ICC_SRGB, buf, size.value, C.byref(size)) == 0
    prof = bytes(buf)
    assert int.from_bytes(prof[:4], "big") == len(prof) and prof[36:40] == b"acsp" and prof[12:20] == b"mntrRGB "
    assert lib.irp_get_icc(None, 0, None, 0, C.byref(size)) == _ffi.IRP_ERR_BAD_ARG   # other ids need a context
    mine = ImageCms.ImageCmsProfile(io.BytesIO(prof))
    assert ImageCms.getProfileDescription(mine).startswith("sRGB IEC61966-2.1")
    theirs = ImageCms.createProfile("sRGB")
    lab = ImageCms.createProfile("LAB")
    rng = np.random.default_rng(3)
    px = np.concatenate([rng.integers(0, 256, (64, 64, 3), dtype=np.uint8),
                         np.repeat(np.arange(256, dtype=np.uint8)[None, :64, None], 3, axis=2).repeat(4, 0)], axis=0)
    im = Image.fromarray(px)
    a = np.asarray(ImageCms.profileToProfile(im, mine, lab, outputMode="LAB"), dtype=int)
    b = np.asarray(ImageCms.profileToProfile(im, theirs, lab, outputMode="LAB"), dtype=int)
    assert np.abs(a - b).max() <= 1
