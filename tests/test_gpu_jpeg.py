"""Device JPEG decode (irp_decode_jpeg_batch / irp_analyze_jpeg_batch) against libjpeg-turbo itself (Pillow)
and against the pinned CPU restatement: decoded pixels bit-exact for every chroma subsampling, odd sizes,
restart intervals, optimised Huffman tables, and batches; then the classify / preprocess results computed from
the compressed bytes must equal those computed from Pillow's pixels."""
import ctypes as C
import io

import numpy as np
import pytest
from PIL import Image, ImageFile

from conftest import assert_result_parity, rand_image

pytestmark = pytest.mark.gpu


ImageFile.MAXBLOCK = 1 << 24   # progressive files of noisy images outgrow Pillow's default encoder buffer


def _encode(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def _pillow(data):
    return np.asarray(Image.open(io.BytesIO(data)))


def _decode_batch(engine, blobs):
    from irp_b200 import _ffi

    n = len(blobs)
    keep = [np.frombuffer(b, np.uint8) for b in blobs]
    descs = (_ffi.JpegDesc * n)()
    outs = (_ffi.OutDesc * n)()
    arrays = []
    for i, k in enumerate(keep):
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        assert engine._lib.irp_jpeg_info(k.ctypes.data, k.size, C.byref(w), C.byref(h), C.byref(c)) == 0
        a = np.empty((h.value, w.value, c.value), np.uint8)
        arrays.append(a)
        descs[i] = _ffi.JpegDesc(k.ctypes.data, k.size, 1, 0)
        outs[i] = _ffi.OutDesc(a.ctypes.data, 0, a.nbytes, 0, 0, 0, 0)
    engine._check(engine._lib.irp_decode_jpeg_batch(engine._ctx, descs, n, outs))
    return [a[:, :, 0] if a.shape[2] == 1 else a for a in arrays]


SHAPES = [(8, 8), (1, 1), (3, 5), (5, 3), (9, 4), (16, 2), (17, 33), (37, 53), (64, 48), (100, 161), (241, 319), (600, 900)]   # widths 2-4: chroma two samples wide or less takes libjpeg's plain upsampler


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_decode_matches_libjpeg_turbo(engine, subsampling):
    blobs = []
    for i, (h, w) in enumerate(SHAPES):
        for kind, q in (("smooth", 85), ("noise", 60), ("edges", 95)):
            blobs.append(_encode(rand_image(h, w, 3, seed=31 * i + subsampling, kind=kind), quality=q, subsampling=subsampling))
    got = _decode_batch(engine, blobs)
    for i, (g, b) in enumerate(zip(got, blobs)):
        ref = _pillow(b)
        assert g.shape == ref.shape and np.array_equal(g, ref), f"blob {i} shape {ref.shape}: max diff {np.abs(g.astype(int) - ref).max()}"


def test_greyscale_restart_intervals_and_optimised_tables(engine):
    blobs = [_encode(rand_image(130, 70, 1, seed=1, kind="smooth")[:, :, 0], quality=80),
             _encode(rand_image(75, 131, 3, seed=2, kind="smooth"), quality=75, subsampling=2, restart_marker_rows=1),
             _encode(rand_image(75, 131, 3, seed=3, kind="noise"), quality=75, subsampling=0, restart_marker_blocks=3),
             _encode(rand_image(211, 333, 3, seed=4, kind="smooth"), quality=90, subsampling=2, optimize=True),
             _encode(rand_image(40, 56, 3, seed=5, kind="edges"), quality=1, subsampling=1),
             _encode(np.zeros((33, 47, 3), np.uint8), quality=50), _encode(np.full((33, 47, 3), 255, np.uint8), quality=100)]
    for g, b in zip(_decode_batch(engine, blobs), blobs):
        ref = _pillow(b)
        assert g.shape == ref.shape and np.array_equal(g, ref)


def test_a_large_photo_sized_image_synchronises(engine):
    """Many thousand subsequences in one stream: the self-synchronising decode must converge and match."""
    img = rand_image(1500, 2000, 3, seed=8, kind="smooth")
    blob = _encode(img, quality=88, subsampling=2)
    got = _decode_batch(engine, [blob])[0]
    assert np.array_equal(got, _pillow(blob))


def test_analyze_from_compressed_bytes_equals_analyze_from_pixels(engine, oracle):
    from irp_b200 import _ffi

    imgs = [rand_image(900, 1300, 3, seed=11, kind="smooth"), rand_image(2300, 2100, 3, seed=12, kind="smooth")]
    blobs = [_encode(im, quality=85, subsampling=2) for im in imgs]
    orients = [1, 6]
    keep = [np.frombuffer(b, np.uint8) for b in blobs]
    n = len(blobs)
    descs = (_ffi.JpegDesc * n)(*[_ffi.JpegDesc(k.ctypes.data, k.size, o, 0) for k, o in zip(keep, orients)])
    res = (_ffi.Result * n)()
    outs = (_ffi.OutDesc * n)()
    arrays = []
    for i, b in enumerate(blobs):
        px = _pillow(b)
        ow, oh = engine.preprocess_dims(px.shape[1], px.shape[0], orients[i])
        a = np.empty((oh, ow, 3), np.uint8)
        arrays.append(a)
        outs[i] = _ffi.OutDesc(a.ctypes.data, 0, a.nbytes, 0, 0, 0, 0)
    engine._check(engine._lib.irp_analyze_jpeg_batch(engine._ctx, descs, n, res, outs))
    from irp_b200.engine import result_to_dict

    for i, b in enumerate(blobs):
        px = np.ascontiguousarray(_pillow(b))
        assert_result_parity(result_to_dict(res[i]), oracle.classify(px), 3, f"jpeg {i}")
        assert np.array_equal(arrays[i], oracle.preprocess(px, orients[i]))


PROG_SHAPES = [(8, 8), (1, 1), (3, 5), (17, 33), (37, 53), (100, 161), (241, 319), (600, 900)]


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_progressive_matches_libjpeg_turbo(engine, subsampling):
    """SOF2 files (spectral selection + successive approximation: libjpeg's standard script, what `progressive: true`
    writes — imagePreprocess.js:57-61) decoded on the device scan by scan (prog_scan_kernel): pixels equal to
    libjpeg-turbo's and to the pinned restatement, for fixed and optimised tables, with restart intervals, in one batch
    together with baseline files."""
    from oracle import jpeg_oracle

    blobs = []
    for i, (h, w) in enumerate(PROG_SHAPES):
        img = rand_image(h, w, 3, seed=200 + i, kind="smooth" if i % 2 else "noise")
        kw = [{}, {"optimize": True}, {"restart_marker_blocks": 5}][i % 3]
        blobs.append(_encode(img, quality=(35, 80, 95)[i % 3], subsampling=subsampling, progressive=True, **kw))
        if i % 4 == 0:
            blobs.append(_encode(img, quality=85, subsampling=subsampling))          # a baseline neighbour in the same launch
    blobs.append(_encode(rand_image(77, 130, 1, seed=9)[:, :, 0], quality=70, progressive=True))   # greyscale
    got = _decode_batch(engine, blobs)
    for b, g in zip(blobs, got):
        ref = _pillow(b)
        assert np.array_equal(g, ref), f"{ref.shape}: max diff {np.abs(g.astype(int) - ref).max()}"
        assert np.array_equal(g, jpeg_oracle.decode(b))


def test_progressive_photo_sized_and_analyzed(engine, oracle):
    """A 3 MP progressive file — the size preprocessImage hands on — through irp_analyze_jpeg_batch: scores and
    preprocessed pixels equal those computed from libjpeg-turbo's pixels."""
    from irp_b200 import _ffi
    from irp_b200.engine import result_to_dict

    img = rand_image(1536, 2048, 3, seed=77, kind="smooth")
    blob = _encode(img, quality=85, subsampling=2, progressive=True, optimize=True)
    px = np.ascontiguousarray(_pillow(blob))
    assert np.array_equal(_decode_batch(engine, [blob])[0], px)
    keep = np.frombuffer(blob, np.uint8)
    descs = (_ffi.JpegDesc * 1)(_ffi.JpegDesc(keep.ctypes.data, keep.size, 1, 0))
    res = (_ffi.Result * 1)()
    ow, oh = engine.preprocess_dims(px.shape[1], px.shape[0], 1)
    a = np.empty((oh, ow, 3), np.uint8)
    outs = (_ffi.OutDesc * 1)(_ffi.OutDesc(a.ctypes.data, 0, a.nbytes, 0, 0, 0, 0))
    engine._check(engine._lib.irp_analyze_jpeg_batch(engine._ctx, descs, 1, res, outs))
    assert_result_parity(result_to_dict(res[0]), oracle.classify(px), 3, "progressive")
    assert np.array_equal(a, oracle.preprocess(px, 1))


def test_progressive_upload_through_the_whole_upload_route(engine, oracle):
    """analyze(buf) + preprocessImage(buf) of ONE progressive upload with files on both sides (irp_transcode_jpeg_batch):
    the returned file is libjpeg-turbo's encoding of the oracle's preprocessed pixels of libjpeg-turbo's decode."""
    img = rand_image(900, 1300, 3, seed=31, kind="smooth")
    blob = _encode(img, quality=88, subsampling=2, progressive=True)
    px = np.ascontiguousarray(_pillow(blob))
    res, files = engine.transcode_jpeg_batch([blob], quality=85)
    assert_result_parity(res[0], oracle.classify(px), 3, "progressive upload")
    b = io.BytesIO()
    Image.fromarray(oracle.preprocess(px, 1)).save(b, "JPEG", quality=85, subsampling=0)
    assert files[0] == b.getvalue()


@pytest.mark.parametrize("denom", [2, 4, 8])
def test_reduced_size_decode_matches_libjpeg_turbo(engine, denom):
    """irp_jpeg_desc.scale_denom: libjpeg's scale 1/2, 1/4, 1/8 (libvips' shrink-on-load) on the device — reduced IDCTs
    (jidctred.c), per-component IDCT sizes, plain upsampling at 1/8 — against Pillow's draft mode (libjpeg-turbo itself)
    and the pinned restatement; baseline and progressive files of every sampling in one batch."""
    from oracle import jpeg_oracle

    blobs = []
    for i, (h, w) in enumerate([(64, 64), (333, 517), (17, 9), (100, 161), (255, 257), (8, 8), (5, 3), (31, 33), (600, 900)]):
        img = rand_image(h, w, 3, seed=300 + i, kind="smooth" if i % 2 else "noise")
        for sub in (0, 1, 2):
            blobs.append(_encode(img, quality=85, subsampling=sub, progressive=bool((i + sub) % 2)))
    blobs.append(_encode(rand_image(77, 130, 1, seed=9)[:, :, 0], quality=70))
    got = engine.decode_jpeg_batch(blobs, scale_denom=denom)
    pinned = 0
    for b, g in zip(blobs, got):
        assert np.array_equal(g, jpeg_oracle.decode_scaled(b, denom)), f"{g.shape}"
        im = Image.open(io.BytesIO(b))
        w, h = im.size
        im.draft(im.mode, (max(1, w // denom), max(1, h // denom)))
        ref = np.asarray(im)
        if ref.shape[:2] == g.shape[:2]:     # Pillow settled for this scale
            assert np.array_equal(g, ref)
            pinned += 1
    assert pinned >= 10
    for b in (blobs[2], blobs[14]):          # a 4:2:0 file ALONE in its call (no other image brings the generic colour kernel along)
        assert np.array_equal(engine.decode_jpeg_batch([b], scale_denom=denom)[0], jpeg_oracle.decode_scaled(b, denom))
    with pytest.raises(Exception):           # the analysing entry points need the full picture
        from irp_b200 import _ffi

        k = np.frombuffer(blobs[0], np.uint8)
        descs = (_ffi.JpegDesc * 1)(_ffi.JpegDesc(k.ctypes.data, k.size, 1, denom))
        res = (_ffi.Result * 1)()
        engine._check(engine._lib.irp_analyze_jpeg_batch(engine._ctx, descs, 1, res, None))


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_sequential_file_with_one_scan_per_component(engine, subsampling):
    """A baseline file whose components come in separate scans takes the multi-scan route (prog_scan_kernel's first-pass
    handler with DC and AC in one scan); same pixels as the interleaved original."""
    from jpeg_rescan import one_scan_per_component

    blobs, refs = [], []
    for i, (h, w) in enumerate([(37, 53), (64, 64), (100, 161), (9, 17), (241, 319)]):
        data = _encode(rand_image(h, w, 3, seed=70 + i, kind="smooth"), quality=85, subsampling=subsampling)
        blobs.append(one_scan_per_component(data))
        refs.append(_pillow(data))
    for g, r in zip(_decode_batch(engine, blobs), refs):
        assert np.array_equal(g, r)


def test_progressive_scan_scripts_beyond_the_standard_one(engine):
    """Scan scripts with several predecessors per scan, refinements of sub-bands in another order than their first
    passes, DC scans per component and three approximation levels: the block-wise scan pipeline (a scan trails every
    earlier scan that touched its coefficients) has to give the baseline file's pixels back for each of them — files of
    all scripts and samplings in ONE batch, so that many dependency chains run side by side."""
    from jpeg_rescan import SCRIPTS, progressive_with_script

    blobs, refs = [], []
    for i, (h, w) in enumerate([(37, 53), (64, 64), (100, 161), (241, 319), (9, 17)]):
        for sub in (0, 1, 2):
            data = _encode(rand_image(h, w, 3, seed=40 + i, kind="smooth" if (i + sub) % 2 else "noise"), quality=90, subsampling=sub)
            for script in SCRIPTS.values():
                blobs.append(progressive_with_script(data, script))
                refs.append(_pillow(data))
    for g, r in zip(_decode_batch(engine, blobs), refs):
        assert np.array_equal(g, r)
    big = _encode(rand_image(320, 480, 3, seed=55, kind="noise"), quality=92, subsampling=2)
    # long scans: the dependents really do run while their predecessors are still at it
    assert np.array_equal(_decode_batch(engine, [progressive_with_script(big, SCRIPTS["fine bands refined one by one"])])[0], _pillow(big))


def test_random_progressive_scan_scripts(engine):
    """Random valid progressions in one batch: whatever order the scans of a file come in, every scan only trails the
    scans that wrote its coefficients before it."""
    from jpeg_rescan import progressive_with_script, random_script

    rng = np.random.default_rng(12)
    blobs, refs = [], []
    for t in range(24):
        h, w, sub = int(rng.integers(8, 140)), int(rng.integers(8, 200)), int(rng.integers(0, 3))
        data = _encode(rand_image(h, w, 3, seed=t, kind="noise" if t % 2 else "smooth"), quality=int(rng.integers(60, 96)), subsampling=sub)
        blobs.append(progressive_with_script(data, random_script(rng)))
        refs.append(_pillow(data))
    for t, (g, r) in enumerate(zip(_decode_batch(engine, blobs), refs)):
        assert np.array_equal(g, r), f"case {t}"


def test_unsupported_kinds_are_refused_loudly(engine):
    from irp_b200 import _ffi

    cmyk = np.frombuffer(_encode_cmyk(), np.uint8)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    assert engine._lib.irp_jpeg_info(cmyk.ctypes.data, cmyk.size, C.byref(w), C.byref(h), C.byref(c)) == _ffi.IRP_ERR_UNSUPPORTED


def _encode_cmyk():
    b = io.BytesIO()
    Image.fromarray(rand_image(64, 64, 3, seed=1)).convert("CMYK").save(b, "JPEG", quality=80)
    return b.getvalue()


def test_truncated_and_corrupted_streams_do_not_poison_the_context(engine):
    """Damaged entropy data may decode to anything (or be refused) but must neither crash nor hang, and the
    next clean file must still decode exactly."""
    import irp_b200

    good = _encode(rand_image(300, 420, 3, seed=21, kind="smooth"), quality=85, subsampling=2)
    cut = good[: int(len(good) * 0.6)]
    rng = np.random.default_rng(3)
    noisy = bytearray(good)
    for i in rng.integers(len(good) // 2, len(good) - 4, 40):
        noisy[i] = int(rng.integers(0, 256))
    for bad in (cut, bytes(noisy)):
        try:
            out = engine.decode_jpeg_batch([bad])[0]
            assert out.shape == (300, 420, 3)
        except irp_b200.IrpError:
            pass
    assert np.array_equal(engine.decode_jpeg_batch([good])[0], _pillow(good))


def test_engine_wrappers_and_mixed_batch(engine, oracle):
    blobs = [_encode(rand_image(120, 200, 3, seed=1, kind="smooth"), quality=90, subsampling=2),
             _encode(rand_image(97, 131, 1, seed=2, kind="smooth")[:, :, 0], quality=70),
             _encode(rand_image(2300, 180, 3, seed=3, kind="edges"), quality=80, subsampling=1, restart_marker_rows=2)]
    assert engine.jpeg_info(blobs[0]) == (200, 120, 3) and engine.jpeg_info(blobs[1]) == (131, 97, 1)
    assert engine.jpeg_info(b"not a jpeg at all") is None
    res, outs = engine.analyze_jpeg_batch(blobs, orientations=[1, 3, 4])
    for i, b in enumerate(blobs):
        px = np.ascontiguousarray(_pillow(b))
        assert_result_parity(res[i], oracle.classify(px), 1 if px.ndim == 2 else 3, f"blob {i}")
        ref = oracle.preprocess(px, [1, 3, 4][i])
        assert np.array_equal(outs[i].reshape(ref.shape), ref)
