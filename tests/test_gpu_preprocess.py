"""GPU parity, preprocess path (EXIF orient -> lanczos3 fit-inside -> normalise) and fusion canvases,
through the C ABI vs the CPU oracle. Bar: resized pixels within +-1 LSB (orientation-only outputs are
pure permutations and must be bit-exact)."""
import numpy as np
import pytest

from conftest import rand_image

pytestmark = pytest.mark.gpu
TOL = 1  # LSB, BASELINE.json north_star


def _maxdiff(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) if a.size else 0


@pytest.mark.parametrize("orientation", range(1, 9))
def test_orientation_only_is_exact(engine, oracle, orientation):
    img = rand_image(37, 53, 3, seed=orientation)
    out = engine.preprocess_batch([img], orientations=[orientation])[0]
    ref = oracle.preprocess(img, orientation)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("h,w", [(2049, 100), (100, 2049), (2100, 2100), (3000, 2200), (1500, 4000), (2160, 3840)])
def test_resize_parity(engine, oracle, h, w):
    img = rand_image(h, w, 3, seed=h + w, kind="smooth")
    out = engine.preprocess_batch([img])[0]
    ref = oracle.preprocess(img)
    assert _maxdiff(out, ref) <= TOL
    assert np.array_equal(out, ref), "expected bit-exact with the oracle's fixed-point arithmetic"


@pytest.mark.parametrize("orientation", [2, 3, 5, 6, 7, 8])
def test_resize_with_orientation(engine, oracle, orientation):
    img = rand_image(2300, 2500, 3, seed=40 + orientation, kind="noise")
    out = engine.preprocess_batch([img], orientations=[orientation])[0]
    ref = oracle.preprocess(img, orientation)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("c", [1, 4])
def test_resize_other_channels(engine, oracle, c):
    img = rand_image(2500, 2100, c, seed=60 + c, kind="noise")
    out = engine.preprocess_batch([img], orientations=[6])[0]
    ref = oracle.preprocess(img, 6)
    assert out.shape[2] == (1 if c == 1 else 3)
    assert np.array_equal(out, ref)


def test_coefficient_mode_truncate(oracle):
    import irp_b200

    img = rand_image(2400, 2600, 3, seed=70, kind="smooth")
    with irp_b200.Engine(0, coef_mode=1) as eng:
        out = eng.preprocess_batch([img])[0]
    assert np.array_equal(out, oracle.preprocess(img, 1, coef_mode=1))


def test_analyze_batch_matches_separate_calls(engine, oracle):
    imgs = [rand_image(2200, 2600, 3, seed=80, kind="smooth"), rand_image(300, 200, 3, seed=81)]
    res, outs = engine.analyze_batch(imgs, orientations=[1, 8])
    for img, o, r, ori in zip(imgs, outs, res, [1, 8]):
        assert np.array_equal(o, oracle.preprocess(img, ori))
        ref = oracle.classify(img)
        assert r["sum"][:3] == ref["sum"][:3] and r["luma_hist"] == ref["luma_hist"] and r["b_sumsq"] == ref["b_sumsq"]


def test_device_resident_outputs(engine, oracle):
    img = rand_image(2300, 3100, 3, seed=90, kind="smooth")
    ow, oh = engine.preprocess_dims(3100, 2300)
    d_in, d_out = engine.upload(img), engine.alloc_device(ow, oh, 3)
    try:
        engine.preprocess_batch([d_in], device_outputs=[d_out])
        got = engine.download(d_out)
    finally:
        engine.free(d_in)
        engine.free(d_out)
    assert np.array_equal(got, oracle.preprocess(img))


def test_fusion_canvases(engine, oracle):
    a = rand_image(2600, 3400, 3, seed=100, kind="smooth")
    b = rand_image(3000, 2100, 3, seed=101, kind="noise")
    c = rand_image(500, 700, 1, seed=102)
    groups = engine.fusion_prepare_batch([[a, b, c], [b]], orientations=[[1, 6, 1], [3]])
    for got, (img, ori) in zip(groups[0] + groups[1], [(a, 1), (b, 6), (c, 1), (b, 3)]):
        assert got.shape == (2048, 2048, 3)
        assert np.array_equal(got, oracle.fusion_canvas(img, ori))


@pytest.mark.parametrize("h,w,c,o", [(9000, 12000, 3, 1), (300, 24000, 3, 1), (300, 24000, 3, 6), (9001, 9003, 1, 1), (40, 9000, 4, 3)])
def test_shrinks_of_four_and_more_take_the_box_pre_shrink(engine, oracle, h, w, c, o):
    """Sides beyond 8192 px: vips_resize box-averages by floor(shrink / 2) first (shrinkv, shrinkh with "ceil"), then
    lanczos3 does the remaining factor in [2, 4) — imagePreprocess.js:46-53 resizes whatever the upload cap lets in."""
    img = rand_image(h, w, c, seed=h + w + o, kind="smooth" if h * w > 50_000_000 else "noise")
    got = engine.preprocess_batch([img], orientations=[o])[0]
    ref = oracle.preprocess(img, o)
    assert got.shape == ref.shape and max(got.shape[:2]) <= 2048
    assert np.array_equal(got, ref)


def test_switched_modes_follow_the_oracle(oracle):
    """IRP_BLUR_VECTOR and IRP_REDUCE_VECTOR_2_6 (the libvips SIMD-path candidates) change the same bytes on the device
    as in the oracle."""
    import irp_b200
    from conftest import assert_result_parity

    a = rand_image(301, 517, 3, seed=77, kind="noise")
    b = rand_image(2300, 2600, 3, seed=78, kind="smooth")
    with irp_b200.Engine(0, blur_mode=1, reduce_mode=1) as eng:
        res = eng.classify_batch([a])[0]
        out = eng.preprocess_batch([b])[0]
    ref = oracle.classify(a, blur_mode=1)
    assert ref["b_sumsq"] != oracle.classify(a)["b_sumsq"]
    assert_result_parity(res, ref, 3, "blur_mode 1")
    assert np.array_equal(out, oracle.preprocess(b, 1, reduce_mode=1))
    assert not np.array_equal(out, oracle.preprocess(b, 1))
