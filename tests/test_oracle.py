"""Pins the CPU oracle (oracle/irp_oracle.c) — the checker every GPU parity test relies on.

The reference holds NO golden vectors for this path: its tests are thresholds
(server-node/tests/classifierService.test.js:23,32,39,46,53-56).  So the oracle is pinned by
 (a) those thresholds on re-synthesised fixtures (tests/utils/imageFixtures.js),
 (b) the hand-derived known answers of SURVEY.md §8c,
 (c) independent implementations available offline: PIL exif_transpose (P2), scipy correlate (K1),
     a numpy restatement of the blur and of the JS reductions.
Exact-value parity with real sharp/libvips remains UNPINNED and is reported as such.
"""
import io

import numpy as np
import pytest

from conftest import rand_image


def _jpeg_roundtrip(a, quality):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="JPEG", quality=quality)
    return np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))


# ---- (a) the reference's own test thresholds -------------------------------------------------
def test_reference_thresholds_on_resynthesised_fixtures(oracle):
    flat = np.full((128, 128, 3), 180, np.uint8)
    # createBlurredImage: blur(4) of a flat JPEG q95, re-encoded q60 (imageFixtures.js:16-19)
    blurred = _jpeg_roundtrip(_jpeg_roundtrip(flat, 95), 60)
    assert oracle.classify(blurred)["scores"]["blur"] > 0.2  # classifierService.test.js:23
    noisy = _jpeg_roundtrip(rand_image(128, 128, 3, seed=1), 80)  # imageFixtures.js:21-37
    assert oracle.classify(noisy)["scores"]["noise"] > 0.3  # :32
    dark = _jpeg_roundtrip(np.full((128, 128, 3), 10, np.uint8), 95)
    assert oracle.classify(dark)["scores"]["lowLight"] > 0.3  # :39
    cast = np.zeros((128, 128, 3), np.uint8)
    cast[...] = (220, 80, 40)
    assert oracle.classify(_jpeg_roundtrip(cast, 95))["scores"]["colorShift"] > 0.25  # :46
    for v in oracle.classify(_jpeg_roundtrip(flat, 95))["scores"].values():  # :53-56
        assert 0.0 <= v <= 1.0


# ---- (b) hand-derived known answers ------------------------------------------
def test_flat_image_known_answers(oracle):
    s = oracle.classify(np.full((64, 96, 3), 180, np.uint8))["scores"]
    assert s == {"blur": 1.0, "noise": 0.0, "lowLight": 0.0, "compression": 0.0, "scratch": 0.0, "fade": 1.0, "colorShift": 0.0}
    assert oracle.classify(np.full((64, 96, 3), 10, np.uint8))["scores"]["lowLight"] == pytest.approx((0.3 - 10 / 255) * 2, abs=1e-15)
    cast = np.zeros((64, 96, 3), np.uint8)
    cast[...] = (220, 80, 40)
    s = oracle.classify(cast)["scores"]
    assert s["colorShift"] == 1.0 and s["lowLight"] == 0.0


def test_uniform_noise_saturates_noise_score(oracle):
    assert oracle.classify(rand_image(128, 128, 3, seed=3))["scores"]["noise"] == 1.0


def test_grey_ramp_is_identity_and_weights_sum_to_one(oracle):
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)
    assert np.array_equal(oracle.grey(ramp)[0], np.arange(256))
    assert np.array_equal(oracle.grey(ramp, luma_mode=1)[0], np.arange(256))


def test_vertical_step_edge_lap8(oracle):
    g = np.zeros((8, 8), np.uint8)
    g[:, 4:] = 255
    e = oracle.stencil(g, 0)
    assert np.all(e[:, 3] == 0)  # clip(-765) = 0: negative responses vanish (half-rectified)
    assert np.all(e[:, 4] == 255)  # clip(+765) = 255
    assert np.all(e[:, :3] == 0) and np.all(e[:, 5:] == 0)


def test_gaussmat_sigma1_taps():
    import math

    taps = [math.exp(-(x * x) / 2.0) for x in range(0, 3)]
    q = [round(20 * t) for t in taps]  # vips_gaussmat integer: scaled to 20 at the centre
    assert q[:2] == [20, 12] and taps[2] < 0.2  # min_ampl 0.2 cuts the mask at radius 1
    assert 12 + 20 + 12 == 44


def test_blur_div11_identity():
    """(12(l+r) + 20c + 22) / 44 == (3(l+r) + 5c + 5) / 11 == mulhi form, for every reachable value."""
    s = np.arange(0, 511)[:, None]
    c = np.arange(0, 256)[None, :]
    a = (12 * s + 20 * c + 22) // 44
    x = 3 * s + 5 * c + 5
    assert np.array_equal(a, x // 11)
    assert np.array_equal(a, (x.astype(np.uint64) * 390451573) >> 32)
    assert int(x.max()) == 2810


# ---- (c) independent cross-checks ------------------------------------------------------------
@pytest.mark.parametrize("which,kernel", [(0, [-1, -1, -1, -1, 8, -1, -1, -1, -1]), (1, [-1, -1, -1, -1, 9, -1, -1, -1, -1]), (2, [0, -1, 0, -1, 4, -1, 0, -1, 0])])
def test_stencils_vs_scipy(oracle, which, kernel):
    from scipy import ndimage

    g = rand_image(40, 57, 1, seed=which)[:, :, 0]
    ref = ndimage.correlate(g.astype(np.int32), np.array(kernel, np.int32).reshape(3, 3), mode="nearest")
    assert np.array_equal(oracle.stencil(g, which), np.clip(ref, 0, 255).astype(np.uint8))


def test_blur_vs_numpy(oracle):
    a = rand_image(33, 47, 3, seed=8).astype(np.int32)
    p = np.pad(a, ((0, 0), (1, 1), (0, 0)), mode="edge")
    h = (12 * (p[:, :-2] + p[:, 2:]) + 20 * p[:, 1:-1] + 22) // 44
    p = np.pad(h, ((1, 1), (0, 0), (0, 0)), mode="edge")
    v = (12 * (p[:-2] + p[2:]) + 20 * p[1:-1] + 22) // 44
    assert np.array_equal(oracle.blur1(a.astype(np.uint8)), v.astype(np.uint8))


@pytest.mark.parametrize("orientation", range(1, 9))
def test_orientation_vs_pil(oracle, orientation):
    from PIL import Image, ImageOps

    a = rand_image(5, 7, 3, seed=orientation)
    im = Image.fromarray(a)
    exif = im.getexif()
    exif[0x0112] = orientation
    buf = io.BytesIO()
    im.save(buf, format="PNG", exif=exif.tobytes())
    ref = np.asarray(ImageOps.exif_transpose(Image.open(io.BytesIO(buf.getvalue()))))
    assert np.array_equal(oracle.orient(a, orientation), ref)


def test_scores_follow_js_formulas(oracle):
    """Recompute the seven scores in numpy from the oracle's own intermediate buffers."""
    img = rand_image(61, 83, 3, seed=17, kind="smooth")
    r = oracle.classify(img)
    g = oracle.grey(img)
    e1, e2, e3 = (oracle.stencil(g, k).astype(np.float64) for k in range(3))
    assert r["scores"]["blur"] == pytest.approx(max(0, 1 - min(e1.var() / 1000, 1)), rel=1e-12)
    assert r["scores"]["noise"] == pytest.approx(min(e2.std() / 50, 1), rel=1e-12)
    b = oracle.blur1(img).astype(np.float64)
    assert r["scores"]["compression"] == pytest.approx(min(max(0, img.astype(np.float64).var() - b.var()) / 500, 1), rel=1e-9)
    means = img.reshape(-1, 3).mean(0)
    stds = img.reshape(-1, 3).astype(np.float64).std(0, ddof=1)
    assert r["scores"]["fade"] == pytest.approx(min((1 - min(np.sqrt((stds ** 2).sum()) / 255, 1)) * 0.6 + (1 - min(stds.mean() / 64, 1)) * 0.4, 1), rel=1e-12)
    avg = means.mean()
    assert r["scores"]["colorShift"] == pytest.approx(min(2 * np.abs(means - avg).max() / avg, 1), rel=1e-12)
    v = h = 0
    for y in range(0, 61, 4):
        for x in range(0, 83, 4):
            if e3[y, x] > 200:
                v += x + 1 < 83 and e3[y, x + 1] > 200
                h += y + 1 < 61 and e3[y + 1, x] > 200
    assert (r["scratch_v"], r["scratch_h"]) == (v, h)
    assert r["luma_hist"] == np.bincount(g.ravel(), minlength=256).tolist()


def test_resize_dims(oracle):
    assert oracle.preprocess_dims(4000, 3000)[:2] == (2048, 1536)
    assert oracle.preprocess_dims(3840, 2160)[:2] == (2048, 1152)
    assert oracle.preprocess_dims(6000, 4000)[:2] == (2048, 1365)
    assert oracle.preprocess_dims(4000, 3000, 6)[:2] == (1152, 1536)  # the pre-rotation-dims quirk
    assert oracle.preprocess_dims(2048, 2048)[:2] == (2048, 2048)
    assert oracle.preprocess_dims(1024, 768, 5)[:2] == (768, 1024)
    assert oracle.fusion_dims(4000, 3000)[:4] == (2048, 1536, 0, 256)
    assert oracle.fusion_dims(3000, 4000)[:4] == (1536, 2048, 256, 0)
    assert oracle.fusion_dims(640, 480)[:4] == (640, 480, 704, 784)


def test_reduce_plan_properties(oracle):
    for in_size, out_size, shrink in [(3000, 1536, 1.953125), (4000, 1365, 4000 / 1365), (2049, 2048, 2049 / 2048), (8000, 2048, 3.90625)]:
        n, start, phase, coefs = oracle.reduce_plan(in_size, out_size, shrink)
        assert n == 2 * int(np.rint(3 * shrink)) + 1
        assert np.all((phase >= 0) & (phase <= 64))
        assert np.all(np.diff(start) >= 0) and np.all(np.diff(start) <= int(np.ceil(shrink)) + 1)
        assert np.all(coefs[:, :n].sum(axis=1) == 4096)  # vips_vector_to_fixed_point keeps unit DC gain
        assert np.all(coefs[:, n:] == 0)
        n2, _, _, c2 = oracle.reduce_plan(in_size, out_size, shrink, coef_mode=1)
        assert np.all(np.abs(c2[:, :n].astype(int) - coefs[:, :n]) <= 2)


def test_resize_preserves_flat_and_handles_identity(oracle):
    flat = np.full((2300, 2500, 3), 77, np.uint8)
    out = oracle.preprocess(flat)
    assert out.shape == (1884, 2048, 3) and np.all(out == 77)
    small = rand_image(50, 60, 3, seed=2)
    assert np.array_equal(oracle.preprocess(small), small)
    rgba = rand_image(20, 30, 4, seed=4)
    o = oracle.preprocess(rgba)
    assert np.array_equal(o, (rgba[:, :, :3].astype(np.uint32) * rgba[:, :, 3:4] // 255).astype(np.uint8))


def test_resize_downscale_is_close_to_area_average(oracle):
    """Sanity of geometry/centring: a smooth image shrunk by lanczos3 stays close to a box average."""
    y, x = np.mgrid[0:2304, 0:4096].astype(np.float32)
    img = np.repeat((127 + 100 * np.sin(x / 200) * np.cos(y / 150))[:, :, None], 3, 2).astype(np.uint8)
    out = oracle.preprocess(img).astype(np.float32)
    box = img.reshape(1152, 2, 2048, 2, 3).astype(np.float32).mean(axis=(1, 3))
    assert out.shape == (1152, 2048, 3)
    assert np.abs(out - box).max() <= 2.0


def test_batch_threads_agree_with_single(oracle):
    imgs = [rand_image(90 + 10 * i, 120, 3, seed=i) for i in range(5)]
    r1, o1 = oracle.analyze_batch(imgs, threads=1)
    r4, o4 = oracle.analyze_batch(imgs, threads=4)
    assert r1 == r4 and all(np.array_equal(a, b) for a, b in zip(o1, o4))
    assert r1[2] == oracle.classify(imgs[2])


# ---- round 2: the switches for libvips' SIMD-path candidates, and vips_resize's integer box pre-shrink ----
def test_blur_vector_mode_is_the_8_bit_mantissa_mask(oracle):
    """IRP_BLUR_VECTOR: vips_convi_intize bakes {12, 20, 12} / 44 into rint(128 * c) = {35, 58, 35} (sum 128, shared
    exponent 7): (sum + 64) >> 7 per pass, u8 between the passes, replicate edges."""
    assert [round(128 * c / 44) for c in (12, 20, 12)] == [35, 58, 35]
    img = rand_image(37, 53, 3, seed=5, kind="noise").astype(np.int64)
    p = np.pad(img, ((0, 0), (1, 1), (0, 0)), mode="edge")
    hp = (35 * (p[:, :-2] + p[:, 2:]) + 58 * p[:, 1:-1] + 64) >> 7
    q = np.pad(hp, ((1, 1), (0, 0), (0, 0)), mode="edge")
    vp = (35 * (q[:-2] + q[2:]) + 58 * q[1:-1] + 64) >> 7
    assert np.array_equal(oracle.blur1(img.astype(np.uint8), blur_mode=1), vp.astype(np.uint8))
    # the two modes differ by at most one level, and not everywhere
    d = np.abs(oracle.blur1(img.astype(np.uint8), 0).astype(int) - vp)
    assert 0 < d.max() <= 1


def test_reduce_vector_mode_keeps_six_fractional_bits(oracle):
    n, start, phase, coefs = oracle.reduce_plan(4000, 2048, 1.953125, reduce_mode=1)
    assert np.all(coefs % 64 == 0) and np.all(coefs[:, :n].sum(axis=1) == 4096)
    n0, start0, phase0, coefs0 = oracle.reduce_plan(4000, 2048, 1.953125)
    assert n0 == n and np.array_equal(start, start0) and np.array_equal(phase, phase0)
    assert np.abs(coefs.astype(int) - coefs0).max() <= 32 + 64   # quantisation plus the sum-preserving nudge


@pytest.mark.parametrize("h,w,c,kh,kv", [(7, 9, 3, 2, 2), (10, 10, 1, 3, 2), (33, 64, 4, 2, 5)])
def test_box_shrink_is_shrinkv_then_shrinkh_with_ceil(oracle, h, w, c, kh, kv):
    img = rand_image(h, w, c, seed=h * w, kind="noise")
    a = img.astype(np.int64)
    oh, ow = -(-h // kv), -(-w // kh)
    a = np.pad(a, ((0, oh * kv - h), (0, ow * kh - w), (0, 0)), mode="edge")
    v = (a.reshape(oh, kv, ow * kh, c).sum(axis=1) + kv // 2) // kv
    hh = (v.reshape(oh, ow, kh, c).sum(axis=2) + kh // 2) // kh
    assert np.array_equal(oracle.box_shrink(img, kh, kv), hh.astype(np.uint8))


def test_large_shrink_is_box_then_lanczos(oracle):
    """12000 -> 2048: vips_resize pre-shrinks by floor(12000 / 2048 / 2) = 2 and lanczos3 does the remaining 2.93."""
    assert [oracle.box_factor(i, o) for i, o in [(12000, 2048), (8191, 2048), (8192, 2048), (24000, 2048), (4000, 2048)]] == [2, 1, 2, 5, 1]
    img = rand_image(40, 9000, 3, seed=3, kind="smooth")
    ow, oh, f = oracle.preprocess_dims(9000, 40)
    got = oracle.preprocess(img)
    assert got.shape == (oh, ow, 3) and f > 4
    # a flat image stays flat through both stages
    flat = np.full((40, 9000, 3), 77, np.uint8)
    assert np.all(oracle.preprocess(flat) == 77)
