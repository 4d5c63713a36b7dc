"""The JPEG oracle (oracle/jpeg_oracle.c) against libjpeg-turbo itself, through Pillow: bit-exact decoded
pixels for baseline JPEGs of every chroma subsampling, odd sizes, restart intervals, qualities and content.
This is the one PINNED oracle of the repository — Pillow links the same library with the same defaults
(JDCT_ISLOW, fancy upsampling) as the libvips decoder behind the reference's sharp calls
(classifier.js:51-52,107,135,199,296; imagePreprocess.js:40-42)."""
import io

import numpy as np
import pytest
from PIL import Image, ImageFile, features

from conftest import rand_image
from oracle import jpeg_oracle

pytestmark = pytest.mark.skipif(not features.check_feature("libjpeg_turbo"), reason="Pillow without libjpeg-turbo")


ImageFile.MAXBLOCK = 1 << 24   # progressive files of noisy images outgrow Pillow's default encoder buffer


def _encode(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def _pillow(data):
    return np.asarray(Image.open(io.BytesIO(data)))


SHAPES = [(8, 8), (16, 16), (1, 1), (3, 5), (17, 33), (37, 53), (64, 48), (100, 161), (241, 319)]


@pytest.mark.parametrize("h,w", SHAPES)
@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_rgb_matches_libjpeg_turbo(h, w, subsampling):
    for kind, q in (("smooth", 85), ("noise", 60), ("edges", 95)):
        img = rand_image(h, w, 3, seed=h * 7 + w + subsampling, kind=kind)
        data = _encode(img, quality=q, subsampling=subsampling)
        got, ref = jpeg_oracle.decode(data), _pillow(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), f"{h}x{w} subsampling {subsampling} {kind} q{q}: {np.abs(got.astype(int) - ref).max()}"


@pytest.mark.parametrize("h,w", [(8, 8), (5, 9), (37, 53), (130, 70)])
def test_greyscale_matches(h, w):
    img = rand_image(h, w, 1, seed=h + w, kind="smooth")[:, :, 0]
    data = _encode(img, quality=80)
    assert np.array_equal(jpeg_oracle.decode(data), _pillow(data))


@pytest.mark.parametrize("kw", [dict(restart_marker_rows=1), dict(restart_marker_blocks=3), dict(restart_marker_blocks=1), dict(restart_marker_rows=2)])
@pytest.mark.parametrize("subsampling", [0, 2])
def test_restart_intervals(kw, subsampling):
    img = rand_image(75, 131, 3, seed=3, kind="smooth")
    data = _encode(img, quality=75, subsampling=subsampling, **kw)
    assert jpeg_oracle.info(data)["restart_interval"] > 0
    assert np.array_equal(jpeg_oracle.decode(data), _pillow(data))


def test_extreme_content_and_qualities():
    rng = np.random.default_rng(1)
    for q in (1, 10, 50, 100):
        for img in (np.zeros((40, 56, 3), np.uint8), np.full((40, 56, 3), 255, np.uint8), rng.integers(0, 2, (40, 56, 3), dtype=np.uint8) * 255):
            data = _encode(img, quality=q, subsampling=2)
            assert np.array_equal(jpeg_oracle.decode(data), _pillow(data)), f"q{q}"


def test_a_megapixel_photo_like_image():
    img = rand_image(768, 1024, 3, seed=9, kind="smooth")
    data = _encode(img, quality=90, subsampling=2, optimize=True)
    assert np.array_equal(jpeg_oracle.decode(data), _pillow(data))


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_progressive_matches_libjpeg_turbo(subsampling):
    """jdphuff.c restated: DC / AC first passes and refinements, end-of-band runs, correction bits, restart intervals."""
    for i, (h, w) in enumerate([(1, 1), (9, 17), (64, 64), (333, 517), (257, 255)]):
        img = rand_image(h, w, 3, seed=60 + i, kind="smooth" if i % 2 else "noise")
        for q in (30, 85, 95):
            for kw in ({}, {"optimize": True}, {"restart_marker_blocks": 7}):
                data = _encode(img, quality=q, subsampling=subsampling, progressive=True, **kw)
                assert np.array_equal(jpeg_oracle.decode(data), _pillow(data)), f"{h}x{w} q{q} {kw}"
    grey = _encode(rand_image(70, 90, 1, seed=3)[:, :, 0], quality=80, progressive=True)
    assert np.array_equal(jpeg_oracle.decode(grey), _pillow(grey))


@pytest.mark.parametrize("subsampling", [1, 2])
def test_narrow_images_take_the_plain_upsampler(subsampling):
    """jinit_upsampler selects the fancy h2v1 / h2v2 routines only for components more than two samples wide: images of
    four pixels or less across are upsampled by replication."""
    for i, (h, w) in enumerate([(5, 3), (5, 4), (3, 3), (2, 2), (16, 4), (40, 3), (1, 2), (9, 1)]):
        img = rand_image(h, w, 3, seed=80 + i)
        for prog in (False, True):
            data = _encode(img, quality=90, subsampling=subsampling, progressive=prog)
            assert np.array_equal(jpeg_oracle.decode(data), _pillow(data)), f"{h}x{w} progressive={prog}"


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_sequential_file_with_one_scan_per_component(subsampling):
    """SOF0 with three non-interleaved scans (tests/jpeg_rescan.py re-codes a Pillow file's own coefficients that way):
    a scan of one component walks that component's own block grid, not the MCU-padded one."""
    from jpeg_rescan import one_scan_per_component

    for i, (h, w) in enumerate([(37, 53), (64, 64), (100, 161), (9, 17), (8, 8)]):
        data = _encode(rand_image(h, w, 3, seed=70 + i, kind="smooth"), quality=85, subsampling=subsampling)
        multi = one_scan_per_component(data)
        assert np.array_equal(_pillow(multi), _pillow(data))            # libjpeg-turbo reads the rewritten file as the same picture
        assert np.array_equal(jpeg_oracle.decode(multi), _pillow(data))


@pytest.mark.parametrize("name", ["spectral selection only", "dc per component, deep approximation", "fine bands refined one by one"])
def test_progressive_scan_scripts_beyond_the_standard_one(name):
    """Progressive files with scan scripts libjpeg's standard progression does not produce (what mozjpeg's scan search
    may write): DC scans per component, three approximation levels, bands split finely and refined one by one.
    tests/jpeg_rescan.py re-codes a baseline file's own coefficients under the script; libjpeg-turbo and the restatement
    must both give the baseline file's pixels back."""
    from jpeg_rescan import SCRIPTS, progressive_with_script

    for i, (h, w) in enumerate([(37, 53), (64, 64), (100, 161), (9, 17)]):
        for sub in (0, 1, 2):
            data = _encode(rand_image(h, w, 3, seed=40 + i, kind="smooth" if (i + sub) % 2 else "noise"), quality=90, subsampling=sub)
            f = progressive_with_script(data, SCRIPTS[name])
            assert np.array_equal(_pillow(f), _pillow(data))
            assert np.array_equal(jpeg_oracle.decode(f), _pillow(data))


def test_random_progressive_scan_scripts():
    """Random valid progressions (bands, approximation depths and scan order drawn at random, jpeg_rescan.random_script)."""
    from jpeg_rescan import progressive_with_script, random_script

    rng = np.random.default_rng(11)
    for t in range(16):
        h, w, sub = int(rng.integers(8, 90)), int(rng.integers(8, 120)), int(rng.integers(0, 3))
        data = _encode(rand_image(h, w, 3, seed=t, kind="noise" if t % 2 else "smooth"), quality=int(rng.integers(60, 96)), subsampling=sub)
        f = progressive_with_script(data, random_script(rng))
        assert np.array_equal(_pillow(f), _pillow(data))
        assert np.array_equal(jpeg_oracle.decode(f), _pillow(data)), f"case {t}"


def test_four_components_are_reported_unsupported():
    import io

    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(rand_image(32, 32, 3, seed=1)).convert("CMYK").save(b, "JPEG", quality=80)
    with pytest.raises(ValueError):
        jpeg_oracle.decode(b.getvalue())


def test_coefficients_hook_shape():
    data = _encode(rand_image(37, 53, 3, seed=2, kind="smooth"), quality=85, subsampling=2)
    y, cb = jpeg_oracle.coefficients(data, 0), jpeg_oracle.coefficients(data, 1)
    assert y.shape == (6, 8, 64) and cb.shape == (3, 4, 64) and y.any()


def _pillow_scaled(data, denom, mode="RGB"):
    """libjpeg-turbo's own decode at scale 1 / denom (Pillow's draft mode sets scale_denom and keeps JDCT_ISLOW and
    fancy upsampling); None when Pillow settles for another scale."""
    im = Image.open(io.BytesIO(data))
    w, h = im.size
    im.draft(mode, (max(1, w // denom), max(1, h // denom)))
    a = np.asarray(im)
    return a if a.shape[:2] == (-(-h // denom), -(-w // denom)) else None


@pytest.mark.parametrize("denom", [2, 4, 8])
def test_reduced_size_decode_matches_libjpeg_turbo(denom):
    """Shrink-on-load (libvips asks libjpeg-turbo for scale 1/2, 1/4, 1/8): jidctred.c's 4x4 / 2x2 / 1x1 inverse DCTs,
    per-component IDCT sizes (a 2x subsampled chroma component is scaled up by a larger IDCT instead of the upsampler),
    fancy upsampling only while the smallest IDCT is larger than 1x1 — baseline and progressive files, every sampling."""
    n = 0
    for i, (h, w) in enumerate([(64, 64), (333, 517), (17, 9), (100, 161), (255, 257), (8, 8), (5, 3), (31, 33), (16, 4), (9, 20)]):
        img = rand_image(h, w, 3, seed=90 + i, kind="smooth" if i % 2 else "noise")
        for sub in (0, 1, 2):
            for prog in (False, True):
                data = _encode(img, quality=85, subsampling=sub, progressive=prog)
                ref = _pillow_scaled(data, denom)
                if ref is None:
                    continue
                assert np.array_equal(jpeg_oracle.decode_scaled(data, denom), ref), f"{h}x{w} sub {sub} progressive {prog}"
                n += 1
        grey = _encode(img[:, :, 0], quality=85)
        ref = _pillow_scaled(grey, denom, "L")
        if ref is not None:
            assert np.array_equal(jpeg_oracle.decode_scaled(grey, denom), ref)
            n += 1
    assert n >= 40
