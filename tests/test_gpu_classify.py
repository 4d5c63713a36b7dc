"""GPU parity, classify path: libirp_b200.so (through the C ABI) vs the CPU oracle on the same seeded
inputs. Bar (BASELINE.json north_star): integer statistics, histogram and counts bit-exact; the seven
scores within 1e-4 relative."""
import numpy as np
import pytest

from conftest import assert_result_parity, rand_image

pytestmark = pytest.mark.gpu

SIZES = [
    (1, 1), (1, 7), (7, 1), (2, 2), (3, 5), (8, 8), (15, 17), (16, 16), (31, 33), (32, 256), (33, 257),
    (64, 255), (100, 300), (128, 128), (97, 513), (250, 1000), (1024, 1024),
]


@pytest.mark.parametrize("h,w", SIZES)
@pytest.mark.parametrize("kind", ["noise", "smooth"])
def test_rgb_parity(engine, oracle, h, w, kind):
    img = rand_image(h, w, 3, seed=h * 1000 + w, kind=kind)
    got = engine.classify_batch([img])[0]
    ref = oracle.classify(img)
    assert_result_parity(got, ref, 3, f"{h}x{w} {kind}")


@pytest.mark.parametrize("c", [1, 4])
@pytest.mark.parametrize("h,w", [(1, 1), (5, 9), (33, 257), (128, 128), (300, 700)])
def test_other_channel_counts(engine, oracle, c, h, w):
    img = rand_image(h, w, c, seed=c * 77 + h + w, kind="noise")
    got = engine.classify_batch([img])[0]
    ref = oracle.classify(img)
    assert_result_parity(got, ref, c, f"{h}x{w}x{c}")


def test_edges_pattern_and_scratch_counts(engine, oracle):
    img = rand_image(257, 515, 3, seed=5, kind="edges")
    got = engine.classify_batch([img])[0]
    ref = oracle.classify(img)
    assert ref["scratch_v"] + ref["scratch_h"] > 0
    assert ref["block_edges"][0] > 0 and ref["block_edges"][1] > 0
    assert_result_parity(got, ref, 3, "edges")


def test_non_jpeg_has_zero_compression(engine, oracle):
    img = rand_image(64, 64, 3, seed=9)
    got = engine.classify_batch([img], is_jpeg=False)[0]
    ref = oracle.classify(img, is_jpeg=False)
    assert got["scores"]["compression"] == 0.0
    assert_result_parity(got, ref, 3, "png")


def test_strided_rows_and_unaligned_pitch(engine, oracle):
    base = rand_image(70, 101, 3, seed=11)
    wide = np.zeros((70, 101 * 3 + 7), np.uint8)
    wide[:, : 101 * 3] = base.reshape(70, -1)
    view = wide[:, : 101 * 3].reshape(70, 101, 3)  # pitch 310: not a multiple of 16
    got = engine.classify_batch([view])[0]
    assert_result_parity(got, oracle.classify(base), 3, "pitch")


def test_device_resident_matches_host_path(engine, oracle):
    img = rand_image(300, 420, 3, seed=13, kind="smooth")
    d = engine.upload(img)
    try:
        got = engine.classify_batch([d])[0]
    finally:
        engine.free(d)
    assert_result_parity(got, oracle.classify(img), 3, "device")


def test_mixed_batch_keeps_order(engine, oracle):
    shapes = [(40, 50, 3), (33, 257, 1), (64, 64, 4), (1, 1, 3), (200, 300, 3), (17, 19, 4), (90, 90, 1), (512, 640, 3)]
    imgs = [rand_image(h, w, c, seed=i, kind="noise" if i % 2 else "smooth") for i, (h, w, c) in enumerate(shapes)]
    got = engine.classify_batch(imgs)
    for i, (img, g) in enumerate(zip(imgs, got)):
        assert_result_parity(g, oracle.classify(img), img.shape[2], f"batch[{i}]")


def test_known_answers_flat_images(engine):
    """SURVEY.md §8c hand-derived answers, independent of the uncertain libvips details."""
    flat = np.full((128, 128, 3), 180, np.uint8)
    s = engine.classify_batch([flat])[0]["scores"]
    assert s == {"blur": 1.0, "noise": 0.0, "lowLight": 0.0, "compression": 0.0, "scratch": 0.0, "fade": 1.0, "colorShift": 0.0}
    dark = np.full((128, 128, 3), 10, np.uint8)
    assert abs(engine.classify_batch([dark])[0]["scores"]["lowLight"] - (0.3 - 10 / 255) * 2) < 1e-15
    cast = np.zeros((128, 128, 3), np.uint8)
    cast[...] = (220, 80, 40)
    s = engine.classify_batch([cast])[0]["scores"]
    assert s["colorShift"] == 1.0 and s["lowLight"] == 0.0


def test_linearity_of_moments_under_tiling(engine):
    """Size-independent property: the integer moments of an image tiled 2x2 from itself are 4x the
    single-image channel sums (stencil sums differ only through the seams, so only S0 is checked)."""
    img = rand_image(96, 160, 3, seed=21)
    big = np.tile(img, (2, 2, 1))
    a, b = engine.classify_batch([img, big])
    assert [4 * v for v in a["sum"][:3]] == b["sum"][:3]
    assert [4 * v for v in a["sumsq"][:3]] == b["sumsq"][:3]
    assert [4 * v for v in a["luma_hist"]] == b["luma_hist"]


def test_errors_are_loud(engine):
    import irp_b200

    with pytest.raises(irp_b200.IrpError):
        engine.classify_batch([np.zeros((4, 4, 2), np.uint8)])
