"""BASELINE.json's full sizes on the GPU: a 12 MP and a 4K image against the oracle (a few seconds of
CPU each), then size-independent properties on a whole 12 MP batch where the oracle would be too slow:
batch-order independence, duplicate images give identical results, channel sums equal numpy's, the
histogram sums to the pixel count, resized flat images stay flat, and a checksum over the batch."""
import hashlib

import numpy as np
import pytest

from conftest import assert_result_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,idx", [(4000, 3000, 0), (3840, 2160, 1), (1024, 1024, 2)])
def test_configs_against_oracle(engine, oracle, w, h, idx):
    from irp_b200.synth import synth_image

    img = synth_image(w, h, idx)
    res, outs = engine.analyze_batch([img])
    assert_result_parity(res[0], oracle.classify(img), 3, f"{w}x{h}")
    ref = oracle.preprocess(img)
    assert outs[0].shape == ref.shape
    assert int(np.abs(outs[0].astype(np.int16) - ref.astype(np.int16)).max()) <= 1  # north_star: +-1 LSB
    assert np.array_equal(outs[0], ref)


def test_24mp_against_oracle(engine, oracle):
    """The top of the mixed-resolution range (BASELINE.json configs[4]): classify AND preprocess of a 6000x4000 image."""
    from irp_b200.synth import synth_image

    img = synth_image(6000, 4000, 3)
    res, outs = engine.analyze_batch([img])
    assert_result_parity(res[0], oracle.classify(img), 3, "24 MP")
    assert outs[0].shape == (1365, 2048, 3)
    assert np.array_equal(outs[0], oracle.preprocess(img))


def test_12mp_batch_properties(engine):
    from irp_b200.synth import synth_batch

    imgs = synth_batch(4000, 3000, 6, distinct=3)  # images 3..5 are row-rolled copies of 0..2
    order = [4, 0, 5, 1, 3, 2, 0]  # shuffled, with a duplicate of image 0
    res, outs = engine.analyze_batch([imgs[i] for i in order])
    res_sorted, _ = engine.analyze_batch(imgs)
    for pos, i in enumerate(order):  # batch position must not matter
        assert res[pos] == res_sorted[i]
    assert res[1] == res[6] and np.array_equal(outs[1], outs[6])  # duplicates agree bit for bit
    for i, r in enumerate(res_sorted):
        a = imgs[i].reshape(-1, 3).astype(np.uint64)
        assert r["sum"][:3] == a.sum(0).tolist()
        assert r["sumsq"][:3] == (a * a).sum(0).tolist()
        assert sum(r["luma_hist"]) == 4000 * 3000
        for v in r["scores"].values():
            assert 0.0 <= v <= 1.0
    # a vertical roll keeps every per-pixel statistic that ignores row adjacency at the seam
    assert res_sorted[0]["sum"] == res_sorted[3]["sum"] and res_sorted[0]["luma_hist"] == res_sorted[3]["luma_hist"]
    digest = hashlib.sha256(b"".join(o.tobytes() for o in outs)).hexdigest()
    res2, outs2 = engine.analyze_batch([imgs[i] for i in order])
    assert hashlib.sha256(b"".join(o.tobytes() for o in outs2)).hexdigest() == digest  # deterministic


def test_flat_12mp_resizes_to_flat(engine):
    flat = np.full((3000, 4000, 3), 201, np.uint8)
    res, outs = engine.analyze_batch([flat])
    assert outs[0].shape == (1536, 2048, 3) and np.all(outs[0] == 201)
    assert res[0]["scores"] == {"blur": 1.0, "noise": 0.0, "lowLight": 0.0, "compression": 0.0, "scratch": 0.0, "fade": 1.0, "colorShift": 0.0}


def test_fusion_triplet_of_12mp(engine, oracle):
    from irp_b200.synth import synth_image

    trip = [synth_image(4000, 3000, 10), synth_image(3000, 4000, 11), synth_image(3840, 2160, 12)]
    canv = engine.fusion_prepare_batch([trip])[0]
    for k in range(3):   # every aspect of the triplet
        assert np.array_equal(canv[k], oracle.fusion_canvas(trip[k])), f"canvas {k}"
    for c, (ow, oh) in zip(canv, [(2048, 1536), (1536, 2048), (2048, 1152)]):
        ox, oy = (2048 - ow) // 2, (2048 - oh) // 2
        mask = np.ones((2048, 2048), bool)
        mask[oy : oy + oh, ox : ox + ow] = False
        assert not c[mask].any()  # black pad outside the centred image
        assert c[~mask].any()


def test_mixed_resolution_queue_slice(engine, oracle):
    """BASELINE.json configs[4], a slice: mixed sizes/aspects in one submission, checked per image."""
    from irp_b200.synth import mixed_resolution_sizes, synth_image

    every = mixed_resolution_sizes(64)
    sizes = [s for s in every if s[0] * s[1] < 3e6][:5] + sorted(every, key=lambda s: -s[0] * s[1])[:2]   # ... and the two largest (>= 12 MP)
    assert sizes[-1][0] * sizes[-1][1] >= 12e6
    imgs = [synth_image(w, h, 40 + i) for i, (w, h) in enumerate(sizes)]
    res, outs = engine.analyze_batch(imgs)
    for img, r, o in zip(imgs, res, outs):
        assert_result_parity(r, oracle.classify(img), 3, str(img.shape))
        assert np.array_equal(o, oracle.preprocess(img))
