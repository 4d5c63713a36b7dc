"""GPU parity of BOTH data paths of each kernel family against the CPU oracle.

3-channel images whose rows are 16-byte aligned take the streaming kernels (TMA tile loads:
classify_bulk_kernel, resize_tma_kernel); everything else — and every image when IRP_NO_BULK=1 /
IRP_NO_RTMA=1 is set before the context is created — takes the generic kernels (classify_kernel<C>,
resize_kernel<C>).  Both must give the oracle's answers bit for bit, on shapes that sit on and around
the tile boundaries (128x32 source tiles, 64x32 output tiles), on batches that change image inside a
persistent group's tile sequence, and on sources that need replicate rows / columns on every side."""
import os

import numpy as np
import pytest

from conftest import assert_result_parity, rand_image

pytestmark = pytest.mark.gpu

EDGE_SIZES = [(1, 1), (2, 3), (31, 127), (32, 128), (33, 129), (34, 130), (63, 255), (64, 256), (65, 257), (96, 384),
              (97, 385), (40, 640), (200, 131), (513, 97)]


@pytest.fixture(scope="module")
def generic_engine():
    """A second context that is forced onto the generic kernels."""
    import irp_b200

    old = {k: os.environ.get(k) for k in ("IRP_NO_BULK", "IRP_NO_RTMA")}
    os.environ["IRP_NO_BULK"] = "1"
    os.environ["IRP_NO_RTMA"] = "1"
    try:
        eng = irp_b200.Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def tma_engine():
    """A third context without the tensor-core resize: 3-channel aligned images take resize_tma_kernel (the ALU
    formulation, still the path for footprints the tensor-core kernel's shared memory cannot hold)."""
    import irp_b200

    old = os.environ.get("IRP_NO_RMMA")
    os.environ["IRP_NO_RMMA"] = "1"
    try:
        eng = irp_b200.Engine(0)
    finally:
        if old is None:
            os.environ.pop("IRP_NO_RMMA", None)
        else:
            os.environ["IRP_NO_RMMA"] = old
    yield eng
    eng.close()


@pytest.mark.parametrize("h,w", EDGE_SIZES)
def test_classify_streaming_and_generic_agree_with_oracle(engine, generic_engine, oracle, h, w):
    img = rand_image(h, w, 3, seed=7 * h + w, kind="smooth")
    ref = oracle.classify(img)
    assert_result_parity(engine.classify_batch([img])[0], ref, 3, f"streaming {h}x{w}")
    assert_result_parity(generic_engine.classify_batch([img])[0], ref, 3, f"generic {h}x{w}")


def test_classify_batch_of_many_small_images_flushes_per_image(engine, oracle):
    """More images than persistent groups see tiles: accumulators must be flushed at every image change."""
    imgs = [rand_image(40 + 3 * (i % 5), 150 + 11 * (i % 7), 3, seed=i, kind="noise" if i % 2 else "edges") for i in range(40)]
    got = engine.classify_batch(imgs)
    for i, im in enumerate(imgs):
        assert_result_parity(got[i], oracle.classify(im), 3, f"image {i}")


def test_classify_long_image_forces_mid_image_flushes(engine, oracle):
    """One image much larger than 16 tiles per group: the u32 accumulators are flushed inside the image."""
    img = rand_image(2600, 1500, 3, seed=99, kind="noise")
    assert_result_parity(engine.classify_batch([img])[0], oracle.classify(img), 3, "2600x1500")


@pytest.mark.parametrize("h,w", [(2049, 64), (64, 2049), (2050, 2050), (2500, 3100), (3100, 2500), (2200, 4099), (4099, 2200),
                                 (8000, 2100), (2100, 8000)])
def test_resize_streaming_and_generic_agree_with_oracle(engine, generic_engine, tma_engine, oracle, h, w):
    img = rand_image(h, w, 3, seed=h + 3 * w, kind="smooth")
    ref = oracle.preprocess(img, 1)
    a = engine.preprocess_batch([img])[0]
    b = generic_engine.preprocess_batch([img])[0]
    c = tma_engine.preprocess_batch([img])[0]
    assert a.shape == ref.shape and np.array_equal(a, ref), f"tensor-core / streaming {h}x{w}"
    assert b.shape == ref.shape and np.array_equal(b, ref), f"generic {h}x{w}"
    assert c.shape == ref.shape and np.array_equal(c, ref), f"streaming (ALU) {h}x{w}"


def test_plan_cache_is_bounded(tmp_path):
    """Geometry plans are cached per context; past IRP_PLAN_CACHE_MB the cache is flushed whole.  With a 1 MB cap every
    call after the first few rebuilds its plans, and the pixels must not change."""
    import subprocess
    import sys

    script = tmp_path / "cap.py"
    script.write_text(
        "import sys; sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + '/tests')\n"
        "import numpy as np, irp_b200\nfrom conftest import rand_image\nfrom oracle import oracle\n"
        "with irp_b200.Engine(0) as eng:\n"
        "    for k in range(12):\n"
        "        img = rand_image(2100 + 37 * k, 2300 + 53 * (k % 5), 3, seed=k, kind='smooth')\n"
        "        assert np.array_equal(eng.preprocess_batch([img])[0], oracle.preprocess(img, 1)), k\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, env=dict(os.environ, IRP_PLAN_CACHE_MB="1"), timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout + p.stderr


def test_resize_mixed_geometries_in_one_launch(engine, oracle):
    """Jobs with different shrink factors share one launch: the tile box is the launch-wide maximum."""
    shapes = [(2100, 2300), (3000, 4000), (2049, 2049), (6000, 2500), (2500, 7000)]
    imgs = [rand_image(h, w, 3, seed=i, kind="smooth") for i, (h, w) in enumerate(shapes)]
    outs = engine.preprocess_batch(imgs, orientations=[1, 6, 3, 4, 2])
    for i, (im, o) in enumerate(zip(imgs, [1, 6, 3, 4, 2])):
        assert np.array_equal(outs[i], oracle.preprocess(im, o)), f"image {i} {shapes[i]} orientation {o}"


def test_unaligned_device_rows_take_the_generic_path(engine, oracle):
    """A tight device-resident image whose row length is not a multiple of 16 cannot be described to
    the TMA unit; it must silently take the generic kernels and still match."""
    img = rand_image(2101, 2101, 3, seed=5, kind="smooth")   # 6303-byte rows
    d = engine.upload(img, pitch_align=1)
    res, outs = engine.analyze_batch([d])
    assert_result_parity(res[0], oracle.classify(img), 3, "device 2101x2101")
    assert np.array_equal(outs[0], oracle.preprocess(img, 1))
