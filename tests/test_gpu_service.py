"""The reference's own tests, re-run against the drop-in mirrors on the GPU:
server-node/tests/classifierService.test.js:19-57 and tests/middleware.test.js:48-71, with fixtures
re-synthesised as tests/utils/imageFixtures.js builds them (Pillow stands in for sharp's encoder)."""
import asyncio
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _jpeg(a, quality):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


def _flat(color=(180, 180, 180), size=128, quality=95):  # createBaseImage, imageFixtures.js:5-14
    a = np.zeros((size, size, 3), np.uint8)
    a[...] = color
    return _jpeg(a, quality)


@pytest.fixture()
def service(engine):
    from irp_b200 import ClassifierService

    class TestLogger:  # createTestLogger, tests/utils/mocks.js:177-184
        def __init__(self):
            self.calls = []

        def debug(self, *a):
            self.calls.append(("debug", a))

        def warn(self, *a):
            self.calls.append(("warn", a))

        warning = warn

        def error(self, *a):
            self.calls.append(("error", a))

    return ClassifierService(logger=TestLogger(), engine=engine)


def test_detects_motion_blur_in_blurred_images(service):
    from PIL import Image, ImageFilter

    base = Image.open(io.BytesIO(_flat()))
    buf = io.BytesIO()
    base.filter(ImageFilter.GaussianBlur(4)).save(buf, format="JPEG", quality=60)  # createBlurredImage :16-19
    result = asyncio.run(service.analyze(buf.getvalue()))
    assert result["blur"] > 0.2  # classifierService.test.js:23
    assert result["noise"] >= 0
    assert "colorShift" in result
    assert list(result.keys()) == ["blur", "noise", "lowLight", "compression", "scratch", "fade", "colorShift"]


def test_detects_strong_noise_levels(service):
    noise = np.random.default_rng(7).integers(0, 256, (128, 128, 3), dtype=np.uint8)  # createNoisyImage :21-37
    assert asyncio.run(service.analyze(_jpeg(noise, 80)))["noise"] > 0.3  # :32


def test_detects_low_light_conditions(service):
    assert asyncio.run(service.analyze(_flat((10, 10, 10))))["lowLight"] > 0.3  # :39


def test_detects_color_cast_shifts(service):
    assert asyncio.run(service.analyze(_flat((220, 80, 40))))["colorShift"] > 0.25  # :46


def test_returns_normalized_metrics_for_clean_images(service):
    result = asyncio.run(service.analyze(_flat()))
    for v in result.values():  # :53-56
        assert 0 <= v <= 1


def test_service_matches_oracle_on_decoded_pixels(service, oracle):
    from irp_b200.classifier import decode_image

    noise = np.random.default_rng(9).integers(0, 256, (200, 300, 3), dtype=np.uint8)
    buf = _jpeg(noise, 70)
    px, fmt, _ = decode_image(buf)
    got = service.analyze_sync(buf)
    ref = oracle.classify(px, is_jpeg=(fmt == "jpeg"))["scores"]
    for k, v in ref.items():
        assert abs(got[k] - v) <= 1e-4 * max(abs(v), 1e-8)
    assert any(c[0] == "debug" and c[1][0] == "[classifier] Analysis complete" for c in service.logger.calls)


def test_png_input_has_zero_compression_score(service):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(np.random.default_rng(3).integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(buf, format="PNG")
    assert asyncio.run(service.analyze(buf.getvalue()))["compression"] == 0.0  # classifier.js:180-182


def test_preprocess_auto_orients_compresses_and_records_operations(engine):
    from irp_b200.preprocess import make_request, preprocess_image

    buffer = _flat()
    req = make_request(buffer, engine)
    calls = []
    preprocess_image(req, {}, lambda *a: calls.append(a))
    assert calls == [()]  # next() called with no error            middleware.test.js:55
    assert req.file.originalBuffer is not None  # :56
    assert "auto_orient" in req.file.preprocessOperations  # :57
    assert any(op.startswith("compress_jpeg") for op in req.file.preprocessOperations)  # :58
    assert req.file.mimetype == "image/jpeg"  # :59
    assert req.file.buffer != req.file.originalBuffer  # :60
    assert req.file.preprocessOperations == ["auto_orient", "compress_jpeg_q85", "attach_sRGB_icc"]


def test_preprocess_resizes_and_rotates_large_uploads(engine, oracle):
    """Not exercised by the reference's tests: resize + non-trivial EXIF orientation."""
    from PIL import Image

    from irp_b200.preprocess import make_request, preprocess_image

    y, x = np.mgrid[0:2400, 0:3200]
    a = np.stack([(x // 13) % 256, (y // 7) % 256, (x + y) % 256], axis=2).astype(np.uint8)
    im = Image.fromarray(a)
    exif = im.getexif()
    exif[0x0112] = 6
    buf = io.BytesIO()
    im.save(buf, format="PNG", exif=exif.tobytes())
    req = make_request(buf.getvalue(), engine)
    calls = []
    preprocess_image(req, {}, lambda *a: calls.append(a))
    assert calls == [()]
    assert req.file.preprocessOperations[:2] == ["auto_orient", "resize_2048x1536"]
    assert req.file.processedMetadata == {"width": 1152, "height": 1536, "format": "jpeg", "channels": 3}  # the pre-rotation-dims quirk, SURVEY.md §8a P3
    out = Image.open(io.BytesIO(req.file.buffer))
    assert out.format == "JPEG" and out.size == (1152, 1536) and out.info.get("icc_profile") == engine.icc_bytes(engine.SRGB)
    # the file is libjpeg-turbo's optimised q85 4:4:4 file of the oracle's pixels (one call: pixels in, file out)
    ref = io.BytesIO()
    Image.fromarray(oracle.preprocess(a, 6)).save(ref, "JPEG", quality=85, subsampling=0, optimize=True, icc_profile=out.info["icc_profile"])
    assert req.file.buffer == ref.getvalue()


def test_preprocess_of_a_jpeg_upload_never_touches_a_host_codec(engine, oracle):
    """A baseline JPEG upload: decoded on the device, oriented, resized and re-encoded on the device; the returned
    file is libjpeg-turbo's q85 4:4:4 file of the oracle's pixels with the sRGB profile attached."""
    from PIL import Image, ImageCms

    from irp_b200.preprocess import make_request, preprocess_image

    y, x = np.mgrid[0:2500, 0:3300]
    a = np.stack([(x // 11) % 256, (y // 5) % 256, (x + 2 * y) % 256], axis=2).astype(np.uint8)
    im = Image.fromarray(a)
    exif = im.getexif()
    exif[0x0112] = 3
    buf = io.BytesIO()
    im.save(buf, format="JPEG", quality=92, subsampling=1, exif=exif.tobytes())
    req = make_request(buf.getvalue(), engine)
    calls = []
    preprocess_image(req, {}, lambda *a: calls.append(a))
    assert calls == [()]
    assert req.file.originalMetadata["format"] == "jpeg" and req.file.originalMetadata["orientation"] == 3
    px = np.asarray(Image.open(io.BytesIO(buf.getvalue())))
    want_px = oracle.preprocess(px, 3)
    ref = io.BytesIO()
    icc = Image.open(io.BytesIO(req.file.buffer)).info["icc_profile"]
    assert icc == engine.icc_bytes(engine.SRGB) and ImageCms.getProfileDescription(ImageCms.ImageCmsProfile(io.BytesIO(icc))).startswith("sRGB")
    Image.fromarray(want_px).save(ref, "JPEG", quality=85, subsampling=0, optimize=True, icc_profile=icc)
    assert req.file.buffer == ref.getvalue()
