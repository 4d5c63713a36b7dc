"""Damaged JPEG files (sharp is opened with failOnError: false, imagePreprocess.js:42): truncated scans, flipped
bytes in the entropy-coded data, stray markers.  The device decoder may refuse a file or return a picture with
damage in it, but it must not fault, hang or poison the context — a valid file decoded afterwards is still
bit-exact — and the part of a truncated picture that precedes the damage is still libjpeg-turbo's."""
import io

import numpy as np
import pytest
from PIL import Image

import irp_b200
from conftest import rand_image

pytestmark = pytest.mark.gpu


def _encode(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def _scan_start(data):
    p = 2
    while True:
        m = data[p + 1]
        ln = (data[p + 2] << 8) | data[p + 3]
        if m == 0xDA:
            return p + 2 + ln
        p += 2 + ln


def _try_decode(engine, blob, shape):
    try:
        out = engine.decode_jpeg_batch([blob])[0]
    except irp_b200.IrpError:
        return None
    assert out.shape == shape
    return out


@pytest.mark.parametrize("subsampling", [0, 2])
def test_damaged_files_neither_fault_nor_poison_the_context(engine, subsampling):
    rng = np.random.default_rng(100 + subsampling)
    img = rand_image(480, 640, 3, seed=21, kind="smooth")
    good = _encode(img, quality=85, subsampling=subsampling)
    ref = np.asarray(Image.open(io.BytesIO(good)))
    s0 = _scan_start(good)
    variants = []
    for cut in sorted(rng.integers(s0 + 1, len(good) - 2, 12)):          # truncated inside the scan
        variants.append(good[:int(cut)])
    for k in (1, 3, 17, 200):                                            # flipped bytes
        b = bytearray(good)
        for pos in rng.integers(s0, len(good) - 2, k):
            b[int(pos)] ^= int(rng.integers(1, 256))
        variants.append(bytes(b))
    b = bytearray(good)                                                  # a stray restart marker and a stray 0xFF 0x00
    mid = (s0 + len(good)) // 2
    b[mid:mid] = b"\xff\xd3"
    variants.append(bytes(b))
    variants.append(good[:-2])                                           # no EOI
    variants.append(good[:s0])                                           # no entropy-coded data at all
    survived = 0
    for v in variants:
        out = _try_decode(engine, v, ref.shape)
        survived += out is not None
    assert survived >= 1
    # truncation: the MCU rows that were complete before the cut are still exact
    cut = s0 + (len(good) - s0) // 2
    out = _try_decode(engine, good[:cut], ref.shape)
    if out is not None:
        rows_equal = np.all(out == ref, axis=(1, 2))
        assert rows_equal[:64].all(), "the first MCU rows of a file cut half way must be intact"
    # the context still works and is still exact
    assert np.array_equal(engine.decode_jpeg_batch([good])[0], ref)
    # a batch holding a damaged file and a good one: the good one is still exact if the batch is accepted
    try:
        outs = engine.decode_jpeg_batch([variants[0], good])
        assert np.array_equal(outs[1], ref)
    except irp_b200.IrpError:
        pass
    assert np.array_equal(engine.decode_jpeg_batch([good])[0], ref)


def test_header_over_the_pixel_limit_is_refused(engine):
    good = bytearray(_encode(rand_image(16, 16, 3, seed=1, kind="smooth"), quality=85))
    p = 2
    while good[p + 1] != 0xC0:
        p += 2 + ((good[p + 2] << 8) | good[p + 3])
    good[p + 5:p + 9] = b"\xff\xff\xff\xff"      # 65535 x 65535
    assert engine.jpeg_info(bytes(good)) is None


@pytest.mark.parametrize("restart", [0, 4])
def test_restart_marker_flood_is_refused_without_overrunning_the_upload_slot(engine, restart):
    """every RSTn pads the un-stuffed stream up to 4 bytes, so a scan made of 'X FF Dn' grows 3
    input bytes into 4 output bytes.  With more markers than restart intervals the file is rejected (with DRI) or
    any RSTn is (without DRI) — and the files sharing its batch and the context stay intact."""
    img = rand_image(64, 64, 3, seed=5, kind="smooth")
    good = _encode(img, quality=85, subsampling=0, restart_marker_blocks=restart) if restart else _encode(img, quality=85, subsampling=0)
    ref = np.asarray(Image.open(io.BytesIO(good)))
    s0 = _scan_start(good)
    flood = b"".join(bytes([0x55, 0xFF, 0xD0 + (k & 7)]) for k in range(100000))
    bad = good[:s0] + flood + b"\xff\xd9"
    with pytest.raises(irp_b200.IrpError) as e:
        engine.decode_jpeg_batch([bad])
    assert "restart" in str(e.value)
    neighbour = _encode(rand_image(200, 312, 3, seed=6, kind="smooth"), quality=90, subsampling=2)
    with pytest.raises(irp_b200.IrpError):
        engine.decode_jpeg_batch([neighbour, bad, good])
    outs = engine.decode_jpeg_batch([neighbour, good])
    assert np.array_equal(outs[0], np.asarray(Image.open(io.BytesIO(neighbour)))) and np.array_equal(outs[1], ref)


def _markers(data):
    """(offset, marker, segment length) of every marker segment up to EOI, entropy-coded data skipped."""
    out, p = [], 2
    while p + 4 <= len(data):
        if data[p] != 0xFF:
            p += 1
            continue
        m = data[p + 1]
        if m in (0x00, 0xFF) or 0xD0 <= m <= 0xD7:
            p += 2 if m != 0xFF else 1
            continue
        if m == 0xD9:
            break
        ln = (data[p + 2] << 8) | data[p + 3]
        out.append((p, m, ln))
        p += 2 + ln
    return out


@pytest.mark.parametrize("subsampling", [0, 2])
def test_damaged_progressive_files_neither_fault_nor_poison_the_context(engine, subsampling):
    """The same for multi-scan files: cuts inside any scan, flipped bytes in the entropy-coded data of every scan,
    scan headers with out-of-range parameters, a missing scan.  prog_scan_kernel's loops are bounded by the frame
    geometry, so damaged data can only produce wrong coefficients, never a hang or an out-of-range store."""
    rng = np.random.default_rng(300 + subsampling)
    img = rand_image(320, 480, 3, seed=23, kind="smooth")
    good = _encode(img, quality=85, subsampling=subsampling, progressive=True)
    ref = np.asarray(Image.open(io.BytesIO(good)))
    segs = _markers(good)
    sos = [(p, ln) for p, m, ln in segs if m == 0xDA]
    assert len(sos) >= 6
    variants = []
    for cut in sorted(rng.integers(sos[0][0] + 20, len(good) - 2, 16)):   # truncated inside some scan
        variants.append(good[:int(cut)])
    for k in (1, 5, 50, 400):                                            # flipped bytes anywhere after the first SOS
        b = bytearray(good)
        for pos in rng.integers(sos[0][0] + sos[0][1] + 2, len(good) - 2, k):
            b[int(pos)] ^= int(rng.integers(1, 256))
        variants.append(bytes(b))
    for (p, ln), (off, val) in zip(sos, [(-3, 70), (-2, 99), (-1, 0xFE), (-3, 0), (2, 9), (-2, 0)]):   # scan parameters
        b = bytearray(good)
        b[p + 2 + ln + off if off < 0 else p + 4 + off] = val
        variants.append(bytes(b))
    p, ln = sos[3]
    nxt = [q for q, m, _ in segs if q > p][0]
    variants.append(good[:p] + good[nxt:])                               # one scan missing altogether
    for v in variants:
        _try_decode(engine, v, ref.shape)
    # the complete scans of a file cut inside its LAST scan were all applied: the picture is close to the full one
    last = sos[-1][0]
    out = _try_decode(engine, good[: last + (len(good) - last) // 2], ref.shape)
    if out is not None:
        assert np.abs(out.astype(int) - ref).mean() < 12.0
        assert np.all(out[:32] == ref[:32]), "rows whose last refinement precedes the cut must be exact"
    assert np.array_equal(engine.decode_jpeg_batch([good])[0], ref)
    try:
        outs = engine.decode_jpeg_batch([variants[0], good])
        assert np.array_equal(outs[1], ref)
    except irp_b200.IrpError:
        pass
