"""The sharp pin.  tools/sharp_golden.mjs (run by a maintainer who has Node: this image does not) dumps what the
reference's own calls produce on the seeded fixtures of tests/golden/sharp_cases.py; this file compares the CPU oracle
with that dump and says, switch by switch (luma_mode, blur_mode, coef_mode, reduce_mode — every libvips detail SURVEY.md
section 8a tags MED or LOW), which setting reproduces sharp.  Without tests/golden/sharp_golden.json the comparison is
skipped with that reason; the harness itself is always tested, on a dump fabricated from the oracle under NON-default
switches, which it has to identify.

Bars (BASELINE.json north_star): grey / stencil / blur buffers and counts bit-exact, scores 1e-4 relative, resized
pixels within 1 LSB."""
import base64
import hashlib
import itertools
import json
import os
import sys

import numpy as np
import pytest

from conftest import rand_image, rel_close

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from sharp_cases import ALL_CASES, CLASSIFY_CASES, PREPROCESS_CASES  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "sharp_golden.json")
CASES = {c[0]: c for c in ALL_CASES}
STENCILS = (("lap8", 0), ("sharp9", 1), ("lap4", 2))


def pixels_of(name):
    _, h, w, c, seed, kind, orientation = CASES[name]
    return rand_image(h, w, c, seed, kind), orientation


def pack(buf: np.ndarray, keep: bool = True) -> dict:
    b = np.ascontiguousarray(buf, np.uint8).ravel()
    d = {"n": int(b.size), "sum": str(int(b.astype(np.uint64).sum())), "sumsq": str(int((b.astype(np.uint64) ** 2).sum())),
         "sha256": hashlib.sha256(b.tobytes()).hexdigest()}
    if keep:
        d["base64"] = base64.b64encode(b.tobytes()).decode()
    return d


def differing(rec: dict, mine: np.ndarray) -> int:
    """Bytes of `mine` that differ from the dumped buffer (exact count when the dump carries it, else 0 / size from the
    hash and the moments)."""
    m = np.ascontiguousarray(mine, np.uint8).ravel()
    if rec["n"] != m.size:
        return max(rec["n"], m.size)
    if "base64" in rec:
        return int((np.frombuffer(base64.b64decode(rec["base64"]), np.uint8) != m).sum())
    return 0 if rec["sha256"] == hashlib.sha256(m.tobytes()).hexdigest() else m.size


def analyse(golden: dict, oracle) -> dict:
    """{switch: {setting: total differing bytes over the cases that exercise it}} + per-case detail."""
    rep = {"luma_mode": {0: 0, 1: 0}, "blur_mode": {0: 0, 1: 0}, "resize": {k: 0 for k in itertools.product((0, 1), (0, 1))},
           "resize_max_abs": {k: 0 for k in itertools.product((0, 1), (0, 1))}, "stencils": 0, "detail": []}
    for rec in golden["cases"]:
        img, orientation = pixels_of(rec["name"])
        assert hashlib.sha256(img.tobytes()).hexdigest() == rec["pixels_sha256"], f"{rec['name']}: fixture pixels changed"
        c = img.shape[2]
        if rec["group"] == "classify":
            if c >= 3:
                for lm in (0, 1):
                    rep["luma_mode"][lm] += differing(rec["grey"], oracle.grey(img, lm))
            if c != 4:   # sharp premultiplies RGBA around blur / convolve: the documented deviation, not a switch
                for bm in (0, 1):
                    rep["blur_mode"][bm] += differing(rec["blurred"], oracle.blur1(img, bm))
            # the stencils on sharp's OWN grey: K1 / K2 are judged apart from the grey conversion
            if "base64" in rec["grey"] and c != 4:
                g = np.frombuffer(base64.b64decode(rec["grey"]["base64"]), np.uint8).reshape(img.shape[:2])
                for key, which in STENCILS:
                    d = differing(rec[key], oracle.stencil(g, which))
                    rep["stencils"] += d
                    if d:
                        rep["detail"].append(f"{rec['name']}: {key} differs in {d} bytes on sharp's own grey")
        else:
            want = rec["resized"]
            shape = (want["info"]["height"], want["info"]["width"], want["info"]["channels"])
            ref = np.frombuffer(base64.b64decode(want["base64"]), np.uint8).reshape(shape)
            if c == 4:
                continue   # premultiplied float resize in sharp: reported by test_sharp_golden, not part of the switch vote
            for cm, rm in rep["resize"]:
                mine = oracle.preprocess(img, orientation, coef_mode=cm, reduce_mode=rm)
                if mine.shape != ref.shape:
                    rep["resize"][(cm, rm)] += ref.size
                    rep["detail"].append(f"{rec['name']}: output {mine.shape} vs sharp {ref.shape}")
                    continue
                d = np.abs(mine.astype(int) - ref.astype(int))
                rep["resize"][(cm, rm)] += int((d > 0).sum())
                rep["resize_max_abs"][(cm, rm)] = max(rep["resize_max_abs"][(cm, rm)], int(d.max()))
    rep["best"] = {"luma_mode": min(rep["luma_mode"], key=rep["luma_mode"].get), "blur_mode": min(rep["blur_mode"], key=rep["blur_mode"].get)}
    rep["best"]["coef_mode"], rep["best"]["reduce_mode"] = min(rep["resize"], key=rep["resize"].get)
    return rep


def fabricate(oracle, luma_mode, blur_mode, coef_mode, reduce_mode, names) -> dict:
    """A dump in sharp_golden.mjs's format, made by the oracle under the given switches."""
    out = {"versions": {"fabricated": True}, "cases": []}
    for name in names:
        img, orientation = pixels_of(name)
        h, w, c = img.shape
        rec = {"name": name, "group": "classify" if name.startswith("c_") else "preprocess", "width": w, "height": h, "channels": c,
               "orientation": orientation, "pixels_sha256": hashlib.sha256(img.tobytes()).hexdigest()}
        if rec["group"] == "classify":
            g = oracle.grey(img, luma_mode)
            rec["grey"] = pack(g)
            for key, which in STENCILS:
                rec[key] = pack(oracle.stencil(g, which))
            rec["blurred"] = pack(oracle.blur1(img, blur_mode))
            r = oracle.classify(img, is_jpeg=False, luma_mode=luma_mode, blur_mode=blur_mode)
            rec["analysis"] = r["scores"]
            rec["blockiness"] = oracle.classify(img, is_jpeg=True, luma_mode=luma_mode, blur_mode=blur_mode)["scores"]["compression"]
        else:
            o = oracle.preprocess(img, orientation, coef_mode=coef_mode, reduce_mode=reduce_mode)
            rec["resized"] = {"info": {"width": o.shape[1], "height": o.shape[0], "channels": o.shape[2]}, **pack(o)}
        out["cases"].append(rec)
    return out


@pytest.mark.parametrize("combo", [(1, 1, 0, 1), (0, 0, 1, 0), (1, 0, 0, 0)])
def test_harness_identifies_the_switches(oracle, combo):
    names = ["c_33x257_smooth", "c_128_noise", "c_120x200_grey", "p_2600x32_noise", "p_32x2600_smooth", "p_2600x48_rot6"]
    rep = analyse(fabricate(oracle, *combo, names), oracle)
    lm, bm, cm, rm = combo
    assert rep["best"] == {"luma_mode": lm, "blur_mode": bm, "coef_mode": cm, "reduce_mode": rm}, rep
    assert rep["luma_mode"][lm] == 0 and rep["blur_mode"][bm] == 0 and rep["resize"][(cm, rm)] == 0 and rep["stencils"] == 0
    assert rep["luma_mode"][1 - lm] > 0 and rep["blur_mode"][1 - bm] > 0


def test_fixture_cases_are_what_the_writer_writes():
    """tools/sharp_fixtures.py and this file must agree on the pixels (same case table, same generator)."""
    assert len({c[0] for c in ALL_CASES}) == len(ALL_CASES) == len(CLASSIFY_CASES) + len(PREPROCESS_CASES)
    for name, h, w, c, *_ in ALL_CASES:
        img, _ = pixels_of(name)
        assert img.shape == (h, w, c) and img.dtype == np.uint8
    assert all(max(h, w) > 2048 or name == "p_300x200_keep" for name, h, w, *_ in PREPROCESS_CASES)


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="tests/golden/sharp_golden.json absent: needs Node + the reference's sharp "
                    "(python tools/sharp_fixtures.py DIR; node tools/sharp_golden.mjs REF DIR > tests/golden/sharp_golden.json) — "
                    "until then classify / preprocess parity with real sharp stays UNPINNED")
def test_sharp_golden(oracle):
    golden = json.load(open(GOLDEN))
    rep = analyse(golden, oracle)
    print("sharp", golden.get("versions"), "\nswitch votes:", {k: v for k, v in rep.items() if k != "detail"}, "\n".join(rep["detail"]))
    default = {"luma_mode": 0, "blur_mode": 0, "coef_mode": 0, "reduce_mode": 0}
    assert rep["best"] == default, f"sharp's output is reproduced best by {rep['best']}, not by the defaults {default}: flip irp_opts / the oracle defaults"
    assert rep["luma_mode"][0] == 0, f"grey differs from sharp in {rep['luma_mode'][0]} bytes under the best luma_mode"
    assert rep["stencils"] == 0, rep["detail"]
    assert rep["blur_mode"][0] == 0, f"blur(1) differs from sharp in {rep['blur_mode'][0]} bytes under the best blur_mode"
    assert rep["resize_max_abs"][(0, 0)] <= 1, f"resized pixels differ from sharp by up to {rep['resize_max_abs'][(0, 0)]} LSB (bar: 1)"
    for rec in golden["cases"]:
        if rec["group"] != "classify" or rec["channels"] == 4:
            continue
        img, _ = pixels_of(rec["name"])
        mine = oracle.classify(img, is_jpeg=False)
        for k, v in rec["analysis"].items():
            assert rel_close(mine["scores"][k], v), f"{rec['name']} {k}: oracle {mine['scores'][k]} vs sharp {v}"
        assert rel_close(oracle.classify(img, is_jpeg=True)["scores"]["compression"], min(rec["blockiness"], 1.0)), rec["name"]
        st = rec["stats"]
        assert [int(s["sum"]) for s in st] == mine["sum"][:len(st)] and [int(s["squaresSum"]) for s in st] == mine["sumsq"][:len(st)]
