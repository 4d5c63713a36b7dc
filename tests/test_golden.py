"""Committed fixtures (tests/golden/golden.json, made by tests/golden/make_golden.py): the oracle must
keep producing them (CPU) and the CUDA path must reproduce them through the C ABI (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import assert_result_parity, rand_image

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("g", GOLD["classify"], ids=lambda g: "x".join(map(str, g["case"][:3])))
def test_oracle_classify_matches_golden(oracle, g):
    h, w, c, seed, kind, jpeg = g["case"]
    img = rand_image(h, w, c, seed, kind)
    assert sha(img) == g["input_sha256"], "seeded input generator drifted"
    got = oracle.classify(img, is_jpeg=jpeg)
    for k, v in g["result"].items():
        if k == "scores":  # a 1x1 image has stdev 0/0 = NaN, as vips_stats would: NaN == NaN here
            for sk, sv in v.items():
                assert got["scores"][sk] == sv or (sv != sv and got["scores"][sk] != got["scores"][sk]), (k, sk)
        else:
            assert got[k] == v, k


@pytest.mark.parametrize("g", GOLD["preprocess"], ids=lambda g: "x".join(map(str, g["case"][:3])) + f"o{g['case'][5]}")
def test_oracle_preprocess_matches_golden(oracle, g):
    h, w, c, seed, kind, o = g["case"]
    r = oracle.preprocess(rand_image(h, w, c, seed, kind), o)
    assert list(r.shape) == g["shape"] and sha(r) == g["sha256"]


@pytest.mark.parametrize("g", GOLD["fusion"], ids=lambda g: "x".join(map(str, g["case"][:3])))
def test_oracle_fusion_matches_golden(oracle, g):
    h, w, c, seed, kind, o = g["case"]
    assert sha(oracle.fusion_canvas(rand_image(h, w, c, seed, kind), o)) == g["sha256"]


@pytest.mark.gpu
def test_cuda_classify_matches_golden(engine):
    imgs = [rand_image(*g["case"][:5]) for g in GOLD["classify"]]
    got = engine.classify_batch(imgs, is_jpeg=[g["case"][5] for g in GOLD["classify"]])
    for g, r in zip(GOLD["classify"], got):
        assert_result_parity(r, g["result"], g["case"][2], str(g["case"]))


@pytest.mark.gpu
def test_cuda_preprocess_and_fusion_match_golden(engine):
    imgs = [rand_image(*g["case"][:5]) for g in GOLD["preprocess"]]
    outs = engine.preprocess_batch(imgs, orientations=[g["case"][5] for g in GOLD["preprocess"]])
    for g, o in zip(GOLD["preprocess"], outs):
        assert list(o.shape) == g["shape"] and sha(o) == g["sha256"], g["case"]
    groups = engine.fusion_prepare_batch([[rand_image(*g["case"][:5]) for g in GOLD["fusion"]]],
                                         orientations=[[g["case"][5] for g in GOLD["fusion"]]])
    for g, canvas in zip(GOLD["fusion"], groups[0]):
        assert sha(canvas) == g["sha256"], g["case"]
