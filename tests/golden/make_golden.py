#!/usr/bin/env python
"""Regenerate tests/golden/golden.json from the CPU oracle.

The reference ships no golden vectors for this path and sharp/libvips cannot run offline (SURVEY.md
§8c), so these fixtures pin the ORACLE's outputs on seeded inputs: classify integer statistics and
scores verbatim, preprocess / fusion outputs as SHA-256 of the pixel arrays.  tests/test_golden.py
checks the oracle (CPU) and the CUDA path (GPU) against them.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CLASSIFY_CASES = [  # (h, w, c, seed, kind, is_jpeg)
    (1, 1, 3, 1, "noise", True), (7, 5, 3, 2, "noise", True), (33, 257, 3, 3, "smooth", True), (128, 128, 3, 4, "noise", True),
    (97, 513, 3, 5, "edges", True), (300, 700, 1, 6, "smooth", True), (64, 64, 4, 7, "noise", False), (512, 640, 3, 8, "smooth", True),
]
PREPROCESS_CASES = [  # (h, w, c, seed, kind, orientation)
    (37, 53, 3, 11, "noise", 1), (37, 53, 3, 12, "noise", 6), (2100, 2300, 3, 13, "smooth", 1), (2300, 2500, 3, 14, "noise", 7),
    (2160, 3840, 3, 15, "smooth", 1), (2500, 2100, 4, 16, "noise", 8), (2500, 2100, 1, 17, "smooth", 3),
]
FUSION_CASES = [(2600, 3400, 3, 21, "smooth", 1), (3000, 2100, 3, 22, "noise", 6), (500, 700, 1, 23, "noise", 1)]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from conftest import rand_image
    from oracle import oracle

    out = {"classify": [], "preprocess": [], "fusion": []}
    for h, w, c, seed, kind, jpeg in CLASSIFY_CASES:
        img = rand_image(h, w, c, seed, kind)
        out["classify"].append({"case": [h, w, c, seed, kind, jpeg], "input_sha256": sha(img), "result": oracle.classify(img, is_jpeg=jpeg)})
    for h, w, c, seed, kind, o in PREPROCESS_CASES:
        img = rand_image(h, w, c, seed, kind)
        r = oracle.preprocess(img, o)
        out["preprocess"].append({"case": [h, w, c, seed, kind, o], "input_sha256": sha(img), "shape": list(r.shape), "sha256": sha(r)})
    for h, w, c, seed, kind, o in FUSION_CASES:
        img = rand_image(h, w, c, seed, kind)
        r = oracle.fusion_canvas(img, o)
        out["fusion"].append({"case": [h, w, c, seed, kind, o], "input_sha256": sha(img), "sha256": sha(r)})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote golden.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
