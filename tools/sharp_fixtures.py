#!/usr/bin/env python
"""Write the inputs of the sharp pin as PNG files (lossless: sharp sees exactly the seeded pixels) plus a manifest.
    python tools/sharp_fixtures.py /tmp/irp_fixtures
    node tools/sharp_golden.mjs /path/to/image-restoration-platform /tmp/irp_fixtures > tests/golden/sharp_golden.json
    python -m pytest tests/test_sharp_golden.py -q
EXIF orientation travels in the PNG's eXIf chunk (libvips reads it; sharp's .rotate() honours it)."""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main(out_dir: str) -> None:
    from conftest import rand_image
    from sharp_cases import CLASSIFY_CASES, PREPROCESS_CASES

    os.makedirs(out_dir, exist_ok=True)
    manifest = []
    for group, cases in (("classify", CLASSIFY_CASES), ("preprocess", PREPROCESS_CASES)):
        for name, h, w, c, seed, kind, orientation in cases:
            a = rand_image(h, w, c, seed, kind)
            img = Image.fromarray(a[:, :, 0] if c == 1 else a, {1: "L", 3: "RGB", 4: "RGBA"}[c])
            kw = {}
            if orientation != 1:
                ex = Image.Exif()
                ex[0x0112] = orientation
                kw["exif"] = ex.tobytes()
            path = os.path.join(out_dir, name + ".png")
            img.save(path, "PNG", compress_level=1, **kw)
            back = np.asarray(Image.open(path))
            assert np.array_equal(back.reshape(a.shape), a), name
            manifest.append({"name": name, "group": group, "file": name + ".png", "width": w, "height": h, "channels": c,
                             "orientation": orientation, "pixels_sha256": hashlib.sha256(a.tobytes()).hexdigest()})
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(f"wrote {len(manifest)} fixtures and manifest.json to {out_dir}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/irp_sharp_fixtures")
