"""Timing probe of the multi-scan (progressive) JPEG route: decode a batch of progressive 12 MP / 3 MP files on the
device, compare with Pillow on the first, and report the call time next to the baseline route's on the same pixels."""
import io
import sys
import time

import numpy as np
from PIL import Image, ImageFile

sys.path.insert(0, ".")
import irp_b200
from irp_b200.synth import synth_image

ImageFile.MAXBLOCK = 1 << 26


def enc(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for (w, h) in ((2048, 1536), (4000, 3000)):
        imgs = [synth_image(w, h, idx=i) for i in range(min(n, 4))]
        prog = [enc(imgs[i % len(imgs)], quality=90, subsampling=2, progressive=True, optimize=True) for i in range(n)]
        base = [enc(imgs[i % len(imgs)], quality=90, subsampling=2) for i in range(n)]
        with irp_b200.Engine(0) as eng:
            for name, blobs in (("baseline", base), ("progressive", prog)):
                out = eng.decode_jpeg_batch(blobs)
                assert np.array_equal(out[0], np.asarray(Image.open(io.BytesIO(blobs[0])))), name
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    eng.decode_jpeg_batch(blobs)
                    ts.append((time.perf_counter() - t0) * 1e3)
                print(f"{w}x{h} x {n} {name}: {min(ts):.1f} ms per call (host pixels back), {sum(len(b) for b in blobs) / 1e6:.1f} MB of files")
        t0 = time.perf_counter()
        np.asarray(Image.open(io.BytesIO(prog[0])))
        print(f"  Pillow, one progressive file on one core: {(time.perf_counter() - t0) * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
