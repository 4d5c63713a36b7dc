"""One progressive decode of a few 3 MP files (for ncu: `ncu -k regex:prog_scan -c 1 ... python tools/prog_once.py`)."""
import io
import sys

import numpy as np
from PIL import Image, ImageFile

sys.path.insert(0, ".")
import irp_b200
from irp_b200.synth import synth_image

ImageFile.MAXBLOCK = 1 << 26
blobs = []
for i in range(4):
    b = io.BytesIO()
    Image.fromarray(synth_image(2048, 1536, idx=i)).save(b, "JPEG", quality=90, subsampling=2, progressive=True, optimize=True)
    blobs.append(b.getvalue())
with irp_b200.Engine(0) as eng:
    out = eng.decode_jpeg_batch(blobs)
    assert np.array_equal(out[0], np.asarray(Image.open(io.BytesIO(blobs[0]))))
print("ok")
