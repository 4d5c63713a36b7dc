import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, irp_b200
from irp_b200.synth import synth_image
with irp_b200.Engine(0) as eng:
    base = synth_image(6600, 6100, 3)
    eng.preprocess_batch([eng.upload(np.ascontiguousarray(base[:3000, :4000]))])
    for (w, h) in [(5986, 3991), (4597, 4597), (3001, 2001), (2500, 2500), (4001, 3003), (3841, 2161)]:
        d = eng.upload(np.ascontiguousarray(base[:h, :w]))
        ow, oh = eng.preprocess_dims(w, h)
        o = eng.alloc_device(ow, oh, 3)
        ts = []
        for k in range(3):
            t0 = time.perf_counter(); eng.preprocess_batch([d], device_outputs=[o]); ts.append(1e3 * (time.perf_counter() - t0))
        print(f"{w}x{h}: first call {ts[0]:.2f} ms, then {ts[1]:.2f} / {ts[2]:.2f} ms")
