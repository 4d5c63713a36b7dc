import sys, numpy as np
sys.path.insert(0, '.')
import irp_b200
from irp_b200.synth import synth_image
from oracle import oracle
w, h, o = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
img = synth_image(w, h, idx=3)
with irp_b200.Engine(0) as eng:
    outs = eng.preprocess_batch([img], orientations=[o])
ref = oracle.preprocess(img, o)
print(w, h, o, outs[0].shape, "equal" if np.array_equal(outs[0], ref) else f"DIFF max {np.abs(outs[0].astype(int)-ref.astype(int)).max()} n {(outs[0]!=ref).sum()}")
