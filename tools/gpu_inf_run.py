import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from tests._pkg import load
m = load()
W, H, D = bench.W, bench.H, bench.D
img, lab = bench.synth_sample(0)
inf = m.UNet3d(1, 1, None)
inf.init_params(0)
inf.set_dim(W, H, D)
inf.prepare_for_inference()
for s in range(2):
    y = inf.forward(img, n_levels=1)[0]
print("ok", float(np.abs(y).mean()))
