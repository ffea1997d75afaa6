"""Workload for the per-layer profile: 3 cfg-2 training micro-batches + updates, then 3 cfg-1 inference forwards of one 160x192x160
window.  Run under ncu with U3D_TRACE_LAUNCHES=<file> U3D_ONE_STREAM=1 U3D_NO_GRAPH=1 (tools/per_layer_table.py joins the two)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tests._pkg import load

m = load()
W, H, D = bench.W, bench.H, bench.D
img, lab = bench.synth_sample(0)
net = m.UNet3d(1, 2, None)
net.init_params(0)
net.set_dim(W, H, D)
net.train(True)
net.create_optimizer(1e-3)
for s in range(3):
    loss = net.train_microbatch(img, lab)
    net.step(1, 1e-3)
print("train loss", loss)
del net
inf = m.UNet3d(1, 1, None)
inf.init_params(0)
inf.set_dim(W, H, D)
inf.prepare_for_inference()
for s in range(3):
    y = inf.forward(img, n_levels=1)[0]
print("inference finite", bool(np.isfinite(y).all()))
