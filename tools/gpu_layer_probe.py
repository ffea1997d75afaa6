"""Runs ONE layer (forward, optionally backward) through the op-level C-ABI at a given size, for ncu captures of a single kernel.
usage: gpu_layer_probe.py transposed ks stride cin0 cin1 cout W H D [bwd]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._pkg import load

tr, ks, st, c0, c1, co, W, H, D = [int(v) for v in sys.argv[1:10]]
bwd = len(sys.argv) > 10
m = load()
rng = np.random.default_rng(0)
x0 = rng.standard_normal((c0, D, H, W), dtype=np.float32)
x1 = rng.standard_normal((c1, D, H, W), dtype=np.float32) if c1 else None
kk = 2 if tr else ks
shape = (c0 + c1, co, kk, kk, kk) if tr else (co, c0 + c1, kk, kk, kk)
w = (rng.standard_normal(shape, dtype=np.float32) / np.sqrt((c0 + c1) * kk ** 3)).astype(np.float32)
b = rng.standard_normal(co, dtype=np.float32)
y = m.conv_forward(x0, w, b, x1, transposed=bool(tr), ks=ks, stride=st)
print("fwd ok", y.shape, float(np.abs(y).mean()))
if bwd:
    dy = rng.standard_normal(y.shape, dtype=np.float32)
    gx0, gx1, gw = m.conv_backward(x0, w, dy, x1, transposed=bool(tr), ks=ks, stride=st)
    print("bwd ok", float(np.abs(gx0).mean()), float(np.abs(gw).mean()))
