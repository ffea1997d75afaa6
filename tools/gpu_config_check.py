"""Runs the other BASELINE.json configurations once on the GPU (finite results, timings): cfg3 (6 classes, 160x192x160 training),
cfg4 (rodent grid 128x160x96 training), cfg5 (320^3 single-pass inference and its 8 windows of 160x192x160)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._pkg import load

m = load()


def sample(W, H, D, nclass, seed=0):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    r = np.sqrt(((z - D / 2) / (0.42 * D)) ** 2 + ((y - H / 2) / (0.40 * H)) ** 2 + ((x - W / 2) / (0.38 * W)) ** 2)
    img = (np.clip(1.1 - r, 0, 1) + 0.05 * rng.random(r.shape, dtype=np.float32)).astype(np.float32)
    img /= img.max()
    lab = np.zeros_like(r)
    for k in range(1, nclass):
        lab += (r < 1.0 - (k - 1) * 0.18)
    return img[None, None], lab[None].astype(np.float32)


def train_cfg(name, out_c, W, H, D, steps=3):
    net = m.UNet3d(1, out_c, None, gpu=0)
    net.init_params(0)
    net.set_dim(W, H, D)
    net.train(True)
    net.create_optimizer(1e-3)
    img, lab = sample(W, H, D, out_c)
    t0 = time.time()
    for s in range(steps):
        loss = m.train_microbatch_augmented(net, img, lab, seed=s)
        net.step(1, 1e-3)
        assert np.isfinite(loss).all() and not net.last_step_skipped(), (name, loss)
    print(f"{name}: UNet3d(1,{out_c}) {W}x{H}x{D} losses {loss} ({(time.time() - t0) / steps * 1e3:.1f} ms per host-buffer step)", flush=True)


train_cfg("cfg3", 6, 160, 192, 160)
train_cfg("cfg4", 2, 128, 160, 96)
net = m.UNet3d(1, 6, None, gpu=0)
net.init_params(0)
net.prepare_for_inference()
img, _ = sample(320, 320, 320, 6)
t0 = time.time()
y = net.forward(img, n_levels=1)[0]
print(f"cfg5 single pass 320^3: logits {y.shape} finite={np.isfinite(y).all()} {time.time() - t0:.2f} s incl. host copies", flush=True)
win = np.ascontiguousarray(img[:, :, 0:160, 0:192, 0:160])
yw = net.forward(win, n_levels=1)[0]
print(f"cfg5 window 160x192x160: finite={np.isfinite(yw).all()} mean|y|={np.abs(yw).mean():.4f}", flush=True)
