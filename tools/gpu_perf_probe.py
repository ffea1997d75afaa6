"""GPU perf probe: default net at BASELINE config sizes; forward and train micro-batch + step timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests._pkg import load
m = load()
W, H, D = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (160, 192, 160))]
out_c = int(sys.argv[4]) if len(sys.argv) > 4 else 2
mode = sys.argv[5] if len(sys.argv) > 5 else "both"
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
net = m.UNet3d(1, out_c)
net.init_params(0)
net.set_dim(W, H, D)
x = torch.rand(1, 1, D, H, W, device="cuda")
lab = torch.randint(0, out_c, (1, D, H, W), device="cuda").float()
torch.cuda.synchronize()
if mode in ("fwd", "both"):
    net.prepare_for_inference()
    out = torch.empty(1, out_c, D, H, W, device="cuda")
    for i in range(iters + 2):
        net.timer_start()
        net.device_forward(x.data_ptr(), [out.data_ptr()])
        ms = net.timer_stop()
        print(f"fwd iter {i}: {ms:.3f} ms  -> {W*H*D/ms/1e3:.1f} Mvoxel/s  finite={bool(torch.isfinite(out).all())}", flush=True)
if mode in ("train", "both"):
    net.train(True); net.create_optimizer(1e-3)
    for i in range(iters + 2):
        net.timer_start()
        l = net.device_train_microbatch(x.data_ptr(), lab.data_ptr())
        ms1 = net.timer_stop()
        net.timer_start()
        gn = net.step(1, 1e-3)
        ms2 = net.timer_stop()
        print(f"train iter {i}: microbatch {ms1:.3f} ms, step {ms2:.3f} ms, loss {l}, gnorm {gn:.4f} skipped={net.last_step_skipped()} scale={net.loss_scale()}", flush=True)
print("launches", net.launch_count(), "mem GB", torch.cuda.mem_get_info())
