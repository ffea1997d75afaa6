"""GPU diagnostic: model-level parity against the golden fixtures (reference unet.cpp outputs)."""
import sys, os, json, glob, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._pkg import load
m = load()
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
for name in names:
    z = np.load(os.path.join(GOLD, name + ".npz")); meta = json.loads(str(z["meta"]))
    print("=====", name, meta, flush=True)
    try:
        net = m.UNet3d(meta["in_c"], meta["out_c"], str(z["feature"]))
        n = net.param_count()
        assert [net.param_name(i) for i in range(n)] == [str(s) for s in z["param_names"]]
        for i in range(n): net.set_param(i, z[f"param_{i:03d}"])
        W, H, D = meta["dim"]; net.set_dim(W, H, D)
        if not meta["train"]:
            net.prepare_for_inference()
            outs = net.forward(z["input"][0:1])
            for k, o in enumerate(outs):
                print(f"  logits[{k}] rel={rel(o, z[f'logits_{k}']):.3e} maxabs={np.abs(o.ravel()-z[f'logits_{k}']).max():.3e} scale={np.abs(z[f'logits_{k}']).max():.3f}", flush=True)
            continue
        net.train(True); net.create_optimizer(meta["lr"])
        B = meta["batch"]
        for s in range(meta["steps"]):
            lr = m.poly_lr(meta["lr"], s, meta["total_steps"])
            logged = np.zeros(3)
            for b in range(B):
                l0, lv = net.train_microbatch(z["input"][b:b+1], z["label"][b:b+1], meta["collapse"], meta["ce"], meta["dice"], meta["mse"], all_levels=True)
                logged += l0
                if s == 0 and b == 0:
                    print("  level losses ours:\n", lv, "\n  ref:\n", z["level_losses"], flush=True)
            print(f"  step {s} logged ours={logged/B} ref={z['logged_losses'][s]}", flush=True)
            if s == 0:
                num = den = 0.0
                for i in range(n):
                    g = net.get_grad(i); gr = z[f"grad_{i:03d}"]
                    num += ((g - gr).astype(np.float64) ** 2).sum(); den += (gr.astype(np.float64) ** 2).sum()
                    print(f"    grad {i:3d} {net.param_name(i):28s} rel={rel(g, gr):.3e} |ref|={np.linalg.norm(gr):.3e} |ours|={np.linalg.norm(g):.3e}", flush=True)
                print(f"  global grad rel err = {np.sqrt(num/den):.3e}")
            gn = net.step(B, lr)
            print(f"  grad norm {gn:.4f} skipped={net.last_step_skipped()} loss_scale={net.loss_scale()}", flush=True)
        worst = max(rel(net.get_param(i), z[f"after_{i:03d}"]) for i in range(n))
        print(f"  params after: worst rel={worst:.3e}", flush=True)
    except Exception as ex:
        print("  EXC", ex, flush=True)
        if "failed" in str(ex) or "timeout" in str(ex): sys.exit(3)
