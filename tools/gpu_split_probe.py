"""Diagnostic: logits error against the reference golden fixtures with the first-layer hi/lo split on and off (U3D_NO_SPLIT_INPUT is
read once per process, so the script re-executes itself)."""
import json
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


if len(sys.argv) > 1 and sys.argv[1] == "child":
    from tests._pkg import load
    m = load()
    for name in ("f1_fwd", "f1_train", "f1_collapse", "f2_eval"):
        z = np.load(os.path.join("tests", "golden", name + ".npz"))
        meta = json.loads(str(z["meta"]))
        net = m.UNet3d(meta["in_c"], meta["out_c"], str(z["feature"]))
        for i in range(net.param_count()):
            net.set_param(i, z[f"param_{i:03d}"])
        W, H, D = meta["dim"]
        net.set_dim(W, H, D)
        net.train(bool(meta["train"]))
        outs = net.forward(z["input"][0:1])
        print(os.environ.get("U3D_NO_SPLIT_INPUT", "split on "), name, "in_c", meta["in_c"], ["%.2e" % rel(o, z[f"logits_{k}"]) for k, o in enumerate(outs)], flush=True)
else:
    for env in ({}, {"U3D_NO_SPLIT_INPUT": "split off"}):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env))
