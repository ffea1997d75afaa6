"""Generates the golden vectors in tests/golden/*.npz by running the reference's own unet.cpp
(compiled unchanged: oracle/_ref/unet_ref, see oracle/build_ref.sh) on seeded synthetic inputs.

Run from the repo root IN THE BUILD CONTAINER (needs /root/reference to (re)build the binary):
    python tests/golden/make_golden.py
The .npz files are committed; tests never read /root/reference.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "unet_ref")

F1 = ("conv8,ks3,stride1+norm,leaky_relu+conv8,ks3,stride1+norm,leaky_relu\n"
      "conv16,ks3,stride2+norm,leaky_relu+conv16,ks3,stride1+norm,leaky_relu\n"
      "conv16,ks3,stride2+norm,leaky_relu+conv16,ks3,stride1+norm,leaky_relu+conv_trans16,ks2,stride2\n"
      "conv16,ks3,stride1+norm,leaky_relu+conv{out},ks1,stride1+conv_trans8,ks2,stride2\n"
      "conv8,ks3,stride1+norm,leaky_relu+conv{out},ks1,stride1")
F2 = ("conv8,ks3,stride1+bnorm,relu\n"
      "max_pool+conv16,ks3,stride1+bnorm,elu\n"
      "max_pool+conv16,ks3,stride1+norm,relu+upsample\n"
      "conv16,ks3,stride1+bnorm,relu+conv{out},ks1,stride1+upsample\n"
      "conv8,ks3,stride1+norm,elu+conv{out},ks1,stride1")

CASES = [
    # name, feature, in_c, out_c, (W,H,D), batch, train, steps, flags
    dict(name="f1_train", feature=F1, in_c=2, out_c=3, dim=(16, 24, 16), batch=2, train=1, steps=2,
         ce=1, dice=1, mse=1, collapse=0, invalid_labels=True),
    dict(name="f1_collapse", feature=F1, in_c=1, out_c=4, dim=(16, 16, 24), batch=1, train=1, steps=1,
         ce=1, dice=1, mse=0, collapse=2, invalid_labels=False),
    dict(name="f2_train", feature=F2, in_c=1, out_c=2, dim=(24, 16, 16), batch=1, train=1, steps=1,
         ce=1, dice=0, mse=1, collapse=0, invalid_labels=False),
    # BatchNorm3d in true eval mode: 3 training steps move the running statistics, then the validation forward + level-0 losses
    # (train.cpp:834-840: output_model->eval(), NoGradGuard) on the first sample
    dict(name="f2_validate", feature=F2, in_c=1, out_c=2, dim=(24, 16, 16), batch=1, train=1, steps=3, validate=1,
         ce=1, dice=1, mse=1, collapse=0, invalid_labels=False),
    dict(name="f2_eval", feature=F2, in_c=1, out_c=2, dim=(24, 16, 16), batch=1, train=0, steps=0, eval=1,
         ce=1, dice=1, mse=1, collapse=0, invalid_labels=False),
    dict(name="f1_fwd", feature=F1, in_c=2, out_c=3, dim=(24, 16, 32), batch=1, train=0, steps=0, eval=1, dump_acts=1,
         ce=1, dice=1, mse=1, collapse=0, invalid_labels=False),
]


def synth(rng, in_c, out_c, dim, batch, invalid):
    W, H, D = dim
    z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    ins, labs = [], []
    for b in range(batch):
        c = np.array([D, H, W]) * (0.5 + 0.1 * rng.uniform(-1, 1, 3))
        r = np.sqrt(((z - c[0]) / (0.4 * D)) ** 2 + ((y - c[1]) / (0.4 * H)) ** 2 + ((x - c[2]) / (0.4 * W)) ** 2)
        lab = np.clip(np.floor((1.0 - r) * out_c * 1.2), 0, out_c - 1).astype(np.float32)
        if invalid:  # labels >= C are masked by `valid` (train.cpp:523)
            lab[rng.uniform(size=lab.shape) < 0.02] = out_c + 1
        img = np.stack([(np.clip(1.2 - r, 0, 1) * (0.5 + 0.5 * ch) + rng.uniform(0, 0.05, r.shape)) for ch in range(in_c)])
        img = (img / img.max()).astype(np.float32)
        ins.append(img)
        labs.append(lab)
    return np.stack(ins), np.stack(labs)


def main():
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        rng = np.random.default_rng(abs(hash(case["name"])) % (2 ** 31) if False else sum(map(ord, case["name"])))
        feature = case["feature"].format(out=case["out_c"])
        ins, labs = synth(rng, case["in_c"], case["out_c"], case["dim"], case["batch"], case["invalid_labels"])
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "feature.txt"), "w").write(feature)
            ins.tofile(os.path.join(td, "input.bin"))
            labs.tofile(os.path.join(td, "label.bin"))
            cmd = [REF, "dump", "--in_c", str(case["in_c"]), "--out_c", str(case["out_c"]), "--feature",
                   os.path.join(td, "feature.txt"), "--dim", *map(str, case["dim"]), "--seed", "0",
                   "--input", os.path.join(td, "input.bin"), "--label", os.path.join(td, "label.bin"),
                   "--outdir", td, "--batch", str(case["batch"]), "--train", str(case["train"]),
                   "--steps", str(max(case["steps"], 1)), "--total_steps", "10", "--lr", "0.01",
                   "--ce", str(case["ce"]), "--dice", str(case["dice"]), "--mse", str(case["mse"]),
                   "--collapse", str(case["collapse"]), "--eval", str(case.get("eval", 0)),
                   "--dump_acts", str(case.get("dump_acts", 0)), "--validate", str(case.get("validate", 0))]
            subprocess.check_call(cmd)
            man = json.load(open(os.path.join(td, "manifest.json")))
            out = dict(feature=np.array(feature), input=ins, label=labs,
                       meta=np.array(json.dumps({k: v for k, v in case.items() if k != "feature"} | {"lr": 0.01, "total_steps": 10,
                                                                                             "torch": man["torch"]})))
            for i, p in enumerate(man["params"]):
                out[f"param_{i:03d}"] = np.fromfile(os.path.join(td, f"param_{i:03d}.bin"), np.float32).reshape(p["shape"])
                if case["train"]:
                    out[f"grad_{i:03d}"] = np.fromfile(os.path.join(td, f"grad_{i:03d}.bin"), np.float32).reshape(p["shape"])
                    out[f"after_{i:03d}"] = np.fromfile(os.path.join(td, f"param_after_{i:03d}.bin"), np.float32).reshape(p["shape"])
            out["param_names"] = np.array([p["name"] for p in man["params"]])
            k = 0
            while os.path.exists(os.path.join(td, f"logits_{k}.bin")):
                out[f"logits_{k}"] = np.fromfile(os.path.join(td, f"logits_{k}.bin"), np.float32)
                k += 1
            if case["train"]:
                out["level_losses"] = np.fromfile(os.path.join(td, "level_losses.bin"), np.float32).reshape(-1, 3)
                out["logged_losses"] = np.array(man["losses"], np.float32)
            if case.get("validate"):
                out["validate_losses"] = np.fromfile(os.path.join(td, "validate_losses.bin"), np.float32)
                out["validate_logits_0"] = np.fromfile(os.path.join(td, "validate_logits_0.bin"), np.float32)
            for f in os.listdir(td):
                if f.startswith("act_"):
                    out[f[:-4]] = np.fromfile(os.path.join(td, f), np.float32)
            path = os.path.join(ROOT, "tests", "golden", case["name"] + ".npz")
            np.savez_compressed(path, **out)
            print(case["name"], "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
