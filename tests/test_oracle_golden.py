"""Pins the CPU oracle (oracle/unet3d_oracle.py) to golden vectors produced by the reference's own
unet.cpp + calc_losses compiled against libtorch (tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import unet3d_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", CASES)
def test_structure_matches_reference(name):
    z, meta = load(name)
    net = O.parse_feature(meta["in_c"], meta["out_c"], str(z["feature"]))
    names = [str(s) for s in z["param_names"]]
    assert net.param_names == names
    for i, shp in enumerate(net.param_shapes):
        assert tuple(z[f"param_{i:03d}"].shape) == tuple(shp)


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    z, meta = load(name)
    torch.set_num_threads(4)
    net = O.parse_feature(meta["in_c"], meta["out_c"], str(z["feature"]))
    P = [torch.from_numpy(z[f"param_{i:03d}"].copy()) for i in range(len(net.param_shapes))]
    x = torch.from_numpy(z["input"][0:1].copy())
    taps = {}
    with torch.no_grad():
        outs = O.forward(net, P, x, training=bool(meta["train"]), taps=taps)
    for k, o in enumerate(outs):
        assert rel(o.numpy().ravel(), z[f"logits_{k}"]) < 2e-5, (name, k)
    for key in z.files:
        if key.startswith("act_"):
            assert rel(taps[key[4:]].numpy().ravel(), z[key]) < 2e-5, key


@pytest.mark.parametrize("name", [c for c in CASES if "train" in c or "collapse" in c])
def test_step_matches_reference(name):
    z, meta = load(name)
    torch.set_num_threads(4)
    net = O.parse_feature(meta["in_c"], meta["out_c"], str(z["feature"]))
    n = len(net.param_shapes)
    P = [torch.from_numpy(z[f"param_{i:03d}"].copy()) for i in range(n)]
    mom = [None] * n
    B = meta["batch"]
    xs = [torch.from_numpy(z["input"][b:b + 1].copy()) for b in range(B)]
    ts = [torch.from_numpy(z["label"][b:b + 1].copy()).to(torch.long) for b in range(B)]
    for s in range(meta["steps"]):
        lr = O.poly_lr(meta["lr"], s, meta["total_steps"])
        if s == 0:
            total, per_level, _ = O.micro_batch_loss(net, P, xs[0], ts[0], meta["ce"], meta["dice"], meta["mse"], meta["collapse"])
            got = torch.stack(per_level).numpy()
            np.testing.assert_allclose(got, z["level_losses"], rtol=2e-4, atol=2e-6)
        logged, grads, _ = O.train_step(net, P, mom, xs, ts, lr, meta["ce"], meta["dice"], meta["mse"], meta["collapse"])
        np.testing.assert_allclose(logged.numpy(), z["logged_losses"][s], rtol=2e-4, atol=2e-6)
        if s == 0:
            for i in range(n):
                g = z[f"grad_{i:03d}"]
                err = np.linalg.norm(grads[i].numpy() - g)
                assert err <= 5e-4 * np.linalg.norm(g) + 1e-6, (name, i, net.param_names[i], err, np.linalg.norm(g))
    for i in range(n):
        assert rel(P[i].numpy(), z[f"after_{i:03d}"]) < 2e-5, (name, i, net.param_names[i])


def test_validation_forward_uses_running_statistics():
    """f2_validate: 3 training steps move the BatchNorm3d running statistics; the validation forward (eval(), NoGradGuard,
    train.cpp:834-840) then normalises with them.  Pins the oracle's eval mode to the reference binary's output."""
    z, meta = load("f2_validate")
    torch.set_num_threads(4)
    net = O.parse_feature(meta["in_c"], meta["out_c"], str(z["feature"]))
    n = len(net.param_shapes)
    P = [torch.from_numpy(z[f"param_{i:03d}"].copy()) for i in range(n)]
    mom = [None] * n
    bn = O.init_bn_state(net)
    x = torch.from_numpy(z["input"][0:1].copy())
    t = torch.from_numpy(z["label"][0:1].copy()).to(torch.long)
    for s in range(meta["steps"]):
        logged, _, _ = O.train_step(net, P, mom, [x], [t], O.poly_lr(meta["lr"], s, meta["total_steps"]), meta["ce"], meta["dice"],
                                    meta["mse"], meta["collapse"], bn_state=bn)
        np.testing.assert_allclose(logged.numpy(), z["logged_losses"][s], rtol=2e-4, atol=2e-6)
    with torch.no_grad():
        out0 = O.forward(net, P, x, training=False, bn_state=bn)[0]
        ce, dice, mse = O.calc_losses(out0, t, net.out_count, 0)
    assert rel(out0.numpy().ravel(), z["validate_logits_0"]) < 5e-5
    np.testing.assert_allclose(torch.stack([ce, dice, mse]).numpy(), z["validate_losses"], rtol=2e-4, atol=2e-6)


def test_parser_errors_match_reference():
    # messages from unet.cpp:53,66,88,117
    with pytest.raises(RuntimeError, match="invalid u-net structure"):
        O.parse_feature(1, 1, "conv8\nconv8")
    with pytest.raises(RuntimeError, match="conv supports only ks1 stride1, ks3 stride1, and ks3 stride2"):
        O.parse_feature(1, 1, "conv8,ks5\nconv8\nconv8")
    with pytest.raises(RuntimeError, match="conv_trans supports only ks2 stride2"):
        O.parse_feature(1, 1, "conv8\nconv8+conv_trans8,ks3\nconv8")
    with pytest.raises(RuntimeError, match="unknown layer"):
        O.parse_feature(1, 1, "conv8\nfoo\nconv8")


def test_default_feature_matches_reference_binary_text():
    # default_feature(2) text as printed by the reference binary (train.cpp:1054-1069), frozen here
    f = O.default_feature(2)
    assert f.count("\n") == 10
    net = O.parse_feature(1, 2, f)
    assert len(net.param_shapes) == 108
    assert sum(int(np.prod(s)) for s in net.param_shapes) == 15023818
    net1 = O.parse_feature(1, 1, O.default_feature(1))
    assert sum(int(np.prod(s)) for s in net1.param_shapes) == 15023317
    ref = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "unet_ref")
    if os.path.exists(ref):
        import subprocess
        assert subprocess.check_output([ref, "feature", "--out_c", "2"]).decode() == f
