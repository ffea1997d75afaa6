"""GPU parity of the model-level C-ABI against (a) golden vectors produced by the reference's own unet.cpp
(tests/golden), (b) the CPU oracle / the live reference binary on the default network, (c) size-independent
properties at BASELINE.json's full 160x192x160 grid.

Tolerances (DESIGN.md "precision"): the CUDA path stores activations and weights in fp16 (fp32 accumulate,
fp32 statistics/loss/optimizer).  North-star tolerance for a 16-bit pipeline: logits within 1e-2 relative error.
Gradients of early layers inherit the forward rounding amplified by the network (measured, documented)."""
import glob
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import unet3d_oracle as O
from tests._pkg import load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "unet_ref")


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def build_from_golden(m, z, meta):
    net = m.UNet3d(meta["in_c"], meta["out_c"], str(z["feature"]))
    n = net.param_count()
    assert [net.param_name(i) for i in range(n)] == [str(s) for s in z["param_names"]]
    for i in range(n):
        assert net.param_shape(i) == tuple(z[f"param_{i:03d}"].shape)
        net.set_param(i, z[f"param_{i:03d}"])
    W, H, D = meta["dim"]
    net.set_dim(W, H, D)
    return net


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(name):
    m = load()
    z, meta = golden(name)
    net = build_from_golden(m, z, meta)
    net.train(bool(meta["train"]))
    outs = net.forward(z["input"][0:1])
    for k, o in enumerate(outs):
        e = rel(o, z[f"logits_{k}"])
        assert e < (1e-2 if k == 0 else 2e-2), (name, k, e)


@pytest.mark.parametrize("name", [c for c in CASES if json.loads(str(np.load(os.path.join(GOLD, c + ".npz"))["meta"]))["train"]])
def test_training_step_matches_reference_golden(name):
    m = load()
    z, meta = golden(name)
    net = build_from_golden(m, z, meta)
    n = net.param_count()
    net.train(True)
    net.create_optimizer(meta["lr"])
    B = meta["batch"]
    for s in range(meta["steps"]):
        lr = m.poly_lr(meta["lr"], s, meta["total_steps"])
        logged = np.zeros(3)
        for b in range(B):
            l0, lv = net.train_microbatch(z["input"][b:b + 1], z["label"][b:b + 1], meta["collapse"], meta["ce"], meta["dice"],
                                          meta["mse"], all_levels=True)
            logged += l0
            if s == 0 and b == 0:
                np.testing.assert_allclose(lv, z["level_losses"], rtol=0, atol=1e-3)
        np.testing.assert_allclose(logged / B, z["logged_losses"][s], rtol=0, atol=2e-3)
        if s == 0:
            num = den = 0.0
            for i in range(n):
                g, gr = net.get_grad(i), z[f"grad_{i:03d}"]
                num += float(((g - gr).astype(np.float64) ** 2).sum()); den += float((gr.astype(np.float64) ** 2).sum())
                name_i = net.param_name(i)
                if name_i.startswith(("output", "decode0")) and np.linalg.norm(gr) > 1e-4:
                    assert rel(g, gr) < 1e-2, (name_i, rel(g, gr))   # layers next to the loss: tight
            assert np.sqrt(num / den) < 5e-2, np.sqrt(num / den)          # whole gradient incl. amplified early layers
        net.step(B, lr)
        assert not net.last_step_skipped()
    num = den = 0.0
    for i in range(n):
        d_ours = net.get_param(i) - z[f"param_{i:03d}"]
        d_ref = z[f"after_{i:03d}"] - z[f"param_{i:03d}"]
        num += float(((d_ours - d_ref).astype(np.float64) ** 2).sum()); den += float((d_ref.astype(np.float64) ** 2).sum())
    assert np.sqrt(num / den) < 6e-2, np.sqrt(num / den)


def synth_volume(W, H, D, seed=1):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    r = np.sqrt(((z - D / 2) / (0.42 * D)) ** 2 + ((y - H / 2) / (0.40 * H)) ** 2 + ((x - W / 2) / (0.38 * W)) ** 2)
    img = np.where(r < 1, 0.2 + 0.8 * np.clip(1 - r, 0, 1), 0) + rng.uniform(0, 0.05, r.shape) * (r < 1)
    img = (img / img.max()).astype(np.float32)
    lab = (r < 1).astype(np.float32)
    return img[None, None], lab[None]


def test_default_net_forward_matches_live_reference_or_oracle():
    """Default feature_string (train.cpp:1054-1069), random init from the reference binary when it is here."""
    m = load()
    W, H, D = 64, 96, 64
    img, _ = synth_volume(W, H, D)
    feature = O.default_feature(2)
    onet = O.parse_feature(1, 2, feature)
    with tempfile.TemporaryDirectory() as td:
        if os.path.exists(REF_BIN):
            img.tofile(os.path.join(td, "in.bin"))
            subprocess.check_call([REF_BIN, "dump", "--in_c", "1", "--out_c", "2", "--feature", "default", "--dim", str(W), str(H), str(D),
                                   "--seed", "0", "--input", os.path.join(td, "in.bin"), "--outdir", td, "--eval", "1"])
            P = [np.fromfile(os.path.join(td, f"param_{i:03d}.bin"), np.float32).reshape(s) for i, s in enumerate(onet.param_shapes)]
            ref = [np.fromfile(os.path.join(td, f"logits_{k}.bin"), np.float32) for k in range(5)]
        else:
            Pt = O.init_params(onet, 0)
            P = [p.numpy() for p in Pt]
            with torch.no_grad():
                ref = [o.numpy().ravel() for o in O.forward(onet, Pt, torch.from_numpy(img))]
    net = m.UNet3d(1, 2, feature)
    net.load_parameters(P)
    net.prepare_for_inference()
    outs = net.forward(img)
    errs = [rel(o, r) for o, r in zip(outs, ref)]
    print("default net 64x96x64 logits rel err per level:", errs)
    assert errs[0] < 1e-2, errs
    lab_a = outs[0][0].argmax(0); lab_b = ref[0].reshape(outs[0].shape)[0].argmax(0)
    assert (lab_a == lab_b).mean() > 0.99


def test_full_size_forward_properties_and_oracle_parity():
    """BASELINE config 1 grid (160x192x160): bit-determinism of the forward (fixed reduction orders), replica
    consistency through copy_from, and parity of logits[0] with the CPU oracle on the same weights and volume.
    (The random-init network amplifies a 1e-3 input perturbation ~50x, measured, so perturbation-style
    properties are not usable as tight checks; the oracle comparison is the parity statement.)"""
    m = load()
    W, H, D = 160, 192, 160
    feature = O.default_feature(1)
    onet = O.parse_feature(1, 1, feature)
    Pt = O.init_params(onet, 5)
    img, _ = synth_volume(W, H, D)
    net = m.UNet3d(1, 1, feature)
    net.load_parameters([p.numpy() for p in Pt])
    net.prepare_for_inference()
    a = net.forward(img, n_levels=1)[0]
    b = net.forward(img, n_levels=1)[0]
    assert np.isfinite(a).all()
    assert np.array_equal(a, b), "forward must be deterministic"
    twin = m.UNet3d(1, 1, feature)
    twin.copy_from(net)
    twin.prepare_for_inference()
    assert np.array_equal(twin.forward(img, n_levels=1)[0], a), "replica after copy_from must reproduce the forward bit for bit"
    del twin
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    with torch.no_grad():
        ref = O.forward(onet, Pt, torch.from_numpy(img))[0].numpy()
    e = rel(a, ref)
    print("full-size 160x192x160 logits[0] rel err vs CPU oracle:", e)
    assert e < 1e-2, e


def test_param_roundtrip_momentum_copy_from_and_windows():
    m = load()
    feature = str(np.load(os.path.join(GOLD, "f1_fwd.npz"))["feature"])
    a = m.UNet3d(2, 3, feature)
    a.init_params(11)
    P = a.parameters()
    b = m.UNet3d(2, 3, feature)
    b.copy_from(a)
    for i, p in enumerate(P):
        assert np.array_equal(b.get_param(i), p)
    b.set_momentum(0, np.full(a.param_shape(0), 0.25, np.float32))
    assert np.all(b.get_momentum(0) == 0.25)
    # evaluate window loop == forward()[0] per window (evaluate.cpp:223-230)
    W, H, D = 16, 16, 32
    a.set_dim(W, H, D)
    a.prepare_for_inference()
    wins = [np.random.default_rng(i).random((1, 2, D, H, W), dtype=np.float32) for i in range(5)]
    singles = [a.forward(w, n_levels=1)[0] for w in wins]
    import ctypes
    F = ctypes.POINTER(ctypes.c_float)
    outs = [np.empty_like(s) for s in singles]
    inp = (F * 5)(*[w.ctypes.data_as(F) for w in wins])
    outp = (F * 5)(*[o.ctypes.data_as(F) for o in outs])
    m.check(m.lib().unet3d_evaluate_windows(a._h, inp, outp, 5, 0))   # pipelined: two staging slots each way are reused
    for s, o in zip(singles, outs):
        assert np.array_equal(s, o)


def test_maxpool_indices_bit_exact():
    m = load()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 8, 12, 10, generator=g).half().float()
    x[0, :2] = 1.0          # ties: first maximum in scan order wins
    x[1, 3, 4, 5] = float("nan")
    y_ref, i_ref = torch.nn.functional.max_pool3d(x[None].cuda(), 2, 2, return_indices=True)
    y, idx = m.maxpool_forward(x.numpy())
    assert np.array_equal(idx, i_ref[0].cpu().numpy())
    np.testing.assert_array_equal(y, y_ref[0].cpu().numpy())


@pytest.mark.parametrize("name", ["f1_train", "f1_collapse"])
def test_gradients_match_fp16_emulating_oracle_tightly(name):
    """Separates implementation correctness from precision: against the oracle evaluated with the SAME fp16 storage points
    (oracle.Fp16Emulation) the GPU gradients agree 2.5-3x tighter than against the pure-fp32 reference (measured: global
    5.0e-3 / 7.1e-3 here vs 1.5e-2 / 1.7e-2), i.e. most of the fp32-reference gap is the storage precision amplified by the
    network, not the kernels; what remains is rounding-order differences (split accumulation, fp32-vs-fp16 statistics)."""
    m = load()
    z, meta = golden(name)
    net = build_from_golden(m, z, meta)
    net.train(True)
    net.create_optimizer(meta["lr"])
    x, lab = z["input"][0:1], z["label"][0:1]
    net.train_microbatch(x, lab, meta["collapse"], meta["ce"], meta["dice"], meta["mse"])
    n = net.param_count()
    onet = O.parse_feature(meta["in_c"], meta["out_c"], str(z["feature"]))
    P = [torch.from_numpy(z[f"param_{i:03d}"].copy()).requires_grad_(True) for i in range(n)]
    total, _, _ = O.micro_batch_loss(onet, P, torch.from_numpy(x.copy()), torch.from_numpy(lab.copy()).long(), meta["ce"], meta["dice"],
                                     meta["mse"], meta["collapse"], emu=O.Fp16Emulation(net.loss_scale()))
    total.backward()
    num = den = 0.0
    worst = 0.0
    per = []
    for i in range(n):
        g = net.get_grad(i)
        gr = P[i].grad.numpy() if P[i].grad is not None else np.zeros_like(g)
        num += float(((g - gr).astype(np.float64) ** 2).sum()); den += float((gr.astype(np.float64) ** 2).sum())
        if np.linalg.norm(gr) > 1e-3 and not onet.param_names[i].endswith(".bias"):
            worst = max(worst, rel(g, gr))
            per.append((rel(g, gr), onet.param_names[i], float(np.linalg.norm(gr))))
    print(name, "vs fp16-emulating oracle: global grad rel err", np.sqrt(num / den), "worst weight tensor", worst)
    print("   per tensor (rel err, name, |g|):", sorted(per, reverse=True)[:6])
    # the same gradients against the pure fp32 oracle: the emulation must explain the GPU result at least as well as fp32 does
    P32 = [torch.from_numpy(z[f"param_{i:03d}"].copy()).requires_grad_(True) for i in range(n)]
    t32, _, _ = O.micro_batch_loss(onet, P32, torch.from_numpy(x.copy()), torch.from_numpy(lab.copy()).long(), meta["ce"], meta["dice"],
                                   meta["mse"], meta["collapse"])
    t32.backward()
    n32 = d32 = ne = 0.0
    for i in range(n):
        g = net.get_grad(i)
        g32 = P32[i].grad.numpy() if P32[i].grad is not None else np.zeros_like(g)
        ge = P[i].grad.numpy() if P[i].grad is not None else np.zeros_like(g)
        n32 += float(((g - g32).astype(np.float64) ** 2).sum()); d32 += float((g32.astype(np.float64) ** 2).sum())
        ne += float(((ge - g32).astype(np.float64) ** 2).sum())
    gpu_vs_fp32, emu_vs_fp32 = np.sqrt(n32 / d32), np.sqrt(ne / d32)
    print(f"   GPU vs fp32 oracle {gpu_vs_fp32:.3e}; emulation vs fp32 oracle {emu_vs_fp32:.3e}")
    # small nets (96 voxels at the deepest level) amplify every rounding difference: the three numbers are of the same size; the GPU
    # must not be further from the emulation than the emulation's own distance from fp32
    assert np.sqrt(num / den) < max(1.2e-2, 1.25 * emu_vs_fp32), (np.sqrt(num / den), emu_vs_fp32)
    assert gpu_vs_fp32 < 3e-2, gpu_vs_fp32
    assert worst < 8e-2, worst


def test_default_net_training_microbatch_matches_oracle_at_band_kernel_sizes():
    """Default net (1 in / 2 out) at 128x96x64: large enough that the x-banded conv, the N-stacked weight gradient (levels 0-1
    natively, level 2 as channel-group pairs), the parity-stacked transpose / stride-2 data gradients, the TMA kernel and the
    fused output heads are all on the path.  Losses of all five levels and the parameter gradients against the fp32 CPU
    oracle on the same weights and sample (tolerances as in the golden-fixture training test: layers next to the loss
    tight, the whole gradient incl. the early layers that amplify the fp16 storage error ~50x looser)."""
    m = load()
    W, H, D = 128, 96, 64
    feature = O.default_feature(2)
    onet = O.parse_feature(1, 2, feature)
    P = [p.clone().requires_grad_(True) for p in O.init_params(onet, 9)]
    img, lab = synth_volume(W, H, D, seed=4)
    net = m.UNet3d(1, 2, feature)
    net.load_parameters([p.detach().numpy() for p in P])
    net.set_dim(W, H, D)
    net.train(True)
    net.create_optimizer(1e-3)
    l0, lv = net.train_microbatch(img, lab, all_levels=True)
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    total, per_level, _ = O.micro_batch_loss(onet, P, torch.from_numpy(img), torch.from_numpy(lab).long())
    total.backward()
    ref = torch.stack(per_level).detach().numpy()
    np.testing.assert_allclose(lv, ref, rtol=0, atol=2e-3)
    num = den = 0.0
    n = net.param_count()
    for i in range(n):
        g = net.get_grad(i)
        gr = P[i].grad.numpy() if P[i].grad is not None else np.zeros_like(g)
        num += float(((g - gr).astype(np.float64) ** 2).sum()); den += float((gr.astype(np.float64) ** 2).sum())
        name_i = net.param_name(i)
        if name_i.startswith(("output", "decode0")) and np.linalg.norm(gr) > 1e-4:
            assert rel(g, gr) < 1e-2, (name_i, rel(g, gr))
    print("128x96x64 default net: global gradient rel err vs fp32 oracle", np.sqrt(num / den))
    assert np.sqrt(num / den) < 5e-2, np.sqrt(num / den)
    net.step(1, 1e-3)
    assert not net.last_step_skipped()


def test_trained_net_label_maps_reach_dice_0999_against_oracle():
    """north_star: 'predicted label maps at Dice >= 0.999 against the reference'.  At random init the probabilities sit on the
    decision boundary and no 16-bit or TF32 arithmetic reaches that, so the statement is tested where it is
    meaningful: train the default net for a few dozen steps with THIS library, then compare the arg-max label map of its
    forward with the fp32 CPU oracle evaluated on the same trained weights."""
    m = load()
    W, H, D = 64, 64, 64
    feature = O.default_feature(2)
    net = m.UNet3d(1, 2, feature)
    net.init_params(3)
    net.set_dim(W, H, D)
    net.train(True)
    lr0, steps = 2e-2, 40
    net.create_optimizer(lr0)
    img, lab = synth_volume(W, H, D, seed=2)
    for s in range(steps):
        loss = net.train_microbatch(img, lab)
        net.step(1, m.poly_lr(lr0, s, steps))
        assert not net.last_step_skipped()
    print("loss after", steps, "steps:", loss)
    assert loss[1] < 0.2, loss          # soft-Dice loss: the net has learned the ellipsoid
    onet = O.parse_feature(1, 2, feature)
    P = [torch.from_numpy(p.copy()) for p in net.parameters()]
    net.prepare_for_inference()
    ours = net.forward(img, n_levels=1)[0]
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    with torch.no_grad():
        ref = O.forward(onet, P, torch.from_numpy(img))[0].numpy()
    a = ours[0].argmax(0) == 1
    b = ref[0].argmax(0) == 1
    dice = 2.0 * float((a & b).sum()) / float(a.sum() + b.sum())
    dice_gt = 2.0 * float((a & (lab[0] > 0)).sum()) / float(a.sum() + (lab[0] > 0).sum())
    print(f"label-map Dice vs oracle {dice:.5f} (vs ground truth {dice_gt:.4f}); logits rel err {rel(ours, ref):.2e}")
    assert dice >= 0.999, dice


def test_training_trajectory_tracks_oracle_over_12_steps():
    """12 optimizer steps (batch of 2 micro-batches, poly learning rate, clip, Nesterov momentum, weight decay) with this library
    and with the fp32 CPU oracle from the same initial weights and samples: the logged losses must stay together step by step
    (the golden fixtures cover 2 steps; this covers momentum build-up and the clip factor over a longer run)."""
    m = load()
    W, H, D = 32, 48, 32          # 49152 voxels: the banded kernels are on the path at level 0
    feature = ("conv16,ks3,stride1+norm,leaky_relu+conv16,ks3,stride1+norm,leaky_relu\n"
               "conv32,ks3,stride2+norm,leaky_relu+conv32,ks3,stride1+norm,leaky_relu+conv_trans16,ks2,stride2\n"
               "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
    onet = O.parse_feature(1, 2, feature)
    P = O.init_params(onet, 21)
    net = m.UNet3d(1, 2, feature)
    net.load_parameters([p.numpy() for p in P])
    net.set_dim(W, H, D)
    net.train(True)
    lr0, steps = 1e-2, 12
    net.create_optimizer(lr0)
    samples = [synth_volume(W, H, D, seed=30 + k) for k in range(2)]
    mom = [None] * len(P)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    worst = 0.0
    for s in range(steps):
        lr = m.poly_lr(lr0, s, steps)
        ours = np.zeros(3)
        for img, lab in samples:
            ours += net.train_microbatch(img, lab)
        ours /= len(samples)
        net.step(len(samples), lr)
        assert not net.last_step_skipped()
        logged, _, _ = O.train_step(onet, P, mom, [torch.from_numpy(i) for i, _ in samples], [torch.from_numpy(l).long() for _, l in samples], lr)
        ref = logged.detach().numpy()
        worst = max(worst, float(np.abs(ours - ref).max()))
        assert np.allclose(ours, ref, rtol=0, atol=1e-3), (s, ours, ref)   # measured: 4.6e-5 over the 12 steps
    print("12-step trajectory: max |loss - oracle loss| =", worst, "final", ours, ref)
    assert ours[0] < 0.6 * 0.7          # it actually trained (ce well below ln 2)


def test_validate_uses_running_statistics_like_the_reference():
    """unet3d_validate (train.cpp:826-851) after 3 training steps of the BatchNorm net: the reference binary ran eval() + forward()[0]
    + calc_losses on the first sample (golden f2_validate).  Checks both the handle in training mode (validate must not touch the
    running statistics) and a copy_from replica in eval() mode (the reference's output_model)."""
    m = load()
    z, meta = golden("f2_validate")
    net = build_from_golden(m, z, meta)
    net.train(True)
    net.create_optimizer(meta["lr"])
    x, lab = z["input"][0:1], z["label"][0:1]
    for s in range(meta["steps"]):
        net.train_microbatch(x, lab, meta["collapse"], meta["ce"], meta["dice"], meta["mse"])
        net.step(1, m.poly_lr(meta["lr"], s, meta["total_steps"]))
        assert not net.last_step_skipped()
    v1 = net.validate(x, lab)
    v2 = net.validate(x, lab)
    assert np.array_equal(v1, v2), "validate must not change the running statistics"
    print("validate losses", v1, "reference", z["validate_losses"])
    np.testing.assert_allclose(v1, z["validate_losses"], rtol=0, atol=3e-3)
    twin = m.UNet3d(meta["in_c"], meta["out_c"], str(z["feature"]))
    twin.copy_from(net)            # parameters AND buffers (unet.cpp:195-222)
    twin.eval()
    out0 = twin.forward(x, n_levels=1)[0]
    e = rel(out0, z["validate_logits_0"])
    print("eval() replica logits rel err", e)
    assert e < 1e-2, e
    np.testing.assert_allclose(twin.validate(x, lab), v1, rtol=0, atol=1e-5)
    # prepare_for_inference resets the buffers: the same replica now gives y = gamma*x + beta
    twin.prepare_for_inference()
    out_inf = twin.forward(x, n_levels=1)[0]
    assert rel(out_inf, z["validate_logits_0"]) > 1e-2


def test_validate_default_net_matches_oracle_level0_losses():
    m = load()
    W, H, D = 64, 96, 64
    feature = O.default_feature(2)
    onet = O.parse_feature(1, 2, feature)
    P = O.init_params(onet, 4)
    img, lab = synth_volume(W, H, D, seed=6)
    net = m.UNet3d(1, 2, feature)
    net.load_parameters([p.numpy() for p in P])
    net.set_dim(W, H, D)
    v = net.validate(img, lab)
    with torch.no_grad():
        out0 = O.forward(onet, P, torch.from_numpy(img))[0]
        ref = torch.stack(O.calc_losses(out0, torch.from_numpy(lab).long(), 2, 0)).numpy()
    print("validate", v, "oracle", ref)
    np.testing.assert_allclose(v, ref, rtol=0, atol=2e-3)


def test_validation_replica_runs_beside_the_trainer():
    """train.cpp:773-776, 826-851: the trainer copies its weights into output_model (copy_from) and a second thread validates on that
    replica while training continues.  Two handles on one GPU, two host threads: the asynchronous validation of the replica, started
    before a block of training steps and collected after it, equals a plain validation of the same weights."""
    import threading
    m = load()
    W, H, D = 64, 64, 64
    feature = O.default_feature(2)
    img, lab = synth_volume(W, H, D, seed=8)
    trainer = m.UNet3d(1, 2, feature)
    trainer.init_params(6)
    trainer.set_dim(W, H, D)
    trainer.train(True)
    trainer.create_optimizer(1e-2)
    replica = m.UNet3d(1, 2, feature)
    replica.copy_from(trainer)
    replica.train(False)
    want = replica.validate(img, lab)
    got = {}

    def validator():
        for k in range(4):
            replica.validate_async(img, lab)
            got[k] = replica.validate_result()

    t = threading.Thread(target=validator)
    t.start()
    for s in range(6):                      # the trainer keeps stepping on its own stream meanwhile
        trainer.train_microbatch(img, lab)
        trainer.step(1, 1e-2)
        assert not trainer.last_step_skipped()
    t.join()
    for k in range(4):
        assert np.array_equal(got[k], want), (k, got[k], want)
    # after the next copy_from the replica sees the trained weights
    replica.copy_from(trainer)
    replica.train(False)
    after = replica.validate(img, lab)
    assert after[0] < want[0], (after, want)
