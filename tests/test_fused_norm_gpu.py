"""Norm + activation folded into the output head (SrcTransform): head_fwd_kernel applies gamma*rstd*(x - mean) + beta
and the activation of unet.cpp:74-98 itself when it is the only reader of the activated tensor, and the separate norm_act_fwd pass over
that full-resolution tensor disappears.  The arithmetic is norm_act_fwd_kernel's, operation for operation, so the folded path must give
BIT-identical forward results to the unfolded one (U3D_NO_XF=1, read once per process -> child process)."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._pkg import load

pytestmark = pytest.mark.gpu

W, H, D = 64, 64, 64
NETS = {
    # the default architecture (InstanceNorm3d + LeakyReLU, skips, stride-2 down, transpose-conv up)
    "default": (1, 3, None),
    # BatchNorm3d + ReLU / ELU, max_pool / upsample: batch statistics in training, running statistics in eval
    "bn_relu_elu": (2, 2, "conv16,ks3,stride1+bnorm,relu+conv16,ks3,stride1+bnorm,elu\n"
                          "max_pool+conv32,ks3,stride1+bnorm,relu+upsample\n"
                          "conv16,ks3,stride1+bnorm,relu+conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1"),
}


def _feature(m, name):
    in_c, out_c, feat = NETS[name]
    if feat is None:
        return in_c, out_c, None
    return in_c, out_c, feat


def compute(name):
    m = load()
    in_c, out_c, feat = _feature(m, name)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((1, in_c, D, H, W)).astype(np.float32)
    lab = rng.integers(0, out_c, size=(1, 1, D, H, W)).astype(np.float32)
    out = {}
    # inference with the freshly initialised weights (identical in both processes)
    inf = m.UNet3d(in_c, out_c, feat)
    inf.init_params(3)
    inf.set_dim(W, H, D)
    inf.prepare_for_inference()
    l0 = inf.launch_count()
    out["logits_inf"] = inf.forward(x, n_levels=1)[0]
    out["inf_launches"] = np.array([inf.launch_count() - l0])
    del inf
    net = m.UNet3d(in_c, out_c, feat)
    net.init_params(3)
    net.set_dim(W, H, D)
    net.train(True)
    net.create_optimizer(1e-2)
    losses = []
    for it in range(2):
        losses.append(np.asarray(net.train_microbatch(x, lab), np.float64))
        if it == 0:
            out["grad"] = np.concatenate([net.get_grad(i).ravel() for i in range(net.param_count())])
        net.step(1, 1e-2)
    out["losses"] = np.stack(losses)
    out["train_launches"] = np.array([net.launch_count()])
    out["val"] = np.asarray(net.validate(x, lab), np.float64)
    params = [net.get_param(i) for i in range(net.param_count())]
    inf = m.UNet3d(in_c, out_c, feat)
    inf.load_parameters(params)
    inf.set_dim(W, H, D)
    inf.eval()
    out["logits_eval"] = inf.forward(x, n_levels=1)[0]
    return out


@pytest.mark.parametrize("name", list(NETS))
def test_folded_norm_is_bit_identical_to_the_separate_pass(name, tmp_path):
    if os.environ.get("U3D_NO_XF"):
        pytest.skip("the folded path is switched off in this process")
    got = compute(name)
    ref_file = str(tmp_path / "ref.npz")
    env = dict(os.environ, U3D_NO_XF="1")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), name, ref_file], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref = np.load(ref_file)
    # the fold removes launches: that it is active at all
    assert int(got["inf_launches"][0]) < int(ref["inf_launches"][0]), (got["inf_launches"], ref["inf_launches"])
    assert int(got["train_launches"][0]) < int(ref["train_launches"][0])
    # forward: bit-identical (inference, BatchNorm eval with running statistics, training losses of the first micro-batch)
    assert np.array_equal(got["logits_inf"], ref["logits_inf"])
    assert np.array_equal(got["losses"][0], ref["losses"][0]), (got["losses"], ref["losses"])
    # backward: same operands, the weight-gradient atomics reorder sums (bounded in test_baseline_configs_gpu.py at 1e-5)
    g, q = got["grad"].astype(np.float64), ref["grad"].astype(np.float64)
    rel = np.linalg.norm(g - q) / np.linalg.norm(q)
    print(f"{name}: folded vs separate: gradient rel diff {rel:.2e}, launches inference {int(got['inf_launches'][0])} vs "
          f"{int(ref['inf_launches'][0])}, 2 training steps {int(got['train_launches'][0])} vs {int(ref['train_launches'][0])}")
    assert rel < 1e-5
    # after one update (second micro-batch, validation, eval-mode forward) only that reordering separates the two
    assert np.allclose(got["losses"][1], ref["losses"][1], rtol=2e-4, atol=2e-5)
    assert np.allclose(got["val"], ref["val"], rtol=2e-4, atol=2e-5)
    assert np.allclose(got["logits_eval"], ref["logits_eval"], rtol=0, atol=5e-3)


if __name__ == "__main__":
    res = compute(sys.argv[1])
    np.savez(sys.argv[2], **res)
