"""Inference pre/post-processing on the GPU against oracle/postproc_oracle.py (parity unpinned: the reference's
implementation lives in the un-vendored TIPL, see the oracle header): softmax + create_mask + argmax, window cutting and re-assembly
of a volume that is larger / smaller than the model grid, and resampling."""
import numpy as np
import pytest

from oracle import postproc_oracle as PO
from tests._pkg import load

pytestmark = pytest.mark.gpu

FEATURE = ("conv8,ks3,stride1+norm,leaky_relu\nconv16,ks3,stride2+norm,leaky_relu+conv_trans8,ks2,stride2\n"
           "conv8,ks3,stride1+norm,leaky_relu+conv4,ks1,stride1")


def test_softmax_create_mask_argmax():
    m = load()
    rng = np.random.default_rng(0)
    for C, shape in ((4, (9, 11, 13)), (6, (16, 8, 8)), (2, (5, 5, 5)), (12, (4, 6, 8))):
        logits = (rng.standard_normal((C,) + shape) * 3).astype(np.float32)
        logits[:, 0, 0, 0] = 1.0                      # tie: first arg-max wins, fg = 1 - 1/C
        label, fg, prob = m.postproc(logits, 0.5)
        p = PO.softmax(logits)
        rl, rf = PO.mask_argmax(p, 0.5)
        np.testing.assert_allclose(prob, p, rtol=0, atol=2e-6)
        np.testing.assert_allclose(fg, rf, rtol=0, atol=2e-6)
        near = np.abs(rf - 0.5) < 1e-5               # voxels whose mask decision sits within rounding of the threshold
        assert np.array_equal(label[~near], rl[~near])
        assert label[0, 0, 0] == (0 if 1 - 1 / C <= 0.5 else 0)


@pytest.mark.parametrize("vol_shape,stride", [((40, 48, 56), (0, 0, 0)), ((40, 48, 56), (16, 16, 16)), ((24, 20, 30), (0, 0, 0)), ((32, 32, 32), (0, 0, 0))])
def test_evaluate_volume_windows_and_reassembly(vol_shape, stride):
    m = load()
    net = m.UNet3d(1, 4, FEATURE)
    net.init_params(2)
    net.set_dim(32, 32, 32)
    net.prepare_for_inference()
    rng = np.random.default_rng(1)
    vol = rng.random((1,) + vol_shape, dtype=np.float32)
    label, fg, prob, n = m.evaluate_volume(net, vol, stride, 0.5, want_prob=True)

    def fwd(win):
        return net.forward(win[None], n_levels=1)[0][0]

    rl, rf, rp, rn = PO.evaluate_volume(fwd, vol, (32, 32, 32), stride, 0.5)
    assert n == rn
    np.testing.assert_allclose(prob, rp, rtol=0, atol=3e-6)
    np.testing.assert_allclose(fg, rf, rtol=0, atol=3e-6)
    near = np.abs(rf - 0.5) < 1e-5
    top2 = np.sort(rp, 0)[-2:]
    near |= (top2[1] - top2[0]) < 1e-5
    assert np.array_equal(label[~near], rl[~near])
    D, H, W = vol_shape
    for dim, s, got in ((W, stride[0], None), (H, stride[1], None), (D, stride[2], None)):
        assert m.window_origins(dim, 32, s) == PO.window_origins(dim, 32, s)


def test_resample_linear_and_nearest():
    m = load()
    rng = np.random.default_rng(3)
    src = rng.random((2, 20, 24, 28), dtype=np.float32)
    for dst in ((32, 32, 32), (10, 12, 14), (20, 24, 28)):
        np.testing.assert_allclose(m.resample(src, dst), PO.resample(src, dst), rtol=0, atol=2e-6)
        lab = np.floor(src * 5).astype(np.float32)
        np.testing.assert_array_equal(m.resample(lab, dst, nearest=True), PO.resample(lab, dst, nearest=True))
    np.testing.assert_array_equal(m.resample(src, (20, 24, 28)), src)
