"""simulate_modality (train.cpp:43-180): CUDA path through the C-ABI against the CPU restatement oracle/simulate_oracle.py
(parity unpinned — TIPL is not vendored, see the oracle header), plus size-independent properties at the full grid."""
import numpy as np
import pytest

from oracle import simulate_oracle as SO
from tests._pkg import load
from tests.test_vpa_gpu import phantom

pytestmark = pytest.mark.gpu

# everything up to pow() is bit-exact (same fp32 evaluation order, no FMA contraction); CUDA powf and the host's powf differ by a few
# ulp, and the (v - min) / (max - min) rescale keeps that at the 1e-6 level on the [0,1] output
ATOL = 5e-6


@pytest.mark.parametrize("seed", [0, 1, 7, 123456789])
def test_labelled_overload_matches_oracle(seed):
    m = load()
    img, lab = phantom(40, 48, 32, 1, seed % 97)
    ref = SO.simulate_modality(img[0], lab, 3, seed)
    out = m.simulate_modality(img[0], lab, 3, seed)
    assert np.abs(out - ref).max() <= ATOL, np.abs(out - ref).max()
    assert ((out == 0) == (ref == 0)).mean() > 0.9999


@pytest.mark.parametrize("seed", [0, 2, 31, 4000000000])
def test_image_only_overload_matches_oracle(seed):
    m = load()
    img, _ = phantom(36, 44, 28, 1, seed % 89)   # odd-ish sizes: the star's linear offsets wrap across rows and planes
    ref = SO.simulate_modality(img[0], None, 0, seed)
    out = m.simulate_modality(img[0], None, 0, seed)
    assert np.abs(out - ref).max() <= ATOL, np.abs(out - ref).max()


@pytest.mark.parametrize("labelled", [True, False])
def test_width_not_a_multiple_of_four_takes_the_scalar_star(labelled):
    m = load()
    img, lab = phantom(30, 20, 12, 1, 6)
    ref = SO.simulate_modality(img[0], lab if labelled else None, 3, 17)
    out = m.simulate_modality(img[0], lab if labelled else None, 3, 17)
    assert np.abs(out - ref).max() <= ATOL, np.abs(out - ref).max()


def test_uniform_image_is_left_unscaled():
    """max == min over the selected voxels: the reference skips the rescale (train.cpp:111)."""
    m = load()
    img = np.full((8, 8, 8), 0.5, np.float32)
    lab = np.zeros((8, 8, 8), np.float32)   # no labelled voxel at all: min / max stay at their sentinels
    ref = SO.simulate_modality(img, lab, 1, 3)
    out = m.simulate_modality(img, lab, 1, 3)
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=0)
    assert out.max() > 0


def test_full_size_properties():
    m = load()
    W, H, D = 160, 192, 160
    img, lab = phantom(W, H, D, 1, 1)
    out = m.simulate_modality(img[0], lab, 3, 42)
    again = m.simulate_modality(img[0], lab, 3, 42)
    assert np.array_equal(out, again)                      # deterministic
    assert out.min() >= 0.0 and out.max() <= 1.0
    assert (out[img[0] <= 0.02] == 0).all()                # train.cpp:87-92
    sel = (img[0] > 0.02) & (lab != 0)
    assert out[sel].max() == 1.0 and out[sel].min() == 0.0   # rescaled to the labelled voxels' range
    other = m.simulate_modality(img[0], lab, 3, 43)
    assert not np.array_equal(out, other)


def test_errors():
    m = load()
    img = np.full((8, 8, 8), 0.5, np.float32)
    with pytest.raises(m.U3DError, match="max_label"):
        m.simulate_modality(img, img, 100000, 0)
    feature = ("conv16,ks3,stride1+norm,leaky_relu\nconv32,ks3,stride2+norm,leaky_relu+conv_trans16,ks2,stride2\n"
               "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
    net = m.UNet3d(2, 2, feature, gpu=0)
    with pytest.raises(m.U3DError, match="mode must be"):
        m.set_simulate_modality(net, 3)
    net.init_params(1)
    net.set_dim(32, 32, 32)
    net.train(True)
    m.set_simulate_modality(net, 1)
    x = np.zeros((1, 2, 32, 32, 32), np.float32)
    with pytest.raises(m.U3DError, match="single-channel"):
        m.train_microbatch_augmented(net, x, np.zeros((1, 32, 32, 32), np.float32), seed=0)


@pytest.mark.parametrize("mode", [1, 2])
def test_fused_sample_path_runs_simulate_before_augmentation(mode):
    """set_simulate_modality(mode) + train_microbatch_augmented == simulate_modality on the host arrays, then the fused call with it off."""
    m = load()
    W, H, D = 64, 48, 32
    feature = ("conv16,ks3,stride1+norm,leaky_relu\nconv32,ks3,stride2+norm,leaky_relu+conv_trans16,ks2,stride2\n"
               "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
    img, lab = phantom(W, H, D, 1, 4)
    lab = np.minimum(lab, 1.0)
    nets = []
    for _ in range(2):
        net = m.UNet3d(1, 2, feature, gpu=0)
        net.init_params(13)
        net.set_dim(W, H, D)
        net.train(True)
        nets.append(net)
    m.set_simulate_modality(nets[0], mode)
    fused = m.train_microbatch_augmented(nets[0], img[None], lab[None], seed=9)
    sim = m.simulate_modality(img[0], lab if mode == 1 else None, 2, 9)   # max_label = out_count (train.cpp:459)
    two = m.train_microbatch_augmented(nets[1], sim[None, None], lab[None], seed=9)
    assert np.isfinite(fused).all()
    np.testing.assert_array_equal(fused, two)
    plain = m.train_microbatch_augmented(nets[1], img[None], lab[None], seed=9)
    assert not np.array_equal(plain, two)
    # the prefetch path takes the same stage
    nets[0].create_optimizer(1e-2); nets[0].step(1, 0.0)
    nets[1].create_optimizer(1e-2); nets[1].step(1, 0.0)
    a = np.ascontiguousarray(img[None]); b = np.ascontiguousarray(lab[None])
    m.prefetch_augmented(nets[0], a, b, seed=9)
    pf = m.train_microbatch_prefetched(nets[0])
    np.testing.assert_array_equal(pf, fused)
