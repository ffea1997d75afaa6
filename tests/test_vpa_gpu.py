"""visual_perception_augmentation: CUDA path (through the C-ABI) against the CPU restatement of the reference's
.cpp (oracle/vpa_oracle.py; parity unpinned — TIPL is not vendored, see the oracle header), plus size-independent
properties at the full 160x192x160 grid."""
import numpy as np
import pytest

from oracle import vpa_oracle as VO
from tests._pkg import load

pytestmark = pytest.mark.gpu


def phantom(W, H, D, C=1, seed=0):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    r = np.sqrt(((z - D / 2) / (0.40 * D)) ** 2 + ((y - H / 2) / (0.38 * H)) ** 2 + ((x - W / 2) / (0.36 * W)) ** 2)
    lab = ((r < 1).astype(np.float32) + (r < 0.6) + (r < 0.3)).astype(np.float32)
    img = np.stack([(np.clip(1.1 - r, 0, 1) * (0.6 + 0.4 * c) + 0.05 * rng.random(r.shape)).astype(np.float32) for c in range(C)])
    img /= img.max()
    return img, lab


def compare(m, options, W, H, D, C, seed, is_label=True):
    img, lab = phantom(W, H, D, C, seed)
    ref_i, ref_l = VO.augment(options, img, lab, is_label, (W, H, D), seed)
    out_i, out_l = m.vpa_augment(img, lab, options, is_label, seed)
    # trilinear samples: fp32 with a different association / FMA contraction; a voxel may differ visibly only where a
    # source position sits within rounding of a cell border, a valid/invalid edge or a majority tie
    bad = np.abs(out_i - ref_i) > 2e-3
    assert bad.mean() < 2e-3, (seed, bad.mean(), np.abs(out_i - ref_i).max())
    assert np.median(np.abs(out_i - ref_i)) < 1e-5
    if is_label:
        assert (out_l != ref_l).mean() < 2e-3, (seed, (out_l != ref_l).mean())
    else:   # the "label" is an intensity image: trilinear like the image channels
        assert (np.abs(out_l - ref_l) > 2e-3).mean() < 2e-3
    return out_i, out_l


@pytest.mark.parametrize("seed", range(8))
def test_default_options_match_oracle(seed):
    m = load()
    compare(m, dict(VO.OPTION_DEFAULTS), 40, 48, 32, 1, seed)


@pytest.mark.parametrize("seed", [3, 11])
def test_everything_on_two_channels(seed):
    m = load()
    o = dict(VO.OPTION_DEFAULTS)
    for k in ("cropping", "truncation_z", "downsample_x", "downsample_y", "downsample_z", "noise", "ambient", "diffuse", "specular",
              "distortion", "rubber_stamping", "perlin_texture"):
        o[k] = 4
    o["zero_background"] = 0
    compare(m, o, 32, 40, 48, 2, seed)


def test_zero_background_and_intensity_label_mode():
    m = load()
    o = dict(VO.OPTION_DEFAULTS)
    o["zero_background"] = 4
    out_i, out_l = compare(m, o, 32, 32, 32, 1, 5)
    assert np.all(out_i[0][out_l == 0] == 0)
    compare(m, dict(VO.OPTION_DEFAULTS), 32, 32, 32, 1, 6, is_label=False)


def test_identity_options_give_identity_warp():
    m = load()
    o = {"scaling_up": 1.0, "scaling_down": 1.0, "aspect_ratio": 1.0}
    img, lab = phantom(24, 32, 16)
    out_i, out_l = m.vpa_augment(img, lab, o, True, 9)
    np.testing.assert_array_equal(out_l, lab)
    np.testing.assert_allclose(out_i, img / img.max(), rtol=0, atol=1e-6)


def test_full_size_properties():
    m = load()
    W, H, D = 160, 192, 160
    img, lab = phantom(W, H, D, 1, 1)
    for seed in (0, 1, 2):
        out_i, out_l = m.vpa_augment(img, lab, None, True, seed)
        assert np.isfinite(out_i).all() and out_i.min() >= 0.0 and out_i.max() <= 1.0 + 1e-6
        if out_i.max() > 0:
            assert abs(out_i.max() - 1.0) < 1e-5
        assert set(np.unique(out_l)).issubset(set(np.unique(lab)))
        again_i, again_l = m.vpa_augment(img, lab, None, True, seed)
        assert np.array_equal(again_i, out_i) and np.array_equal(again_l, out_l)   # same seed -> same sample


def test_fused_augment_microbatch_equals_the_two_calls():
    """unet3d_train_microbatch_augmented (one upload, augmentation and micro-batch in HBM) must give the losses of
    vpa_augment (host, in place) followed by train_microbatch on the augmented host buffers."""
    m = load()
    W, H, D = 64, 48, 32
    feature = ("conv16,ks3,stride1+norm,leaky_relu\nconv32,ks3,stride2+norm,leaky_relu+conv_trans16,ks2,stride2\n"
               "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
    img, lab = phantom(W, H, D, 1, 3)
    lab = np.minimum(lab, 1.0)
    nets = []
    for _ in range(2):
        net = m.UNet3d(1, 2, feature, gpu=0)
        net.init_params(11)
        net.set_dim(W, H, D)
        net.train(True)
        nets.append(net)
    fused = m.train_microbatch_augmented(nets[0], img[None], lab[None], seed=5)
    ai, al = m.vpa_augment(img.copy(), lab.copy(), None, True, 5)
    two, _ = nets[1].train_microbatch(ai[None], al[None], all_levels=True)
    assert np.isfinite(fused).all()
    np.testing.assert_allclose(fused, two, rtol=2e-4, atol=2e-5)
    g0 = nets[0].get_grad(0)
    g1 = nets[1].get_grad(0)
    assert np.linalg.norm(g0 - g1) <= 1e-3 * max(np.linalg.norm(g1), 1e-20)


def test_prefetched_samples_equal_the_fused_call_and_come_out_in_order():
    """unet3d_prefetch_augmented / unet3d_train_microbatch_prefetched (upload + augmentation of the next sample on a side stream,
    two staging slots, first-in first-out) must give the losses of the synchronous fused call for the same samples and seeds."""
    m = load()
    W, H, D = 64, 48, 32
    feature = ("conv16,ks3,stride1+norm,leaky_relu\nconv32,ks3,stride2+norm,leaky_relu+conv_trans16,ks2,stride2\n"
               "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
    samples = []
    for k in range(3):
        img, lab = phantom(W, H, D, 1, 10 + k)
        samples.append((np.ascontiguousarray(img[None]), np.ascontiguousarray(np.minimum(lab, 1.0)[None])))
    nets = []
    for _ in range(2):
        net = m.UNet3d(1, 2, feature, gpu=0)
        net.init_params(12)
        net.set_dim(W, H, D)
        net.train(True)
        net.create_optimizer(1e-2)
        nets.append(net)
    ref = []
    for k, (img, lab) in enumerate(samples):
        ref.append(m.train_microbatch_augmented(nets[0], img, lab, seed=20 + k))
        nets[0].step(1, 1e-2)
    got = []
    m.prefetch_augmented(nets[1], samples[0][0], samples[0][1], seed=20)
    for k in range(3):
        if k + 1 < 3:
            m.prefetch_augmented(nets[1], samples[k + 1][0], samples[k + 1][1], seed=20 + k + 1)
        got.append(m.train_microbatch_prefetched(nets[1]))
        nets[1].step(1, 1e-2)
    np.testing.assert_allclose(np.array(got), np.array(ref), rtol=5e-4, atol=5e-5)
    with pytest.raises(m.U3DError, match="no prefetched sample"):
        m.train_microbatch_prefetched(nets[1])


def test_mt19937_noise_stream_is_the_reference_cpu_stream():
    """Library option noise_mt19937 = 1: the noise added per voxel is the reference CPU path's sequential stream
    (visual_perception_augmentation.cpp:252-258: one uniform_dist<float>(0, noise_mag, seed) drawn voxel after voxel, channel after
    channel) generated on the GPU in 624-word blocks.  With every other stage off and the identity warp the output is
    (img + noise) / max per channel, so the stream is checked bit for bit up to the final normalisation."""
    m = load()
    W, H, D, C = 24, 20, 16, 2          # 7680 voxels per channel: 12.3 generator blocks per channel, the stream crosses the channel border
    img, lab = phantom(W, H, D, C, seed=2)
    o = {"scaling_up": 1.0, "scaling_down": 1.0, "aspect_ratio": 1.0, "noise": 4, "noise_mag": 0.2, "noise_mt19937": 1}
    seed = 77
    out_i, out_l = m.vpa_augment(img, lab, o, True, seed)
    noise = VO.noise_field_mt19937(seed, C * W * H * D, 0.2).reshape(C, D, H, W)
    want = img + noise
    want = np.stack([w / w.max() for w in want]).astype(np.float32)
    np.testing.assert_allclose(out_i, want, rtol=0, atol=1.2e-7)
    np.testing.assert_array_equal(out_l, lab)
    ref_i, _ = VO.augment(o, img, lab, True, (W, H, D), seed)
    np.testing.assert_allclose(out_i, ref_i, rtol=0, atol=1e-6)
    # and the default (hash) stream is a different one
    o2 = dict(o); o2["noise_mt19937"] = 0
    assert np.abs(m.vpa_augment(img, lab, o2, True, seed)[0] - out_i).max() > 1e-3
