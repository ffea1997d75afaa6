"""Host-side plan of visual_perception_augmentation (vpa_plan_describe, no GPU needed) against the scalars the CPU restatement draws
(oracle/vpa_oracle.py, trace): the affine, the perspective coefficients and the distortion foci depend on EVERY earlier draw, so this
pins the library's draw order to the oracle's (which restates visual_perception_augmentation.cpp:180-320; TIPL semantics unpinned)."""
import ctypes

import numpy as np
import pytest

from oracle import vpa_oracle as VO
from tests._pkg import load

FP = ctypes.POINTER(ctypes.c_float)


def describe(m, options, W, H, D, C, seed, is_label=1):
    keys = [k.encode() for k in options]
    karr = (ctypes.c_char_p * len(keys))(*keys)
    varr = (ctypes.c_float * len(keys))(*[float(v) for v in options.values()])
    M = np.zeros(12, np.float32); persp = np.zeros(3, np.float32); foci = np.zeros((12, 5), np.float32)
    n = ctypes.c_int(0)
    m.check(m.lib().vpa_plan_describe(karr, varr, len(keys), is_label, W, H, D, C, ctypes.c_uint64(seed), M.ctypes.data_as(FP),
                                      persp.ctypes.data_as(FP), ctypes.byref(n), foci.ctypes.data_as(FP)))
    return M.reshape(3, 4), persp, foci[:n.value]


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 8, 13, 21])
def test_plan_matches_oracle_draws(seed):
    m = load()
    W, H, D = 20, 24, 16
    rng = np.random.default_rng(seed)
    img = rng.random((1, D, H, W), dtype=np.float32)
    lab = (rng.random((D, H, W)) > 0.5).astype(np.float32)
    options = dict(VO.OPTION_DEFAULTS)
    if seed % 2:   # every optional stage on: more draws in front of the affine
        for k in ("cropping", "truncation_z", "downsample_x", "downsample_y", "downsample_z", "noise", "ambient", "diffuse", "specular",
                  "distortion", "rubber_stamping", "perlin_texture"):
            options[k] = 4
    tr = {}
    VO.augment(options, img, lab, True, (W, H, D), seed, trace=tr)
    M, persp, foci = describe(m, options, W, H, D, 1, seed)
    np.testing.assert_allclose(M, tr["M"], rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(persp, np.array(tr["persp"], np.float32), rtol=1e-6, atol=1e-9)
    assert len(foci) == len(tr["foci"])
    for got, (loc, radius, mag) in zip(foci, tr["foci"]):
        assert tuple(int(v) for v in got[:3]) == tuple(int(v) for v in loc)
        assert got[3] == np.float32(radius) and got[4] == np.float32(mag)


def test_unknown_keys_read_as_zero_and_channel_limit():
    m = load()
    M, persp, foci = describe(m, {"not_an_option": 3.0}, 16, 16, 16, 1, 4)
    assert len(foci) == 0 and (persp == 0).all()
    with pytest.raises(m.U3DError, match="channels"):
        describe(m, {}, 16, 16, 16, 9, 0)


@pytest.mark.parametrize("seed", [1, 3, 7, 12345])
def test_perlin_permutation_is_std_shuffle(seed):
    """visual_perception_augmentation.cpp:388-392 shuffles the Perlin table with std::shuffle(p, std::mt19937(seed)).  The host plan
    calls the real std::shuffle; the oracle restates libstdc++'s algorithm (pairs of swap positions per draw, Lemire's bounded
    integers).  Both must produce the same permutation and the same zoom draw behind it."""
    m = load()
    W, H, D = 20, 24, 16
    rng = np.random.default_rng(seed)
    img = rng.random((1, D, H, W), dtype=np.float32)
    lab = (rng.random((D, H, W)) > 0.5).astype(np.float32)
    options = dict(VO.OPTION_DEFAULTS)
    options["perlin_texture"] = 4
    tr = {}
    VO.augment(options, img, lab, True, (W, H, D), seed, trace=tr)
    keys = [k.encode() for k in options]
    karr = (ctypes.c_char_p * len(keys))(*keys)
    varr = (ctypes.c_float * len(keys))(*[float(v) for v in options.values()])
    perm = (ctypes.c_int * 512)()
    applies = ctypes.c_int(0)
    zoom = ctypes.c_float(0)
    m.check(m.lib().vpa_plan_perlin(karr, varr, len(keys), 1, W, H, D, 1, ctypes.c_uint64(seed), ctypes.byref(applies), perm,
                                    ctypes.byref(zoom)))
    assert applies.value == 1
    assert list(perm) == tr["perlin_perm"]
    assert sorted(perm) == sorted(i & 255 for i in range(512))
    assert zoom.value == np.float32(tr["perlin_zoom"])


def test_oracle_mt19937_words_and_uniform_float_match_the_scalar_restatement():
    """noise_field_mt19937 takes its raw words from numpy's MT19937 (legacy init_genrand seeding); they must be std::mt19937's
    words (restated scalar class MT19937, known answer: the 10000th output of mt19937(5489) is 4123659995, ISO C++ [rand.predef])."""
    g = VO.MT19937(5489)
    for _ in range(9999):
        g()
    assert g() == 4123659995
    for seed in (0, 1, 77, 2 ** 31 + 5):
        ud = VO.UniformDist(0.0, 0.2, seed)
        want = np.array([ud() for _ in range(1500)], np.float32)     # crosses two 624-word regenerations
        np.testing.assert_array_equal(VO.noise_field_mt19937(seed, 1500, 0.2), want)
