"""CPU checks of the simulate_modality restatement (oracle/simulate_oracle.py): the generators against the C++ standard's
known answers, and the properties train.cpp:43-180 implies.  The TIPL-side semantics stay unpinned (see the oracle header)."""
import numpy as np

from oracle import simulate_oracle as SO
from oracle.vpa_oracle import MT19937


def test_mt19937_known_answer():
    g = MT19937(5489)
    for _ in range(9999):
        g()
    assert g() == 4123659995          # [rand.predef]: the 10000th invocation of a default-constructed std::mt19937


def test_int_draws_are_the_top_bits_for_powers_of_two():
    a, g = SO.UniformInt(77), MT19937(77)
    assert [a(4) for _ in range(64)] == [g() >> 30 for _ in range(64)]
    b = SO.UniformInt(5)
    assert all(0 <= b(3) < 3 for _ in range(200))


def test_terms_never_have_a_equal_b_equal_zero():
    for seed in range(20):
        terms = SO.draw_terms(SO.UniformInt(seed), SO.UniformDist(0.0, 1.0, seed + 1))
        assert len(terms) == SO.TERM_COUNT
        assert all(a + b != 0 and 0 <= c < 4 and 0 <= d < 4 and 0 <= w < 1 for a, b, c, d, w in terms)


def test_star_filter():
    img = np.ones((6, 5, 4), np.float32)
    out = SO.gaussian(img)
    assert (out[1:-1] == 1).all()                      # all 8/8 of the weight present away from the volume's two ends
    assert out.reshape(-1)[0] == np.float32(5 / 8)     # the first voxel has no -1, -W, -WH neighbour
    assert out.reshape(-1)[-1] == np.float32(5 / 8)
    imp = np.zeros((5, 5, 5), np.float32); imp[2, 2, 2] = 8
    out = SO.gaussian(imp)
    assert out[2, 2, 2] == 2 and out[2, 2, 1] == 1 and out[1, 2, 2] == 1 and out[2, 1, 2] == 1 and out.sum() == 8
    edge = np.zeros((3, 3, 4), np.float32); edge[0, 0, 3] = 8     # the +1 neighbour of a row's last voxel is the next row's first
    assert SO.gaussian(edge)[0, 1, 0] == 1


def test_properties():
    rng = np.random.default_rng(0)
    z, y, x = np.meshgrid(np.arange(16), np.arange(20), np.arange(24), indexing="ij")
    r = np.sqrt(((z - 8) / 7) ** 2 + ((y - 10) / 9) ** 2 + ((x - 12) / 11) ** 2)
    img = (np.clip(1.05 - r, 0, 1) + 0.02 * rng.random(r.shape)).astype(np.float32)
    img /= img.max()
    lab = ((r < 1).astype(np.float32) + (r < 0.5)).astype(np.float32)
    for seed in (0, 5):
        tr = {}
        out = SO.simulate_modality(img, lab, 2, seed, trace=tr)
        assert out.dtype == np.float32 and out.min() >= 0 and out.max() <= 1
        assert (out[img <= 0.02] == 0).all()
        sel = (img > 0.02) & (lab != 0)
        assert out[sel].max() == 1 and out[sel].min() == 0
        assert 0.6 <= tr["gamma"] < 1.8 and (tr["tissue"] <= 0.6).all()
        out2 = SO.simulate_modality(img, None, 0, seed)
        assert out2.min() >= 0 and out2.max() == 1 and (out2[img <= 0.02] == 0).all()
    assert not np.array_equal(SO.simulate_modality(img, lab, 2, 0), SO.simulate_modality(img, lab, 2, 1))


def test_host_plan_of_the_library_draws_what_the_oracle_draws():
    """simulate_modality_plan (host only, no GPU): tissue LUT, the 20 terms and gamma for several seeds and both overloads."""
    import ctypes
    from tests._pkg import load
    m = load()
    L = m.lib()
    img = np.full((2, 2, 2), 0.5, np.float32)
    lab = np.ones((2, 2, 2), np.float32)
    FP = ctypes.POINTER(ctypes.c_float)
    for seed in (0, 1, 77, 4000000000, 0xFFFFFFFF):
        for labelled, max_label in ((1, 3), (1, 40), (0, 0)):
            tr = {}
            SO.simulate_modality(img, lab if labelled else None, max_label, seed, trace=tr)
            lut = np.zeros(max_label + 1, np.float32)
            terms = np.zeros((20, 5), np.float32)
            gamma = ctypes.c_float(0)
            rc = L.simulate_modality_plan(labelled, ctypes.c_uint(max_label), ctypes.c_uint(seed), lut.ctypes.data_as(FP),
                                          terms.ctypes.data_as(FP), ctypes.byref(gamma))
            assert rc == 0
            if labelled:
                assert np.array_equal(lut, tr["lut"])
            assert np.array_equal(terms, np.array([[a, b, c, d, w] for a, b, c, d, w in tr["terms"]], np.float32))
            assert np.float32(gamma.value) == tr["gamma"]
    with np.testing.assert_raises(m.U3DError):
        m.check(L.simulate_modality_plan(1, ctypes.c_uint(100000), ctypes.c_uint(0), None, None, None))
