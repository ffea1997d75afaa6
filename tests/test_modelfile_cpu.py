"""The .nz container (gzip stream of MATLAB Level-4 MAT matrices; /root/reference/main.cpp:157-233 through TIPL's gz_mat_read /
gz_mat_write) checked against an independent implementation of the public Level-4 format: scipy.io.  TIPL itself is not vendored,
so its dialect (the "sloped" tensor encoding in particular) stays parity-unpinned; see csrc/modelfile.cpp."""
import gzip
import io
import os

import numpy as np
import pytest
import scipy.io

from tests._pkg import load


def test_written_container_is_read_by_scipy(tmp_path):
    m = load()
    z = m.NzFile()
    rng = np.random.default_rng(0)
    t0 = rng.standard_normal((27, 16)).astype(np.float32)          # rows = numel/size(0), cols = size(0)
    z.add("channels", np.array([1, 2], np.int32))
    z.add("architecture", "conv16,ks3,stride1+norm,leaky_relu\nconv2,ks1,stride1")
    z.add("voxel_size", np.array([1.0, 0.5, 0.25], np.float32))
    z.add("tensor0", t0)
    z.add("training_errors", np.arange(12, dtype=np.float32).reshape(3, 4))
    path = tmp_path / "m.nz"
    z.save(path)
    raw = gzip.open(path, "rb").read()
    d = scipy.io.loadmat(io.BytesIO(raw))
    assert d["channels"].tolist() == [[1, 2]]
    assert str(d["architecture"][0]) == "conv16,ks3,stride1+norm,leaky_relu\nconv2,ks1,stride1"
    np.testing.assert_array_equal(d["tensor0"], t0)
    np.testing.assert_array_equal(d["training_errors"], np.arange(12, dtype=np.float32).reshape(3, 4))
    # native element order of parameters()[i] ([cols][rows] contiguous) is what sits in the file
    names = [n for n, *_ in z.names()]
    assert names == ["channels", "architecture", "voxel_size", "tensor0", "training_errors"]


def test_scipy_written_container_is_read(tmp_path):
    m = load()
    rng = np.random.default_rng(1)
    t = rng.standard_normal((8, 3))
    q = rng.integers(-3000, 3000, (5, 4)).astype(np.int16)
    b = io.BytesIO()
    scipy.io.savemat(b, {"tensor0": t, "tensor1": q, "tensor1.slope": np.array([[0.5, 1.0, 2.0, 4.0]], np.float32),
                         "tensor1.inter": np.array([[1.0]], np.float32), "architecture": "abc\ndef", "channels": np.array([[3, 4]], np.int32)}, format="4")
    path = tmp_path / "s.nz"
    with gzip.open(path, "wb") as f:
        f.write(b.getvalue())
    z = m.NzFile(path)
    info = {n: (ty, r, c) for n, ty, r, c in z.names()}
    assert info["tensor0"] == (0, 8, 3) and info["tensor1"] == (30, 5, 4) and info["architecture"][0] == 51
    np.testing.assert_array_equal(z.read_f32("tensor0"), t.astype(np.float32))
    # assumed sloped companions: one slope per column, scalar intercept
    np.testing.assert_allclose(z.read_f32("tensor1"), q * np.array([0.5, 1.0, 2.0, 4.0], np.float32) + 1.0, rtol=1e-6)
    # an uncompressed Level-4 file loads too (zlib's transparent read)
    raw = tmp_path / "plain.mat"
    raw.write_bytes(b.getvalue())
    assert len(m.NzFile(raw).names()) == 6


def test_broken_files_fail_loudly(tmp_path):
    m = load()
    with pytest.raises(m.U3DError, match="cannot open"):
        m.NzFile(tmp_path / "missing.nz")
    z = m.NzFile()
    z.add("tensor0", np.ones((4, 4), np.float32))
    p = tmp_path / "t.nz"
    z.save(p)
    raw = gzip.open(p, "rb").read()
    cut = tmp_path / "cut.nz"
    with gzip.open(cut, "wb") as f:
        f.write(raw[:-10])
    with pytest.raises(m.U3DError, match="truncated"):
        m.NzFile(cut)
    junk = tmp_path / "junk.nz"
    junk.write_bytes(os.urandom(256))
    with pytest.raises(m.U3DError):
        m.NzFile(junk)
    with pytest.raises(m.U3DError, match="Level-4"):
        m.load_from_file(junk)
