"""load_from_file / save_to_file (main.cpp:157-233) round trips through the C-ABI, a model file written by an independent Level-4
writer (scipy.io), the reference's error texts, the optimizer file and the raw export."""
import gzip
import io
import json
import os

import numpy as np
import pytest
import scipy.io

from tests._pkg import load

pytestmark = pytest.mark.gpu

FEATURE = ("conv8,ks3,stride1+norm,leaky_relu\nconv16,ks3,stride2+bnorm,relu+conv_trans8,ks2,stride2\n"
           "conv8,ks3,stride1+norm,leaky_relu+conv3,ks1,stride1")


def test_save_load_round_trip(tmp_path):
    m = load()
    a = m.UNet3d(2, 3, FEATURE)
    a.init_params(7)
    a.set_dim(32, 48, 16)
    a.set_info("preproc", "gaussian_smoothing")
    a.set_info("orientation", "LPS")
    a.set_errors(False, np.arange(15, dtype=np.float32).reshape(5, 3))
    a.set_errors(True, np.arange(15, dtype=np.float32).reshape(5, 3) * 2)
    path = tmp_path / "model.nz"
    m.save_to_file(a, path)
    b = m.load_from_file(path)
    assert (b.in_count, b.out_count, b.architecture, b.dim) == (2, 3, FEATURE, (32, 48, 16))
    assert b.param_count() == a.param_count()
    for i in range(a.param_count()):
        assert np.array_equal(a.get_param(i), b.get_param(i)), a.param_name(i)
    assert b.get_info("preproc") == "gaussian_smoothing" and b.get_info("orientation") == "LPS"
    assert b.get_info("postproc") == "softmax+create_mask+argmax" and b.get_info("fov_strategy") == "align_top"   # unet.cpp:110-112
    np.testing.assert_array_equal(b.get_errors(True), a.get_errors(True))
    np.testing.assert_array_equal(b.get_errors(False), a.get_errors(False))
    x = np.random.default_rng(0).random((1, 2, 16, 48, 32), dtype=np.float32)
    a.prepare_for_inference(); b.prepare_for_inference()
    assert np.array_equal(a.forward(x, n_levels=1)[0], b.forward(x, n_levels=1)[0])
    # the file holds tensor{i} as rows = numel/size(0) x cols = size(0) in native element order (main.cpp:225-231)
    d = scipy.io.loadmat(io.BytesIO(gzip.open(path, "rb").read()))
    w0 = a.get_param(0)
    assert d["tensor0"].shape == (w0.size // w0.shape[0], w0.shape[0])
    np.testing.assert_array_equal(d["tensor0"].T.reshape(w0.shape), w0)
    assert d["channels"].tolist() == [[2, 3]] and d["dimension"].tolist() == [[32, 48, 16]]
    assert d["training_errors"].shape == (3, 5)


def test_model_file_from_an_independent_writer_loads(tmp_path):
    m = load()
    ref = m.UNet3d(1, 2, FEATURE.replace("conv3,ks1", "conv2,ks1"))
    ref.init_params(3)
    mats = {"channels": np.array([[1, 2]], np.int32), "architecture": ref.architecture, "dimension": np.array([[16, 16, 32]], np.int32),
            "voxel_size": np.array([[1.0, 1.0, 1.5]], np.float32), "postproc": "softmax+argmax"}
    for i in range(ref.param_count()):
        p = ref.get_param(i)
        mats[f"tensor{i}"] = p.reshape(p.shape[0], -1).T.astype(np.float64)     # a double matrix: read_as_type<float> converts
    b = io.BytesIO()
    scipy.io.savemat(b, mats, format="4")
    path = tmp_path / "indep.nz"
    with gzip.open(path, "wb") as f:
        f.write(b.getvalue())
    net = m.load_from_file(path)
    assert net.dim == (16, 16, 32) and net.get_info("postproc") == "softmax+argmax"
    for i in range(ref.param_count()):
        assert np.array_equal(net.get_param(i), ref.get_param(i))
    # tensor size mismatch (main.cpp:199-201) and missing structure (main.cpp:166)
    mats["tensor2"] = np.zeros((3, 3))
    b = io.BytesIO(); scipy.io.savemat(b, mats, format="4")
    bad = tmp_path / "bad.nz"
    with gzip.open(bad, "wb") as f:
        f.write(b.getvalue())
    with pytest.raises(m.U3DError, match="tensor size mismatch at tensor2 9 not the expected of size 8"):
        m.load_from_file(bad)
    del mats["architecture"]
    b = io.BytesIO(); scipy.io.savemat(b, mats, format="4")
    with gzip.open(bad, "wb") as f:
        f.write(b.getvalue())
    with pytest.raises(m.U3DError, match="invalid format"):
        m.load_from_file(bad)


def test_optimizer_file_and_raw_export(tmp_path):
    m = load()
    a = m.UNet3d(1, 2, FEATURE.replace("conv3,ks1", "conv2,ks1"))
    a.init_params(1)
    a.set_dim(16, 16, 16)
    a.train(True)
    a.create_optimizer(1e-2)
    rng = np.random.default_rng(2)
    x = rng.random((1, 1, 16, 16, 16), dtype=np.float32)
    lab = (rng.random((1, 16, 16, 16)) > 0.5).astype(np.float32)
    for s in range(2):
        a.train_microbatch(x, lab)
        a.step(1, 1e-2)
    m.save_to_file(a, tmp_path / "ck.nz")
    m.save_optimizer(a, tmp_path / "ck.nz.opt")
    b = m.load_from_file(tmp_path / "ck.nz")
    b.create_optimizer(1e-2)
    m.load_optimizer(b, tmp_path / "ck.nz.opt")
    for i in range(a.param_count()):
        assert np.array_equal(a.get_momentum(i), b.get_momentum(i))
    # resumed training continues identically (momentum restored, first-step flag cleared)
    la = a.train_microbatch(x, lab); a.step(1, 1e-2)
    lb = b.train_microbatch(x, lab); b.step(1, 1e-2)
    np.testing.assert_allclose(la, lb, rtol=0, atol=1e-6)
    for i in range(a.param_count()):
        np.testing.assert_allclose(a.get_param(i), b.get_param(i), rtol=0, atol=1e-6)
    out = tmp_path / "raw"
    os.makedirs(out)
    m.export_raw(a, out)
    meta = json.load(open(out / "model.json"))
    assert meta["channels"] == [1, 2] and meta["architecture"] == a.architecture and len(meta["tensors"]) == a.param_count()
    for i, t in enumerate(meta["tensors"]):
        p = a.get_param(i)
        assert t["shape"] == list(p.shape) and t["rows"] * t["cols"] == p.size and t["name"] == a.param_name(i)
        np.testing.assert_array_equal(np.fromfile(out / t["file"], np.float32).reshape(p.shape), p)
