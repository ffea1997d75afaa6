"""include/unet3d.hpp + the INTEGRATION.md binding stub compile with a plain host compiler (no nvcc, no torch) and link against
libunet3d_b200.so.  On a box without a GPU the program must report the loud constructor failure; with a GPU (-m gpu) it runs three
training steps and a forward through the C++ wrapper."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "integration_stub.cpp")
LIBDIR = os.path.join(ROOT, "unet-studio_b200")


def build(tmp_path):
    exe = os.path.join(str(tmp_path), "integration_stub")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe, "-L" + LIBDIR,
           "-l:libunet3d_b200.so", "-Wl,-rpath," + LIBDIR, "-Wl,--allow-shlib-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run(exe):
    env = dict(os.environ)
    import torch  # the CUDA runtime / NCCL the library links against live next to torch in this image
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    nv = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia")
    extra = [tl] + [os.path.join(nv, d, "lib") for d in ("cuda_runtime", "nccl")]
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    return subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)


def test_header_compiles_and_stub_fails_loudly_without_gpu(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libunet3d_b200.so")):
        pytest.skip("library not built")
    import torch
    r = run(build(tmp_path))
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "FEATURE_LINES 11" in r.stdout
    if torch.cuda.is_available():
        assert "CTOR_ERROR conv supports only ks1 stride1, ks3 stride1, and ks3 stride2" in r.stdout
        assert "LOSSES" in r.stdout and "LOGIT0" in r.stdout
    else:
        assert "NO_GPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_integration_stub_trains_through_the_cpp_wrapper(tmp_path):
    r = run(build(tmp_path))
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "CTOR_ERROR conv supports only ks1 stride1, ks3 stride1, and ks3 stride2" in r.stdout
    losses = [float(v) for v in r.stdout.split("LOSSES")[1].split()[:3]]
    assert all(0 < v < 2 for v in losses), losses
