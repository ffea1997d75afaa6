"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and its host-side feature_string parser agrees with the oracle (and with the reference's error messages)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from oracle import unet3d_oracle as O
from tests._pkg import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "unet3d_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:unet3d|u3d_op|vpa)_[a-z0-9_]+|simulate_modality[a-z_]*)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = load().lib()
    names = declared_functions()
    assert len(names) > 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/unet3d_b200.h but not exported"


def describe(in_c, out_c, feature):
    lib = load().lib()
    buf = ctypes.create_string_buffer(1 << 20)
    rc = lib.unet3d_describe(in_c, out_c, feature.encode(), buf, ctypes.c_size_t(len(buf)))
    if rc != 0:
        raise RuntimeError(lib.unet3d_last_error().decode())
    return json.loads(buf.value.decode())


FEATURES = [
    O.default_feature(1), O.default_feature(2), O.default_feature(6),
    "conv8,ks3,stride1+bnorm,relu\nmax_pool+conv16,ks3,stride1+bnorm,elu\nmax_pool+conv16,ks3,stride1+norm,relu+upsample\n"
    "conv16,ks3,stride1+bnorm,relu+conv2,ks1,stride1+upsample\nconv8,ks3,stride1+norm,elu+conv2,ks1,stride1",
    "conv4\nconv8,stride2+conv_trans4\nconv4+conv3,ks1",
]


@pytest.mark.parametrize("fi", range(len(FEATURES)))
def test_parser_matches_oracle(fi):
    f = FEATURES[fi]
    out_c = {0: 1, 1: 2, 2: 6, 3: 2, 4: 3}[fi]
    d = describe(1, out_c, f)
    net = O.parse_feature(1, out_c, f)
    assert [p["name"] for p in d["params"]] == net.param_names
    assert [tuple(p["shape"]) for p in d["params"]] == [tuple(s) for s in net.param_shapes]
    assert [bool(p["decay"]) for p in d["params"]] == [O.is_decay_param(n, s) for n, s in zip(net.param_names, net.param_shapes)]
    assert d["levels"] == len(net.output)


def test_default_net_counts():
    d = describe(1, 1, O.default_feature(1))
    assert len(d["params"]) == 108 and sum(int(np.prod(p["shape"])) for p in d["params"]) == 15023317
    d = describe(1, 2, O.default_feature(2))
    assert sum(int(np.prod(p["shape"])) for p in d["params"]) == 15023818


def test_default_feature_text():
    m = load()
    for oc in (1, 2, 6):
        assert m.default_feature(oc) == O.default_feature(oc)


@pytest.mark.parametrize("feature,msg", [
    ("conv8\nconv8", "invalid u-net structure"),
    ("conv8,ks5\nconv8\nconv8", "conv supports only ks1 stride1, ks3 stride1, and ks3 stride2"),
    ("conv8\nconv8+conv_trans8,ks3\nconv8", "conv_trans supports only ks2 stride2"),
    ("conv8\nfoo\nconv8", "unknown layer"),
])
def test_constructor_errors_carry_reference_messages(feature, msg):
    with pytest.raises(RuntimeError, match=msg):
        describe(1, 1, feature)


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = load()
    with pytest.raises(m.U3DError, match="no CPU fallback"):
        m.UNet3d(1, 2)


def test_window_origins_cover_the_volume_host_only():
    """unet3d_window_origins (host only): windows at 0, stride, ... plus the one ending at the border; every voxel covered."""
    from oracle import postproc_oracle as PO
    m = load()
    for vdim, wdim, stride in ((320, 160, 0), (320, 192, 128), (100, 160, 0), (161, 160, 80), (160, 160, 0), (500, 96, 64)):
        o = m.window_origins(vdim, wdim, stride)
        assert o == PO.window_origins(vdim, wdim, stride)
        covered = set()
        for s in o:
            covered.update(range(s, min(s + wdim, vdim)))
        assert covered == set(range(vdim)) and all(s + wdim <= max(vdim, wdim) for s in o)
