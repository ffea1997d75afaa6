"""Operator-level bindings (u3d_op_* in include/unet3d_b200.h): one reference layer, host numpy buffers."""
import ctypes

import numpy as np


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def conv_forward(x0, weight, bias=None, x1=None, transposed=False, ks=3, stride=1, want_stats=False, planar_fp32=False):
    """x0/x1: [C,D,H,W] fp32; weight in the reference layout.  Returns y [Cout,D',H',W'] (and stats [2,Cout])."""
    from . import lib, check
    x0 = _f32(x0); x1 = _f32(x1); weight = _f32(weight); bias = _f32(bias)
    cin0, d, h, w = x0.shape
    cin1 = 0 if x1 is None else x1.shape[0]
    cout = weight.shape[1] if transposed else weight.shape[0]
    if transposed:
        od, oh, ow = 2 * d, 2 * h, 2 * w
    else:
        p = (ks - 1) // 2
        od, oh, ow = [(n + 2 * p - ks) // stride + 1 for n in (d, h, w)]
    y = np.empty((cout, od, oh, ow), np.float32)
    stats = np.zeros((2, cout), np.float64) if want_stats else None
    check(lib().u3d_op_conv_forward(int(transposed), ks, stride, cin0, cin1, cout, w, h, d, _ptr(x0), _ptr(x1),
                                    _ptr(weight), _ptr(bias), _ptr(y), _ptr(stats, ctypes.c_double), int(planar_fp32)))
    return (y, stats) if want_stats else y


def conv_backward(x0, weight, dy, x1=None, transposed=False, ks=3, stride=1, gx0_init=None, want_gx=True, want_gw=True):
    """Returns (gx0, gx1, gw) of the layer given dy [Cout,D',H',W']."""
    from . import lib, check
    x0 = _f32(x0); x1 = _f32(x1); weight = _f32(weight); dy = _f32(dy)
    cin0, d, h, w = x0.shape
    cin1 = 0 if x1 is None else x1.shape[0]
    cout = weight.shape[1] if transposed else weight.shape[0]
    gx0 = np.array(gx0_init, np.float32, copy=True) if gx0_init is not None else np.empty_like(x0)
    gx1 = None if x1 is None else np.empty_like(x1)
    gw = np.empty_like(weight)
    flags = int(gx0_init is not None)
    check(lib().u3d_op_conv_backward(int(transposed), ks, stride, cin0, cin1, cout, w, h, d, _ptr(x0), _ptr(x1),
                                     _ptr(weight), _ptr(dy), _ptr(gx0) if want_gx else None,
                                     _ptr(gx1) if want_gx else None, _ptr(gw) if want_gw else None, flags))
    return gx0, gx1, gw


def maxpool_forward(x):
    """x [C,D,H,W] -> (y [C,D/2,H/2,W/2], indices int64 in torch's flat-offset convention)."""
    from . import lib, check
    x = _f32(x)
    c, d, h, w = x.shape
    y = np.empty((c, d // 2, h // 2, w // 2), np.float32)
    idx = np.empty(y.shape, np.int64)
    check(lib().u3d_op_maxpool_forward(c, w, h, d, _ptr(x), _ptr(y), _ptr(idx, ctypes.c_int64)))
    return y, idx
