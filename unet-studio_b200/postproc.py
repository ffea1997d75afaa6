"""Inference pre/post-processing bindings (unet3d_evaluate_volume, u3d_postproc, u3d_resample; include/unet3d_b200.h)."""
import ctypes

import numpy as np

_F = ctypes.POINTER(ctypes.c_float)
_B = ctypes.POINTER(ctypes.c_uint8)


def evaluate_volume(net, volume, stride=(0, 0, 0), mask_threshold=0.5, want_fg=True, want_prob=False):
    """volume [in, D, H, W] fp32 (any size) -> label map uint8 [D, H, W] (+ fg_prob, label_prob).  Windows of net.dim."""
    from . import check
    volume = np.ascontiguousarray(volume, np.float32)
    cin, d, h, w = volume.shape
    assert cin == net.in_count
    label = np.empty((d, h, w), np.uint8)
    fg = np.empty((d, h, w), np.float32) if want_fg else None
    prob = np.empty((net.out_count, d, h, w), np.float32) if want_prob else None
    n = ctypes.c_int(0)
    check(net._lib.unet3d_evaluate_volume(net._h, volume.ctypes.data_as(_F), w, h, d, int(stride[0]), int(stride[1]), int(stride[2]),
                                          ctypes.c_float(mask_threshold), label.ctypes.data_as(_B),
                                          fg.ctypes.data_as(_F) if want_fg else None, prob.ctypes.data_as(_F) if want_prob else None, 0,
                                          ctypes.byref(n)))
    return label, fg, prob, n.value


def window_origins(volume_dim, window_dim, stride):
    from . import lib
    buf = (ctypes.c_int * 4096)()
    n = lib().unet3d_window_origins(int(volume_dim), int(window_dim), int(stride), buf, 4096)
    return [buf[i] for i in range(n)]


def postproc(logits, mask_threshold=0.5, gpu=0):
    """logits [C, ...] fp32 -> (label uint8, fg_prob, label_prob) with the default "softmax+create_mask+argmax"."""
    from . import lib, check
    logits = np.ascontiguousarray(logits, np.float32)
    c = logits.shape[0]
    v = int(np.prod(logits.shape[1:]))
    label = np.empty(logits.shape[1:], np.uint8)
    fg = np.empty(logits.shape[1:], np.float32)
    prob = np.empty(logits.shape, np.float32)
    check(lib().u3d_postproc(logits.ctypes.data_as(_F), c, ctypes.c_longlong(v), ctypes.c_float(mask_threshold), label.ctypes.data_as(_B),
                             fg.ctypes.data_as(_F), prob.ctypes.data_as(_F), int(gpu)))
    return label, fg, prob


def resample(src, dst_dhw, nearest=False, gpu=0):
    from . import lib, check
    src = np.ascontiguousarray(src, np.float32)
    c, sd, sh, sw = src.shape
    dd, dh, dw = dst_dhw
    dst = np.empty((c, dd, dh, dw), np.float32)
    check(lib().u3d_resample(src.ctypes.data_as(_F), c, sw, sh, sd, dst.ctypes.data_as(_F), dw, dh, dd, int(bool(nearest)), int(gpu)))
    return dst
