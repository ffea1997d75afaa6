"""visual_perception_augmentation bindings (vpa_augment / unet3d_vpa_augment in include/unet3d_b200.h).
Same argument meaning as the reference call (train.hpp:43-48): options map, image with channels stacked along z,
label, is_label, shape, seed; the work happens in place."""
import ctypes

import numpy as np

# option ids and defaults of the reference's options.txt (lines 1-39)
OPTION_DEFAULTS = {
    "cropping": 0, "cropping_size_min": 0.1, "cropping_size_max": 0.2, "truncation_z": 1,
    "downsample_x": 2, "downsample_x_ratio": 0.5, "downsample_y": 2, "downsample_y_ratio": 0.5,
    "downsample_z": 2, "downsample_z_ratio": 0.5, "noise": 2, "noise_mag": 0.2,
    "ambient": 2, "ambient_mag": 2.0, "diffuse": 2, "diffuse_mag": 2.0,
    "specular": 2, "specular_freq": 2.0, "specular_mag": 0.5,
    "translocation_ratio": 0.2, "rotation_x": 0.2, "rotation_y": 0.2, "rotation_z": 0.2,
    "scaling_up": 1.25, "scaling_down": 0.8, "aspect_ratio": 1.25, "perspective": 0.1, "lens_distortion": 0.1,
    "distortion": 1, "distortion_count": 3, "distortion_radius_min": 0.1, "distortion_radius_max": 0.5,
    "distortion_mag_min": 0.05, "distortion_mag_max": 0.1,
    "zero_background": 1, "rubber_stamping": 2, "rubber_stamping_mag": 0.5, "perlin_texture": 2, "perlin_texture_mag": 0.5,
}

_F = ctypes.POINTER(ctypes.c_float)


def _opts(options):
    options = OPTION_DEFAULTS if options is None else options
    keys = [k.encode() for k in options]
    karr = (ctypes.c_char_p * len(keys))(*keys)
    varr = (ctypes.c_float * len(keys))(*[float(v) for v in options.values()])
    return karr, varr, len(keys)


def vpa_augment(image, label, options=None, is_label=True, seed=0, gpu=0):
    """image [C,D,H,W] fp32, label [D,H,W] fp32 (host); returns augmented copies."""
    from . import lib, check
    image = np.array(image, np.float32, copy=True, order="C")
    label = np.array(label, np.float32, copy=True, order="C")
    c, d, h, w = image.shape
    karr, varr, n = _opts(options)
    check(lib().vpa_augment(karr, varr, n, image.ctypes.data_as(_F), label.ctypes.data_as(_F), int(is_label), w, h, d, c,
                            ctypes.c_uint64(seed), 0, int(gpu)))
    return image, label


def train_microbatch_augmented(net, image, label, options=None, seed=0, collapse_before=0, use_ce=True, use_dice=True, use_mse=True):
    """Raw host sample -> one upload -> augmentation in HBM -> micro-batch (unet3d_train_microbatch_augmented).  Returns (ce, dice, mse)."""
    from . import check
    karr, varr, n = _opts(options)
    image = np.ascontiguousarray(image, np.float32)
    label = np.ascontiguousarray(label, np.float32)
    out = np.zeros(3, np.float32)
    check(net._lib.unet3d_train_microbatch_augmented(net._h, karr, varr, n, image.ctypes.data_as(_F), label.ctypes.data_as(_F),
                                                     ctypes.c_uint64(seed), int(collapse_before), int(use_ce), int(use_dice), int(use_mse),
                                                     out.ctypes.data_as(_F)))
    return out


def prefetch_augmented(net, image, label, options=None, seed=0, where=0):
    """Starts upload + augmentation of the next sample on the handle's side stream (unet3d_prefetch_augmented).  image / label:
    host numpy arrays (where=0; keep them alive and unchanged until train_microbatch_prefetched returns) or raw device pointers."""
    from . import check
    karr, varr, n = _opts(options)
    if where == 0:
        ip, lp = image.ctypes.data_as(_F), label.ctypes.data_as(_F)
    else:
        ip, lp = ctypes.cast(image, _F), ctypes.cast(label, _F)
    check(net._lib.unet3d_prefetch_augmented(net._h, karr, varr, n, ip, lp, ctypes.c_uint64(seed), int(where)))


def train_microbatch_prefetched(net, collapse_before=0, use_ce=True, use_dice=True, use_mse=True):
    from . import check
    out = np.zeros(3, np.float32)
    check(net._lib.unet3d_train_microbatch_prefetched(net._h, int(collapse_before), int(use_ce), int(use_dice), int(use_mse),
                                                      out.ctypes.data_as(_F)))
    return out


def vpa_augment_on(net, image_ptr, label_ptr, w, h, d, channels, options=None, is_label=True, seed=0, where=1):
    """In place on raw pointers (device when where=1), stream-ordered on `net`'s stream."""
    from . import check
    karr, varr, n = _opts(options)
    check(net._lib.unet3d_vpa_augment(net._h, karr, varr, n, ctypes.cast(image_ptr, _F), ctypes.cast(label_ptr, _F), int(is_label),
                                      int(w), int(h), int(d), int(channels), ctypes.c_uint64(seed), int(where)))


def simulate_modality(t1w, label=None, max_label=0, seed=0, gpu=0):
    """simulate_modality (train.cpp:43-180) on host arrays: t1w [D,H,W] fp32 in [0,1]; label [D,H,W] fp32 integers 0..max_label or
    None for the image-only overload.  Returns the simulated copy (the reference works in place)."""
    from . import lib, check
    t1w = np.array(t1w, np.float32, copy=True, order="C")
    d, h, w = t1w.shape
    lp = None
    if label is not None:
        label = np.ascontiguousarray(label, np.float32)
        lp = label.ctypes.data_as(_F)
    check(lib().simulate_modality(t1w.ctypes.data_as(_F), lp, ctypes.c_uint(max_label), ctypes.c_uint(seed & 0xFFFFFFFF), w, h, d, 0, int(gpu)))
    return t1w


def simulate_modality_on(net, t1w_ptr, label_ptr, w, h, d, max_label=0, seed=0, where=1):
    """In place on raw pointers (device when where=1), stream-ordered on `net`'s stream; label_ptr None / 0 = image-only overload."""
    from . import check
    lp = ctypes.cast(label_ptr, _F) if label_ptr else None
    check(net._lib.unet3d_simulate_modality(net._h, ctypes.cast(t1w_ptr, _F), lp, ctypes.c_uint(max_label), ctypes.c_uint(seed & 0xFFFFFFFF),
                                            int(w), int(h), int(d), int(where)))


def set_simulate_modality(net, mode):
    """0 off, 1 labelled template overload, 2 image-only overload: run by train_microbatch_augmented / prefetch_augmented on the
    uploaded sample before the augmentation (train.cpp:459-462)."""
    from . import check
    check(net._lib.unet3d_set_simulate_modality(net._h, int(mode)))
