"""unet-studio_b200 — Python host-side mirror of the reference's UNet3d interface over the C-ABI of
libunet3d_b200.so (include/unet3d_b200.h).  The compute path is the CUDA library; there is no CPU
fallback: loading fails loudly if the shared library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunet3d_b200.so")


class U3DError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libunet3d_b200.so (built by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise U3DError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (no CPU fallback exists)")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.unet3d_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise U3DError(lib().unet3d_last_error().decode(errors="replace"))


from .ops import conv_forward, conv_backward, maxpool_forward  # noqa: E402,F401
from .model import UNet3d, default_feature, poly_lr  # noqa: E402,F401
from .vpa import (vpa_augment, vpa_augment_on, train_microbatch_augmented, prefetch_augmented, train_microbatch_prefetched,  # noqa: E402,F401
                  simulate_modality, simulate_modality_on, set_simulate_modality, OPTION_DEFAULTS)  # noqa: E402,F401
from .modelfile import NzFile, load_from_file, save_to_file, save_optimizer, load_optimizer, export_raw  # noqa: E402,F401
from .postproc import evaluate_volume, window_origins, postproc, resample  # noqa: E402,F401
from . import dist  # noqa: E402,F401
