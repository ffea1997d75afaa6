"""Model file (.nz) bindings: load_from_file / save_to_file of the reference (main.cpp:157-233) and the Level-4 MAT-in-gzip
container itself (host only)."""
import ctypes

import numpy as np

_TYPES = {0: np.float64, 10: np.float32, 20: np.int32, 30: np.int16, 40: np.uint16, 50: np.uint8,
          1: np.float64, 11: np.float32, 21: np.int32, 31: np.int16, 41: np.uint16, 51: np.uint8}


class NzFile:
    """The container: an ordered list of named matrices (column-major)."""

    def __init__(self, path=None):
        from . import lib, check
        self._lib = lib()
        self._lib.u3d_nz_free.restype = None
        self._z = ctypes.c_void_p()
        if path is None:
            check(self._lib.u3d_nz_create(ctypes.byref(self._z)))
        else:
            check(self._lib.u3d_nz_load(str(path).encode(), ctypes.byref(self._z)))

    def __del__(self):
        if getattr(self, "_z", None):
            self._lib.u3d_nz_free(self._z)
            self._z = None

    def names(self):
        out = []
        for i in range(self._lib.u3d_nz_count(self._z)):
            out.append(self.info(i))
        return out

    def info(self, i):
        from . import check
        name = ctypes.create_string_buffer(256)
        t, r, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self._lib.u3d_nz_info(self._z, i, name, ctypes.c_size_t(256), ctypes.byref(t), ctypes.byref(r), ctypes.byref(c)))
        return name.value.decode(), t.value, r.value, c.value

    def add(self, name, array, type_code=None):
        """array: 2-D (rows, cols) or 1-D (stored as 1 x n); str -> text matrix."""
        from . import check
        if isinstance(array, str):
            a = np.frombuffer(array.encode(), np.uint8)
            type_code, rows, cols = 51, 1, a.size
        else:
            a = np.asarray(array)
            if type_code is None:
                type_code = {np.dtype(np.float64): 0, np.dtype(np.float32): 10, np.dtype(np.int32): 20, np.dtype(np.int16): 30,
                             np.dtype(np.uint16): 40, np.dtype(np.uint8): 50}[a.dtype]
            a = a.astype(_TYPES[type_code], copy=False)
            if a.ndim == 1:
                a = a[None]
            rows, cols = a.shape
            a = np.asfortranarray(a).ravel(order="F")
        a = np.ascontiguousarray(a)
        check(self._lib.u3d_nz_add(self._z, name.encode(), int(type_code), int(rows), int(cols), a.ctypes.data_as(ctypes.c_void_p)))

    def read_f32(self, name):
        from . import check
        info = {n: (t, r, c) for n, t, r, c in self.names()}
        t, r, c = info[name]
        out = np.empty(r * c, np.float32)
        check(self._lib.u3d_nz_read_f32(self._z, name.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_size_t(out.size)))
        return out.reshape(c, r).T   # column-major -> (rows, cols)

    def save(self, path):
        from . import check
        check(self._lib.u3d_nz_save(self._z, str(path).encode()))


def load_from_file(path, gpu=0):
    """main.cpp:157-206 -> UNet3d (training mode, parameters / dim / voxel_size / metadata from the file)."""
    from . import lib, check
    from .model import UNet3d
    h = ctypes.c_void_p()
    check(lib().unet3d_load_from_file(str(path).encode(), int(gpu), ctypes.byref(h)))
    return UNet3d._from_handle(h, gpu)


def save_to_file(net, path):
    """main.cpp:207-233."""
    from . import check
    check(net._lib.unet3d_save_to_file(net._h, str(path).encode()))


def save_optimizer(net, path):
    from . import check
    check(net._lib.unet3d_save_optimizer(net._h, str(path).encode()))


def load_optimizer(net, path):
    from . import check
    check(net._lib.unet3d_load_optimizer(net._h, str(path).encode()))


def export_raw(net, directory):
    from . import check
    check(net._lib.unet3d_export_raw(net._h, str(directory).encode()))
