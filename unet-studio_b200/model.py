"""Python mirror of the reference's UNet3d surface (unet.hpp:13-70) over the C-ABI (include/unet3d_b200.h).

Same names and argument meaning as the reference: UNet3d(in_count, out_count, feature_string), forward(),
train(), prepare_for_inference(), create_optimizer(), copy_from(); plus the training-step body
(train.cpp:628-706) as train_microbatch() and the update (train.cpp:755-766) as step()."""
import ctypes

import numpy as np

_F = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return a.ctypes.data_as(_F)


def default_feature(out_count):
    """train.cpp:1054-1069 (text produced by the library)."""
    from . import lib, U3DError
    L = lib()
    buf = ctypes.create_string_buffer(8192)
    if L.unet3d_default_feature(int(out_count), buf, ctypes.c_size_t(len(buf))) != 0:
        raise U3DError("default_feature buffer too small")
    return buf.value.decode()


def poly_lr(lr0, step, total_steps):
    """train.cpp:566."""
    return lr0 * (1.0 - step / total_steps) ** 0.9


class UNet3d:
    def __init__(self, in_count, out_count, feature_string=None, gpu=0):
        from . import lib, check
        self._lib = lib()
        L = self._lib
        L.unet3d_param_name.restype = ctypes.c_char_p
        L.unet3d_architecture.restype = ctypes.c_char_p
        L.unet3d_last_grad_norm.restype = ctypes.c_double
        L.unet3d_loss_scale.restype = ctypes.c_float
        L.unet3d_param_total.restype = ctypes.c_longlong
        L.unet3d_launch_count.restype = ctypes.c_longlong
        L.unet3d_destroy.restype = None
        if feature_string is None:
            feature_string = default_feature(out_count)
        self._h = ctypes.c_void_p()
        check(L.unet3d_create(int(in_count), int(out_count), feature_string.encode(), int(gpu), ctypes.byref(self._h)))
        self.in_count, self.out_count, self.architecture, self.gpu = int(in_count), int(out_count), feature_string, gpu
        self.levels = L.unet3d_levels(self._h)
        self.dim = (192, 224, 192)  # unet.hpp:38

    @classmethod
    def _from_handle(cls, h, gpu=0):
        """Wraps a handle made by the library (unet3d_load_from_file)."""
        from . import lib
        self = cls.__new__(cls)
        self._lib = L = lib()
        L.unet3d_param_name.restype = ctypes.c_char_p
        L.unet3d_architecture.restype = ctypes.c_char_p
        L.unet3d_last_grad_norm.restype = ctypes.c_double
        L.unet3d_loss_scale.restype = ctypes.c_float
        L.unet3d_param_total.restype = ctypes.c_longlong
        L.unet3d_launch_count.restype = ctypes.c_longlong
        L.unet3d_destroy.restype = None
        self._h = h
        self.in_count, self.out_count = L.unet3d_in_count(h), L.unet3d_out_count(h)
        self.architecture, self.gpu = L.unet3d_architecture(h).decode(), gpu
        self.levels = L.unet3d_levels(h)
        d = (ctypes.c_int * 3)()
        L.unet3d_get_dim(h, d)
        self.dim = (d[0], d[1], d[2])
        return self

    def get_info(self, key):
        from . import check
        buf = ctypes.create_string_buffer(65536)
        check(self._lib.unet3d_get_info(self._h, key.encode(), buf, ctypes.c_size_t(len(buf))))
        return buf.value.decode()

    def set_info(self, key, value):
        from . import check
        check(self._lib.unet3d_set_info(self._h, key.encode(), value.encode()))

    def set_errors(self, testing, errors):
        from . import check
        e = np.ascontiguousarray(errors, np.float32).reshape(-1, 3)
        check(self._lib.unet3d_set_errors(self._h, int(bool(testing)), _fp(e), int(e.shape[0])))

    def get_errors(self, testing):
        n = self._lib.unet3d_get_errors(self._h, int(bool(testing)), None, 0)
        e = np.zeros((max(n, 0), 3), np.float32)
        if n > 0:
            self._lib.unet3d_get_errors(self._h, int(bool(testing)), _fp(e), n)
        return e

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.unet3d_destroy(h)
            self._h = None

    # ---- parameters (tensorN order) ----
    def param_count(self):
        return self._lib.unet3d_param_count(self._h)

    def param_total(self):
        return self._lib.unet3d_param_total(self._h)

    def param_shape(self, i):
        dims = (ctypes.c_int64 * 5)()
        nd = ctypes.c_int()
        from . import check
        check(self._lib.unet3d_param_shape(self._h, i, dims, ctypes.byref(nd)))
        return tuple(int(dims[k]) for k in range(nd.value))

    def param_name(self, i):
        return self._lib.unet3d_param_name(self._h, i).decode()

    def _get(self, fn, i):
        from . import check
        a = np.empty(self.param_shape(i), np.float32)
        check(fn(self._h, i, _fp(a)))
        return a

    def get_param(self, i):
        return self._get(self._lib.unet3d_get_param, i)

    def get_grad(self, i):
        return self._get(self._lib.unet3d_get_grad, i)

    def get_momentum(self, i):
        return self._get(self._lib.unet3d_get_momentum, i)

    def set_param(self, i, a):
        from . import check
        a = np.ascontiguousarray(a, np.float32)
        assert a.shape == self.param_shape(i), (a.shape, self.param_shape(i))
        check(self._lib.unet3d_set_param(self._h, i, _fp(a)))

    def set_momentum(self, i, a):
        from . import check
        a = np.ascontiguousarray(a, np.float32)
        check(self._lib.unet3d_set_momentum(self._h, i, _fp(a)))

    def parameters(self):
        return [self.get_param(i) for i in range(self.param_count())]

    def load_parameters(self, arrays):
        for i, a in enumerate(arrays):
            self.set_param(i, a)

    def init_params(self, seed=0):
        from . import check
        check(self._lib.unet3d_init_params(self._h, ctypes.c_uint64(seed)))

    # ---- mode / geometry ----
    def set_dim(self, w, h, d):
        from . import check
        check(self._lib.unet3d_set_dim(self._h, int(w), int(h), int(d)))
        self.dim = (int(w), int(h), int(d))

    def train(self, on=True):
        """unet.hpp:58-62: train(False) is eval() (BatchNorm3d reads its running statistics)."""
        from . import check
        check(self._lib.unet3d_set_mode(self._h, 1 if on else 2))

    def prepare_for_inference(self):
        """unet.cpp:7-22 (eval + every BatchNorm3d reset to running_mean 0 / running_var 1)."""
        from . import check
        check(self._lib.unet3d_set_mode(self._h, 0))

    def eval(self):
        """torch Module::eval(): BatchNorm3d normalises with its tracked running statistics (the validation replica, train.cpp:836)."""
        from . import check
        check(self._lib.unet3d_set_mode(self._h, 2))

    # ---- forward / training ----
    def _level_shape(self, k):
        w, h, d = self.dim
        return (1, self.out_count, d >> k, h >> k, w >> k)

    def forward(self, x, n_levels=None, out=None):
        """x: [1,in,D,H,W] fp32 -> list of logits (results[k] of unet.cpp:168-193)."""
        from . import check
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 5 and x.shape[0] == 1 and x.shape[1] == self.in_count
        d, h, w = x.shape[2:]
        if (w, h, d) != self.dim:
            self.set_dim(w, h, d)
        n = self.levels if n_levels is None else n_levels
        outs = out if out is not None else [np.empty(self._level_shape(k), np.float32) for k in range(n)]
        ptrs = (_F * n)(*[_fp(o) for o in outs])
        check(self._lib.unet3d_forward(self._h, _fp(x), ptrs, n, 0))
        return outs

    def evaluate_windows(self, windows, outs=None):
        """forward()[0] of a list of windows [1,in,D,H,W] (evaluate.cpp:223-230); uploads / downloads overlap the compute."""
        from . import check
        windows = [np.ascontiguousarray(w, np.float32) for w in windows]
        d, h, w = windows[0].shape[2:]
        if (w, h, d) != self.dim:
            self.set_dim(w, h, d)
        if outs is None:
            outs = [np.empty(self._level_shape(0), np.float32) for _ in windows]
        n = len(windows)
        inp = (_F * n)(*[_fp(a) for a in windows])
        outp = (_F * n)(*[_fp(o) for o in outs])
        check(self._lib.unet3d_evaluate_windows(self._h, inp, outp, n, 0))
        return outs

    def create_optimizer(self, lr):
        from . import check
        check(self._lib.unet3d_create_optimizer(self._h, ctypes.c_float(lr)))

    def train_microbatch(self, x, label, collapse_before=0, use_ce=True, use_dice=True, use_mse=True, all_levels=False):
        from . import check
        x = np.ascontiguousarray(x, np.float32)
        label = np.ascontiguousarray(label, np.float32)
        d, h, w = x.shape[2:]
        if (w, h, d) != self.dim:
            self.set_dim(w, h, d)
        out = np.zeros(3, np.float32)
        lv = np.zeros((self.levels, 3), np.float32)
        check(self._lib.unet3d_train_microbatch(self._h, _fp(x), _fp(label), int(collapse_before), int(use_ce), int(use_dice),
                                                int(use_mse), _fp(out), _fp(lv), 0))
        return (out, lv) if all_levels else out

    def validate(self, x, label, collapse_before=0):
        from . import check
        x = np.ascontiguousarray(x, np.float32)
        label = np.ascontiguousarray(label, np.float32)
        out = np.zeros(3, np.float32)
        check(self._lib.unet3d_validate(self._h, _fp(x), _fp(label), int(collapse_before), _fp(out), 0))
        return out

    def validate_async(self, x, label, collapse_before=0):
        """Enqueues the validation on this handle's own stream and returns; keep x / label alive until validate_result()."""
        from . import check
        self._val_keep = (np.ascontiguousarray(x, np.float32), np.ascontiguousarray(label, np.float32))
        check(self._lib.unet3d_validate_async(self._h, _fp(self._val_keep[0]), _fp(self._val_keep[1]), int(collapse_before), 0))

    def validate_result(self):
        from . import check
        out = np.zeros(3, np.float32)
        check(self._lib.unet3d_validate_result(self._h, _fp(out)))
        self._val_keep = None
        return out

    def attach_comm(self, nccl_comm, microbatches_per_step=1):
        from . import check
        check(self._lib.unet3d_attach_comm(self._h, nccl_comm, int(microbatches_per_step)))

    def step(self, batch_size, lr, nccl_comm=None):
        from . import check
        check(self._lib.unet3d_step(self._h, int(batch_size), ctypes.c_double(lr), nccl_comm))
        return self._lib.unet3d_last_grad_norm(self._h)

    def last_step_skipped(self):
        return bool(self._lib.unet3d_last_step_skipped(self._h))

    def loss_scale(self):
        return float(self._lib.unet3d_loss_scale(self._h))

    def launch_count(self):
        return int(self._lib.unet3d_launch_count(self._h))

    def copy_from(self, other):
        from . import check
        check(self._lib.unet3d_copy_from(self._h, other._h))
        self.dim = other.dim

    def timer_start(self):
        from . import check
        check(self._lib.unet3d_timer_start(self._h))

    def timer_stop(self):
        from . import check
        ms = ctypes.c_float()
        check(self._lib.unet3d_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def device_forward(self, x_dev_ptr, out_dev_ptrs):
        """forward with device-resident fp32 NCDHW buffers (raw pointers); stream-ordered, no sync."""
        from . import check
        n = len(out_dev_ptrs)
        ptrs = (_F * n)(*[ctypes.cast(p, _F) for p in out_dev_ptrs])
        check(self._lib.unet3d_forward(self._h, ctypes.cast(x_dev_ptr, _F), ptrs, n, 1))

    def device_train_microbatch(self, x_dev_ptr, label_dev_ptr, collapse_before=0, use_ce=True, use_dice=True, use_mse=True):
        from . import check
        out = np.zeros(3, np.float32)
        check(self._lib.unet3d_train_microbatch(self._h, ctypes.cast(x_dev_ptr, _F), ctypes.cast(label_dev_ptr, _F),
                                                int(collapse_before), int(use_ce), int(use_dice), int(use_mse), _fp(out), None, 1))
        return out

    def profile(self, enable=True):
        from . import check
        check(self._lib.unet3d_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        from . import check
        out = (ctypes.c_double * 24)()
        check(self._lib.unet3d_profile_read(self._h, out, int(reset)))
        return [float(v) for v in out]

    def sync(self):
        from . import check
        check(self._lib.unet3d_sync(self._h))
