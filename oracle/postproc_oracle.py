"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the inference post-processing "softmax+create_mask+argmax" and the window
cutting / re-assembly around the forward.  Not product code.

**parity unpinned**: the reference runs these steps inside tipl::ml3d::evalution_set (handle_fov_pre, handle_fov_post,
run_postproc; /root/reference/evaluate.cpp:201-204,274), which is part of the un-vendored TIPL library.  What the reference itself
shows: the default postproc string (unet.cpp:112), the argmax call `tipl::argmax(prob4d, shape, mask > threshold)`
(evaluate.cpp:315-319), the window loop (evaluate.cpp:223-230) and that label 0 is background with out_count = max(label)+1
(train.cpp:1125).  Assumed semantics [TIPL]:
  softmax      over the out_count channels of each voxel
  create_mask  fg_prob = 1 - p[0]
  argmax       label = first arg-max channel where fg_prob > threshold, else 0
  windows      origins 0, stride, 2*stride, ... while the window ends inside the volume, then the window ending at the border;
               a volume smaller than the window is zero-padded at the far end; probabilities of overlapping windows are averaged
  resample     tipl::scale-style: source position = destination index * src_dim / dst_dim (clamped), trilinear or nearest
"""
import numpy as np

F = np.float32


def softmax(logits):
    l = logits.astype(F)
    e = np.exp(l - l.max(0, keepdims=True), dtype=F)
    return (e / e.sum(0, keepdims=True, dtype=F)).astype(F)


def mask_argmax(prob, threshold):
    fg = (F(1) - prob[0]).astype(F)
    label = np.where(fg > F(threshold), prob.argmax(0), 0).astype(np.uint8)
    return label, fg


def window_origins(vdim, wdim, stride):
    if vdim <= wdim:
        return [0]
    if stride < 1:
        stride = wdim
    o = list(range(0, vdim - wdim, stride))
    o.append(vdim - wdim)
    return o


def evaluate_volume(forward0, volume, win_dhw, stride_xyz, threshold):
    """forward0(window [C_in, wd, wh, ww]) -> logits [C, wd, wh, ww].  volume [C_in, D, H, W].  Returns (label, fg, prob, n_windows)."""
    cin, D, H, W = volume.shape
    wd, wh, ww = win_dhw
    acc = None
    cnt = np.zeros((D, H, W), F)
    n = 0
    for z in window_origins(D, wd, stride_xyz[2]):
        for y in window_origins(H, wh, stride_xyz[1]):
            for x in window_origins(W, ww, stride_xyz[0]):
                win = np.zeros((cin, wd, wh, ww), F)
                dz, dy, dx = min(wd, D - z), min(wh, H - y), min(ww, W - x)
                win[:, :dz, :dy, :dx] = volume[:, z:z + dz, y:y + dy, x:x + dx]
                p = softmax(forward0(win))
                if acc is None:
                    acc = np.zeros((p.shape[0], D, H, W), F)
                acc[:, z:z + dz, y:y + dy, x:x + dx] += p[:, :dz, :dy, :dx]
                cnt[z:z + dz, y:y + dy, x:x + dx] += 1
                n += 1
    prob = (acc / cnt).astype(F)
    label, fg = mask_argmax(prob, threshold)
    return label, fg, prob, n


def resample(src, dst_dhw, nearest=False):
    """src [C, sd, sh, sw] -> [C, dd, dh, dw]."""
    C, sd, sh, sw = src.shape
    dd, dh, dw = dst_dhw
    fz = np.minimum(np.arange(dd, dtype=F) * F(F(sd) / F(dd)), F(sd - 1))
    fy = np.minimum(np.arange(dh, dtype=F) * F(F(sh) / F(dh)), F(sh - 1))
    fx = np.minimum(np.arange(dw, dtype=F) * F(F(sw) / F(dw)), F(sw - 1))
    if nearest:
        iz = np.minimum((fz + F(0.5)).astype(np.int64), sd - 1)
        iy = np.minimum((fy + F(0.5)).astype(np.int64), sh - 1)
        ix = np.minimum((fx + F(0.5)).astype(np.int64), sw - 1)
        return src[:, iz][:, :, iy][:, :, :, ix]
    z0, y0, x0 = fz.astype(np.int64), fy.astype(np.int64), fx.astype(np.int64)
    z1, y1, x1 = np.minimum(z0 + 1, sd - 1), np.minimum(y0 + 1, sh - 1), np.minimum(x0 + 1, sw - 1)
    az, ay, ax = (fz - z0).astype(F)[None, :, None, None], (fy - y0).astype(F)[None, None, :, None], (fx - x0).astype(F)[None, None, None, :]

    def g(zi, yi, xi):
        return src[:, zi][:, :, yi][:, :, :, xi]
    c00 = g(z0, y0, x0) * (1 - ax) + g(z0, y0, x1) * ax
    c01 = g(z0, y1, x0) * (1 - ax) + g(z0, y1, x1) * ax
    c10 = g(z1, y0, x0) * (1 - ax) + g(z1, y0, x1) * ax
    c11 = g(z1, y1, x0) * (1 - ax) + g(z1, y1, x1) * ax
    return ((c00 * (1 - ay) + c01 * ay) * (1 - az) + (c10 * (1 - ay) + c11 * ay) * az).astype(F)
