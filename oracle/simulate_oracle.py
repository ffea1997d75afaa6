"""TEST INFRASTRUCTURE ONLY — CPU oracle for simulate_modality (both overloads).  Not product code.

**parity unpinned**: the reference (/root/reference/train.cpp:43-117 labelled-template overload, :119-180 image-only
overload; called once per sample right before visual_perception_augmentation, train.cpp:459-462) is written against
the un-vendored, unpinned TIPL library and has no tests or golden vectors, and train.cpp cannot be compiled here
(Qt + TIPL).  This file restates the two functions literally (same draw order, same float32 evaluation order) on
top of EXPLICIT assumptions about the TIPL primitives; each is tagged [TIPL].  The CUDA path
(unet-studio_b200/csrc/simulate.cu) is tested against this restatement: bit-exact up to the pow() call, 2e-6 after it.

[TIPL] uniform_dist<float>(0,1,seed)  = std::mt19937(seed) + std::uniform_real_distribution<float>(0,1)
                                        (libstdc++: u = float(x)/2^32 in float32, clipped below 1)
[TIPL] uniform_dist<int>(seed)(n)     = std::uniform_int_distribution<int>(0,n-1) on std::mt19937(seed); libstdc++ (gcc >= 11)
                                        maps one 32-bit draw x to (x*n) >> 32 and redraws while uint32(x*n) < (2^32-n) % n
                                        (never for n = 4)
[TIPL] filter::gaussian(I) (3-D)      = one pass of the 7-point star on LINEAR offsets: dest = 2*I, then += the neighbour at
                                        +1, -1, +W, -W, +W*H, -W*H in that order (terms whose index leaves [0,size) are
                                        dropped, x/y neighbours wrap across row / plane ends), then /8 whatever was dropped
[TIPL] upper_lower_threshold(I,0,1)   = clamp to [0,1]
"""
from __future__ import annotations

import numpy as np

from .vpa_oracle import MT19937, UniformDist

F = np.float32
TERM_COUNT = 20   # train.cpp:50


class UniformInt:
    def __init__(self, seed):
        self.gen = MT19937(seed)

    def __call__(self, n):
        thr = ((1 << 32) - n) % n
        while True:
            prod = self.gen() * n
            if (prod & 0xFFFFFFFF) >= thr:
                return prod >> 32


def gaussian(img):
    """[TIPL] filter::gaussian on a (D,H,W) float32 volume (x fastest)."""
    d, h, w = img.shape
    src = img.reshape(-1)
    dest = src * F(2.0)
    for shift in (1, -1, w, -w, w * h, -w * h):
        if shift > 0:      # dest[i] += src[i - shift]
            dest[shift:] += src[:-shift]
        else:              # dest[i] += src[i + |shift|]
            dest[:shift] += src[-shift:]
    return (dest / F(8.0)).reshape(img.shape)


def draw_terms(rand_int, rand_float):
    """train.cpp:65-78 / :131-144 — (a,b) redrawn until a+b != 0, then c, d, w."""
    terms = []
    for _ in range(TERM_COUNT):
        while True:
            a = rand_int(4)
            b = rand_int(4)
            if a + b != 0:
                break
        c = rand_int(4)
        d = rand_int(4)
        terms.append((a, b, c, d, rand_float()))
    return terms


def simulate_modality(t1w, label=None, max_label=0, seed=0, trace=None):
    """t1w: (D,H,W) float32 in [0,1]; label: (D,H,W) float32 integers 0..max_label, or None for the image-only overload.
    Returns the new t1w (the reference works in place)."""
    t1w = np.ascontiguousarray(t1w, F)
    rand_int = UniformInt(seed & 0xFFFFFFFF)
    rand_float = UniformDist(0.0, 1.0, (seed + 1) & 0xFFFFFFFF)
    if label is not None:
        lut = np.array([F(F(0.4) + F(rand_float() * F(0.2))) for _ in range(max_label + 1)], F)   # train.cpp:56-58
        tissue = lut[label.astype(np.int64)]                                                        # :59-60
    else:
        lut = np.zeros(0, F)
        tissue = t1w.copy()                                                                         # :127
    tissue = gaussian(gaussian(tissue))                                                             # :62-63
    terms = draw_terms(rand_int, rand_float)
    gamma = F(F(0.6) + F(F(1.2) * rand_float()))                                                    # :80
    x = t1w
    z = tissue
    rx = F(1.0) - x
    rz = F(1.0) - z
    one = np.ones_like(x)
    px = [one, x, x * x, x * x * x]
    pz = [one, z, z * z, z * z * z]
    qx = [one, rx, rx * rx, rx * rx * rx]
    qz = [one, rz, rz * rz, rz * rz * rz]
    s = np.zeros_like(x)
    for a, b, c, d, w in terms:                                                                      # :99-101
        s = s + ((F(w) * px[a]) * pz[b]) * qx[c] * qz[d]
    with np.errstate(invalid="ignore"):
        v = np.power(s, gamma).astype(F)                                                             # :103
    keep = ~(x <= F(0.02))                                                                              # :87-92
    out = np.where(keep, v, F(0.0)).astype(F)
    sel = keep & (label != 0) if label is not None else keep                                         # :104-108 / :169-170
    sel = sel & ~np.isnan(out)            # std::min/max keep the running value when the new one is NaN
    if trace is not None:
        trace.update(tissue=tissue, s=np.where(keep, s, F(0.0)), gamma=gamma, terms=terms, lut=lut, pre=out.copy())
    if sel.any():
        mn, mx = F(out[sel].min()), F(out[sel].max())
        if mx > mn:                                                                                  # :111-116
            out = out - mn
            out = out * F(F(1.0) / F(mx - mn))
            out = np.minimum(np.maximum(out, F(0.0)), F(1.0))
    return out
