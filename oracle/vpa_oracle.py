"""TEST INFRASTRUCTURE ONLY — CPU oracle for visual_perception_augmentation.  Not product code.

**parity unpinned**: the reference implementation (/root/reference/visual_perception_augmentation.cpp:163-438)
is written against the un-vendored, unpinned TIPL library (frankyeh/TIPL @ HEAD) and has no tests or golden
vectors, and it cannot be compiled here.  This file restates the .cpp literally (same stage order, same RNG draw
order) on top of EXPLICIT assumptions about the TIPL primitives it calls; each assumption is tagged [TIPL].
The CUDA path (unet-studio_b200/csrc/vpa.cu) is tested for exact agreement with this restatement.

[TIPL] uniform_dist<float>(a,b,seed)      = std::mt19937(uint32(seed)) + std::uniform_real_distribution<float>(a,b)
                                            (libstdc++: u = float(x)/2^32 in float32, clipped below 1; a + (b-a)*u)
[TIPL] scale(src,dst)                     = dst[p] = trilinear(src, p * src_dim/dst_dim), position clamped to the volume
[TIPL] interpolator::linear::get_location = valid iff 0 <= p <= dim-1 on every axis; upper neighbour clamped to dim-1
[TIPL] estimate<majority>                 = label with the largest summed trilinear weight (first in z,y,x order on ties)
[TIPL] transformation_matrix(arg,..)      = p -> Rz*Ry*Rx*diag(scale)*(p - dim/2) + dim/2 + translocation
[TIPL] normalize(I,upper=1)               = I *= upper/max(I) when max != 0;  lower_threshold(I,0) = max(I,0)
[TIPL] preserve(I,mask)                   = I = 0 where mask == 0;   masking(I,mask) = I = 0 where mask != 0
[TIPL] resample(src,dst,T)                = dst[p] = trilinear(src, T(p)) where get_location succeeds, else 0
[TIPL] for_each_neighbors(c,shape,r,f)    = every voxel of the cube [c-r, c+r]^3 clipped to the volume, r truncated to int
Unspecified in the reference itself (C++ argument evaluation order): the three draws of random_location and the
(location, radius, magnitude) draws of create_distortion_at are taken left to right.
Deliberate deviations, shared with the CUDA path and documented in DESIGN.md:
  * per-voxel noise: by default a counter-based hash of (seed, index); the library option noise_mt19937 = 1 selects the reference
    CPU path's sequential mt19937 stream bit for bit (noise_field_mt19937).  The default stays the hash (the
    reference's own CUDA path already differs from its CPU path there: curand_init(0,index,0), .cu:64-73);
  * std::shuffle of the Perlin permutation is implementation-defined: restated from libstdc++ 13 (std_shuffle below), which is what
    the library's host plan calls;
  * a distortion focus voxel itself (length 0, 0/0 in the reference) gets no displacement.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

OPTION_DEFAULTS = {  # /root/reference/options.txt:1-39 (id -> default)
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


class MT19937:
    """std::mt19937 (init_genrand seeding)."""

    def __init__(self, seed):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


class UniformDist:
    def __init__(self, lo, hi, seed):
        self.gen, self.lo, self.hi = MT19937(seed), F(lo), F(hi)

    def __call__(self):
        u = F(self.gen()) / F(4294967296.0)
        if u >= F(1.0):
            u = np.nextafter(F(1.0), F(0.0))
        return F((self.hi - self.lo) * u + self.lo)


def hash32(x):
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def noise_field_mt19937(seed, n, mag):
    """The reference CPU path's noise stream (visual_perception_augmentation.cpp:252-258): ONE tipl::uniform_dist<float>(0, mag, seed)
    [TIPL: std::mt19937(seed) + std::uniform_real_distribution<float>] drawn once per voxel, channel after channel.  libstdc++:
    u = float(word) / 2^32 (clipped below 1), value = (mag - 0) * u + 0, all in float32.  Raw words from numpy's MT19937 with the
    classic init_genrand seeding (checked against the MT19937 class above in tests/test_vpa_plan_cpu.py)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    words = bg.random_raw(n).astype(np.uint32)
    u = words.astype(F) / F(4294967296.0)
    u = np.where(u >= F(1.0), np.nextafter(F(1.0), F(0.0)), u).astype(F)
    return (F(mag) * u).astype(F)


def noise_field(seed, n):
    k = hash32(np.uint32(seed & 0xFFFFFFFF))
    h = hash32(np.arange(n, dtype=np.uint32) ^ k)
    return (h >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)


def trilinear(src, px, py, pz, clamp):
    """src [D,H,W]; positions float32 arrays.  Returns (values, valid)."""
    D, H, W = src.shape
    if clamp:
        px = np.clip(px, F(0), F(W - 1)); py = np.clip(py, F(0), F(H - 1)); pz = np.clip(pz, F(0), F(D - 1))
        valid = np.ones(px.shape, bool)
    else:
        valid = (px >= 0) & (px <= W - 1) & (py >= 0) & (py <= H - 1) & (pz >= 0) & (pz <= D - 1)
        px = np.where(valid, px, F(0)); py = np.where(valid, py, F(0)); pz = np.where(valid, pz, F(0))
    x0 = np.floor(px).astype(np.int64); y0 = np.floor(py).astype(np.int64); z0 = np.floor(pz).astype(np.int64)
    fx = (px - x0.astype(F)).astype(F); fy = (py - y0.astype(F)).astype(F); fz = (pz - z0.astype(F)).astype(F)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1); z1 = np.minimum(z0 + 1, D - 1)
    out = np.zeros(px.shape, F)
    for zz, wz in ((z0, F(1) - fz), (z1, fz)):
        for yy, wy in ((y0, F(1) - fy), (y1, fy)):
            for xx, wx in ((x0, F(1) - fx), (x1, fx)):
                out = out + src[zz, yy, xx] * ((wz * wy).astype(F) * wx).astype(F)
    return np.where(valid, out, F(0)).astype(F), valid


def majority(lab, px, py, pz):
    D, H, W = lab.shape
    valid = (px >= 0) & (px <= W - 1) & (py >= 0) & (py <= H - 1) & (pz >= 0) & (pz <= D - 1)
    px = np.where(valid, px, F(0)); py = np.where(valid, py, F(0)); pz = np.where(valid, pz, F(0))
    x0 = np.floor(px).astype(np.int64); y0 = np.floor(py).astype(np.int64); z0 = np.floor(pz).astype(np.int64)
    fx = (px - x0.astype(F)).astype(F); fy = (py - y0.astype(F)).astype(F); fz = (pz - z0.astype(F)).astype(F)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1); z1 = np.minimum(z0 + 1, D - 1)
    vals, wts = [], []
    for zz, wz in ((z0, F(1) - fz), (z1, fz)):
        for yy, wy in ((y0, F(1) - fy), (y1, fy)):
            for xx, wx in ((x0, F(1) - fx), (x1, fx)):
                vals.append(lab[zz, yy, xx]); wts.append(((wz * wy).astype(F) * wx).astype(F))
    vals = np.stack(vals); wts = np.stack(wts)
    best = vals[0].copy(); best_w = np.full(px.shape, F(-1))
    for i in range(8):
        tot = np.zeros(px.shape, F)
        for j in range(8):
            tot = tot + np.where(vals[j] == vals[i], wts[j], F(0))
        upd = tot > best_w
        best = np.where(upd, vals[i], best); best_w = np.where(upd, tot, best_w)
    return np.where(valid, best, F(0)).astype(F)


def scale(src, dst_shape):
    D, H, W = src.shape
    d, h, w = dst_shape
    z, y, x = np.meshgrid(np.arange(d, dtype=F), np.arange(h, dtype=F), np.arange(w, dtype=F), indexing="ij")
    v, _ = trilinear(src, (x * (F(W) / F(w))).astype(F), (y * (F(H) / F(h))).astype(F), (z * (F(D) / F(d))).astype(F), True)
    return v


def affine_matrix(t, r, s, dim):
    """[TIPL] p -> Rz*Ry*Rx*diag(s)*(p - c) + c + t, c = dim/2; returns 3x4 float32."""
    cx, sx = math.cos(float(r[0])), math.sin(float(r[0]))
    cy, sy = math.cos(float(r[1])), math.sin(float(r[1]))
    cz, sz = math.cos(float(r[2])), math.sin(float(r[2]))
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    A = (Rz @ Ry @ Rx) @ np.diag([float(s[0]), float(s[1]), float(s[2])])
    c = np.array([dim[0] * 0.5, dim[1] * 0.5, dim[2] * 0.5])
    b = c + np.array([float(t[0]), float(t[1]), float(t[2])]) - A @ c
    return np.concatenate([A, b[:, None]], 1).astype(F)


def apply_affine(M, px, py, pz):
    ox = (M[0, 0] * px + M[0, 1] * py + M[0, 2] * pz + M[0, 3]).astype(F)
    oy = (M[1, 0] * px + M[1, 1] * py + M[1, 2] * pz + M[1, 3]).astype(F)
    oz = (M[2, 0] * px + M[2, 1] * py + M[2, 2] * pz + M[2, 3]).astype(F)
    return ox, oy, oz


_fade = lambda t: t * t * t * (t * (t * F(6) - F(15)) + F(10))
_lerp = lambda t, a, b: a + t * (b - a)


def _grad(h, x, y, z):
    h = h & 15
    u = np.where(h < 8, x, y)
    v = np.where(h < 4, y, np.where((h == 12) | (h == 14), x, z))
    return np.where(h & 1, -u, u) + np.where(h & 2, -v, v)


def perlin(x, y, z, p):
    """visual_perception_augmentation.cpp:110-155 (classic Perlin improved noise)."""
    xf0, yf0, zf0 = np.floor(x), np.floor(y), np.floor(z)
    xi = xf0.astype(np.int64) & 255; yi = yf0.astype(np.int64) & 255; zi = zf0.astype(np.int64) & 255
    xf = (x - xf0).astype(F); yf = (y - yf0).astype(F); zf = (z - zf0).astype(F)
    u, v, w = _fade(xf), _fade(yf), _fade(zf)
    p = np.asarray(p, np.int64)
    A = p[xi] + yi; B = p[xi + 1] + yi
    aaa = p[p[A] + zi]; aba = p[p[A + 1] + zi]; aab = p[p[A] + zi + 1]; abb = p[p[A + 1] + zi + 1]
    baa = p[p[B] + zi]; bba = p[p[B + 1] + zi]; bab = p[p[B] + zi + 1]; bbb = p[p[B + 1] + zi + 1]
    x1 = _lerp(u, _grad(aaa, xf, yf, zf), _grad(baa, xf - 1, yf, zf))
    x2 = _lerp(u, _grad(aba, xf, yf - 1, zf), _grad(bba, xf - 1, yf - 1, zf))
    y1 = _lerp(v, x1, x2)
    x1 = _lerp(u, _grad(aab, xf, yf, zf - 1), _grad(bab, xf - 1, yf, zf - 1))
    x2 = _lerp(u, _grad(abb, xf, yf - 1, zf - 1), _grad(bbb, xf - 1, yf - 1, zf - 1))
    y2 = _lerp(v, x1, x2)
    return _lerp(w, y1, y2).astype(F)


def _lemire_u32(g, rng_range):
    """libstdc++ >= 11 uniform_int_distribution over a 32-bit engine: Lemire's nearly-divisionless method (bits/uniform_int_dist.h,
    _S_nd<uint64_t>): uniform integer in [0, rng_range)."""
    product = g() * rng_range
    low = product & 0xFFFFFFFF
    if low < rng_range:
        threshold = ((1 << 32) - rng_range) % rng_range
        while low < threshold:
            product = g() * rng_range
            low = product & 0xFFFFFFFF
    return product >> 32


def std_shuffle(p, g):
    """std::shuffle(first, last, std::mt19937) as implemented by libstdc++ 13 (bits/stl_algo.h): because the engine's range (2^32-1)
    divided by the element count is >= the element count, swap positions are produced TWO per distribution call
    (__gen_two_uniform_ints: x = uniform[0, b0*b1), (x / b1, x % b1)).  std::shuffle is implementation-defined; the library's host
    plan calls the real std::shuffle of the same libstdc++, this is its restatement."""
    n = len(p)
    if n == 0:
        return p
    assert 0xFFFFFFFF // n >= n
    i = 1
    if n % 2 == 0:
        j = _lemire_u32(g, 2)
        p[i], p[j] = p[j], p[i]
        i += 1
    while i != n:
        swap_range = i + 1
        x = _lemire_u32(g, swap_range * (swap_range + 1))
        a, b = x // (swap_range + 1), x % (swap_range + 1)
        p[i], p[a] = p[a], p[i]
        i += 1
        p[i], p[b] = p[b], p[i]
        i += 1
    return p


def perlin_table(seed):
    """visual_perception_augmentation.cpp:388-392: p[i] = i & 255, std::shuffle(p, std::mt19937(seed))."""
    return std_shuffle([i & 255 for i in range(512)], MT19937(seed))


def augment(options, image, label, is_label, shape, seed, trace=None):
    """visual_perception_augmentation.cpp:163-438.  image [C,D,H,W] fp32, label [D,H,W] fp32, shape = (W,H,D).
    Returns (image_out, label_out).  `trace` (dict) receives the drawn scalars for debugging."""
    opt = lambda k: F(options.get(k, 0.0))
    W, H, D = shape
    C = image.shape[0]
    img = image.astype(F).copy()
    lab = label.astype(F).copy()
    V = W * H * D
    one = UniformDist(-1.0, 1.0, seed)
    rng = lambda a, b: F(F(F(one() * F(F(b) - F(a))) * F(0.5)) + F(F(F(b) + F(a)) * F(0.5)))

    def apply(name):
        idx = int(opt(name))
        if idx == 0:
            return False
        if idx >= 4:
            return True
        return abs(one()) < F(idx) * F(0.25)

    def random_location(a, b):
        return (int(F(W - 1) * rng(a, b)), int(F(H - 1) * rng(a, b)), int(F(D - 1) * rng(a, b)))

    zz, yy, xx = np.meshgrid(np.arange(D, dtype=F), np.arange(H, dtype=F), np.arange(W, dtype=F), indexing="ij")
    maxdim = max(W, H, D)
    # ---- downsample (:205-220)
    dsx, dsy, dsz = apply("downsample_x"), apply("downsample_y"), apply("downsample_z")
    if dsx or dsy or dsz:
        lw = int(F(W) * (opt("downsample_x_ratio") if dsx else F(1)))
        lh = int(F(H) * (opt("downsample_y_ratio") if dsy else F(1)))
        ld = int(F(D) * (opt("downsample_z_ratio") if dsz else F(1)))
        for c in range(C):
            img[c] = scale(scale(img[c], (ld, lh, lw)), (D, H, W))
    # ---- cropping (:222-230)
    if apply("cropping"):
        size = rng(opt("cropping_size_min"), opt("cropping_size_max")) * F(W)
        value = rng(0.0, 2.0)
        loc = random_location(size, F(1.0) - size)
        r = int(size)
        x0, x1 = max(loc[0] - r, 0), min(loc[0] + r, W - 1)
        y0, y1 = max(loc[1] - r, 0), min(loc[1] + r, H - 1)
        z0, z1 = max(loc[2] - r, 0), min(loc[2] + r, D - 1)
        for c in range(C):
            if x0 <= x1 and y0 <= y1 and z0 <= z1:
                sub = lab[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
                m = sub != 0
                img[c, z0:z1 + 1, y0:y1 + 1, x0:x1 + 1][m] = value
                sub[m] = 0
    # ---- truncation (:231-250)
    if apply("truncation_z"):
        top = int(abs(F(one() * F(0.5)) * F(D)))
        bot = int(abs(F(one() * F(0.5)) * F(D)))
        if top:
            lab[D - top:] = 0; img[:, D - top:] = 0
        if bot:
            lab[:bot] = 0; img[:, :bot] = 0
    # ---- noise (:252-258)
    if apply("noise"):
        if opt("noise_mt19937") != 0:   # library option: the reference CPU path's sequential stream instead of the counter-based hash
            img += noise_field_mt19937(seed, C * V, opt("noise_mag")).reshape(C, D, H, W)
        else:
            img += (noise_field(seed, C * V) * opt("noise_mag")).reshape(C, D, H, W)
    # ---- lighting (:260-277)
    if apply("ambient"):
        img += rng(0.0, 1.0) * opt("ambient_mag")
    if apply("diffuse"):
        d = np.array([rng(-0.5, 0.5), rng(-0.5, 0.5), rng(-0.5, 0.5)], F)
        d = (d / F(math.sqrt(float(d[0]) ** 2 + float(d[1]) ** 2 + float(d[2]) ** 2))).astype(F)
        f = (d * (opt("diffuse_mag") / F(maxdim))).astype(F)
        g = np.maximum(F(0), F(1) + ((xx - F(W * 0.5)) * f[0] + (yy - F(H * 0.5)) * f[1] + (zz - F(D * 0.5)) * f[2])).astype(F)
        img *= g
    if apply("specular"):
        loc = random_location(0.4, 0.6)
        mag = opt("specular_mag")
        b = F(F(1.0) - mag - mag)
        freq = F(float(opt("specular_freq")) * (math.acos(-1.0) * 0.5 / maxdim))
        dist = np.sqrt((xx - F(loc[0])) ** 2 + (yy - F(loc[1])) ** 2 + (zz - F(loc[2])) ** 2).astype(F)
        img *= ((np.cos(dist * freq) + F(1)) * mag + b).astype(F)
    # ---- rigid motion + view port (:280-336)
    resolution = rng(F(1) / opt("scaling_up"), F(1) / opt("scaling_down"))
    tr = opt("translocation_ratio")
    t = (one() * tr * F(W), one() * tr * F(H), one() * tr * F(D))
    r = (one() * opt("rotation_x"), one() * opt("rotation_y"), one() * opt("rotation_z"))
    asp = opt("aspect_ratio")
    s = (resolution * rng(F(1) / asp, asp), resolution * rng(F(1) / asp, asp), resolution * rng(F(1) / asp, asp))
    M = affine_matrix(t, r, s, (W, H, D))
    persp = (rng(-0.5, 0.5) * opt("perspective") / F(W), rng(-0.5, 0.5) * opt("perspective") / F(H),
             rng(-0.5, 0.5) * opt("perspective") / F(D))
    dx = np.zeros((D, H, W), F); dy = np.zeros((D, H, W), F); dz = np.zeros((D, H, W), F)
    lens_mag = None
    if opt("lens_distortion") != 0:
        lens_mag = rng(0.0, 1.0) * opt("lens_distortion")
        radius = F(maxdim // 2)
        k = F(-(lens_mag / (radius * radius)))
        ex, ey, ez = xx - F(W // 2), yy - F(H // 2), zz - F(D // 2)
        l2 = (ex * ex + ey * ey + ez * ez).astype(F)
        dx, dy, dz = (ex * (k * l2)).astype(F), (ey * (k * l2)).astype(F), (ez * (k * l2)).astype(F)
    foci = []
    if apply("distortion"):
        num = int(rng(1.0, opt("distortion_count") + F(1.0)))
        for _ in range(num):
            loc = random_location(0.3, 0.7)
            radius = F(W) * rng(opt("distortion_radius_min"), opt("distortion_radius_max"))
            mag = rng(opt("distortion_mag_min"), opt("distortion_mag_max"))
            foci.append((loc, radius, mag))
            ri = int(radius)
            x0, x1 = max(loc[0] - ri, 0), min(loc[0] + ri, W - 1)
            y0, y1 = max(loc[1] - ri, 0), min(loc[1] + ri, H - 1)
            z0, z1 = max(loc[2] - ri, 0), min(loc[2] + ri, D - 1)
            if x0 > x1 or y0 > y1 or z0 > z1:
                continue
            sl = (slice(z0, z1 + 1), slice(y0, y1 + 1), slice(x0, x1 + 1))
            ex, ey, ez = xx[sl] - F(loc[0]), yy[sl] - F(loc[1]), zz[sl] - F(loc[2])
            ln = np.sqrt(ex * ex + ey * ey + ez * ez).astype(F)
            ok = (ln <= radius) & (ln > 0)
            coef = np.where(ok, F(-(radius * mag)) * np.sin(ln * F(math.acos(-1.0) / float(radius))) / np.where(ok, ln, F(1)), F(0)).astype(F)
            dx[sl] += ex * coef; dy[sl] += ey * coef; dz[sl] += ez * coef
    px, py, pz = xx.copy(), yy.copy(), zz.copy()
    if opt("lens_distortion") > 0:
        px, py, pz = px + dx, py + dy, pz + dz
    if opt("perspective") > 0:
        den = (persp[0] * (px - F(W / 2.0)) + persp[1] * (py - F(H / 2.0)) + persp[2] * (pz - F(D / 2.0)) + F(1)).astype(F)
        px, py, pz = (px / den).astype(F), (py / den).astype(F), (pz / den).astype(F)
    px, py, pz = apply_affine(M, px, py, pz)
    out = np.zeros_like(img)
    if is_label:
        out_lab = majority(lab, px, py, pz)
    else:
        out_lab, _ = trilinear(lab, px, py, pz, False)
    for c in range(C):
        out[c], _ = trilinear(img[c], px, py, pz, False)
    if trace is not None:
        trace.update(dict(M=M, persp=persp, lens_mag=lens_mag, foci=foci, resolution=resolution))

    def normalize(a, upper=F(1)):
        mx = a.max()
        if mx != 0:
            a *= F(upper) / mx

    for c in range(C):
        np.maximum(out[c], 0, out=out[c])
        normalize(out[c])
    # ---- background (:345-425)
    if is_label:
        if apply("zero_background"):
            out *= (out_lab != 0)
            return out, out_lab
        bg_mask = out_lab == 0

        def blend(dst, bg):
            dst[bg_mask] += (bg * np.maximum(F(0.1), F(1.0) - dst))[bg_mask]

        if apply("rubber_stamping"):
            pi2 = F(math.acos(-1.0) * 2.0)
            args = []
            for _ in range(5):
                tt = (one() * F(W) * F(0.5), one() * F(H) * F(0.5), one() * F(D) * F(0.5))
                rr = (one() * pi2, one() * pi2, one() * pi2)
                ss = (rng(0.8, 1.25), rng(0.8, 1.25), rng(0.8, 1.25))
                args.append(affine_matrix(tt, rr, ss, (W, H, D)))
            for c in range(C):
                img[c] = np.where(lab != 0, F(0), img[c])  # masking
                for it in range(5):
                    qx, qy, qz = apply_affine(args[it], xx, yy, zz)
                    bg, _ = trilinear(img[c], qx, qy, qz, False)
                    np.maximum(bg, 0, out=bg)
                    normalize(bg, rng(0.0, 1.0) * opt("rubber_stamping_mag"))
                    blend(out[c], bg)
        if apply("perlin_texture"):
            p = perlin_table(seed)
            zoom = rng(0.005, 0.05)
            if trace is not None:
                trace.update(dict(perlin_perm=list(p), perlin_zoom=zoom))
            bg = np.zeros((D, H, W), F)
            for octave in range(4):
                po = F(0.5 ** octave)
                sc = F(zoom * po)
                bg += perlin(xx * sc, yy * sc, zz * sc, p) * po
            bg = (bg * F(2)).astype(F)
            bg = (bg - np.floor(bg)).astype(F)
            normalize(bg, rng(0.0, 1.0) * opt("perlin_texture_mag"))
            for c in range(C):
                blend(out[c], bg)
        for c in range(C):
            np.maximum(out[c], 0, out=out[c])
            normalize(out[c])
    return out, out_lab
