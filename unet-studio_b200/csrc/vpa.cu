// visual_perception_augmentation on the GPU (north-star kernel 5): random affine / perspective / lens / local
// distortion warp with trilinear (image) and majority (label) gathers, intensity stages (truncation, noise,
// ambient / diffuse / specular light), background synthesis (zero, rubber-stamping, Perlin) and renormalisation.
// Semantics = /root/reference/visual_perception_augmentation.cpp:163-438 (the CPU path; the reference's own .cu
// differs from it in places) with the TIPL assumptions listed in oracle/vpa_oracle.py.
// All random scalars are drawn on the host in the reference's draw order (std::mt19937 +
// uniform_real_distribution<float>(-1,1)); the kernels are pure functions of the resulting plan, so the
// displacement field is evaluated analytically per voxel instead of being materialised (3 x fp32 x V saved).
// Bandwidth-bound: every pass is a coalesced fp32 planar sweep; the warp is an 8-tap gather through L1/L2.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "common.cuh"
#include "vpa.h"

namespace u3d {

// ------------------------------------------------------------------------------------------------
// host: plan
// ------------------------------------------------------------------------------------------------
namespace {

struct Rng {
    std::mt19937 g;
    float lo, hi;
    Rng(float a, float b, uint32_t seed) : g(seed), lo(a), hi(b) {}
    float operator()() {
        float u = float(uint32_t(g())) / 4294967296.0f;
        if (u >= 1.0f) u = std::nextafterf(1.0f, 0.0f);
        float r = (hi - lo) * u;
        return r + lo;
    }
};

void affine_matrix(const float t[3], const float r[3], const float s[3], int W, int H, int D, float M[12]) {
    const double cx = std::cos(double(r[0])), sx = std::sin(double(r[0]));
    const double cy = std::cos(double(r[1])), sy = std::sin(double(r[1]));
    const double cz = std::cos(double(r[2])), sz = std::sin(double(r[2]));
    const double Rx[3][3] = {{1, 0, 0}, {0, cx, -sx}, {0, sx, cx}};
    const double Ry[3][3] = {{cy, 0, sy}, {0, 1, 0}, {-sy, 0, cy}};
    const double Rz[3][3] = {{cz, -sz, 0}, {sz, cz, 0}, {0, 0, 1}};
    double T[3][3], R[3][3], A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T[i][j] = 0;
            for (int k = 0; k < 3; ++k) T[i][j] += Rz[i][k] * Ry[k][j];
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            R[i][j] = 0;
            for (int k = 0; k < 3; ++k) R[i][j] += T[i][k] * Rx[k][j];
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = R[i][j] * double(s[j]);
    const double c[3] = {W * 0.5, H * 0.5, D * 0.5};
    for (int i = 0; i < 3; ++i) {
        double b = c[i] + double(t[i]);
        for (int j = 0; j < 3; ++j) b -= A[i][j] * c[j];
        for (int j = 0; j < 3; ++j) M[i * 4 + j] = float(A[i][j]);
        M[i * 4 + 3] = float(b);
    }
}

}  // namespace

int vpa_make_plan(const char* const* keys, const float* vals, int n_opts, int is_label, int W, int H, int D, int C, uint64_t seed64,
                  VpaPlan& P) {
    std::memset(&P, 0, sizeof(P));
    if (C < 1 || C > kVpaMaxC) { set_error("vpa: 1..8 image channels supported"); return 1; }
    std::map<std::string, float> options;
    for (int i = 0; i < n_opts; ++i) options[keys[i]] = vals[i];
    auto opt = [&](const char* k) -> float {  // unordered_map::operator[] semantics: missing key reads as 0
        auto it = options.find(k);
        return it == options.end() ? 0.f : it->second;
    };
    const uint32_t seed = uint32_t(seed64);
    P.W = W; P.H = H; P.D = D; P.C = C; P.is_label = is_label; P.seed = seed;
    Rng one(-1.0f, 1.0f, seed);
    auto range = [&](float from, float to) {
        float a = one() * (to - from);
        a = a * 0.5f;
        float b = (to + from) * 0.5f;
        return a + b;
    };
    auto apply = [&](const char* name) {
        const int index = int(opt(name));
        if (index == 0) return false;
        if (index >= 4) return true;
        return std::fabs(one()) < float(index) * 0.25f;
    };
    auto random_location = [&](float from, float to, int loc[3]) {
        loc[0] = int(float(W - 1) * range(from, to));
        loc[1] = int(float(H - 1) * range(from, to));
        loc[2] = int(float(D - 1) * range(from, to));
    };
    const int maxdim = std::max(W, std::max(H, D));
    // downsample (:205-220)
    const bool dsx = apply("downsample_x"), dsy = apply("downsample_y"), dsz = apply("downsample_z");
    if (dsx || dsy || dsz) {
        P.ds = 1;
        P.lw = int(float(W) * (dsx ? opt("downsample_x_ratio") : 1.0f));
        P.lh = int(float(H) * (dsy ? opt("downsample_y_ratio") : 1.0f));
        P.ld = int(float(D) * (dsz ? opt("downsample_z_ratio") : 1.0f));
        if (P.lw < 1 || P.lh < 1 || P.ld < 1) { set_error("vpa: downsample ratio yields an empty volume"); return 1; }
    }
    // cropping (:222-230)
    if (apply("cropping")) {
        P.crop = 1;
        const float size = range(opt("cropping_size_min"), opt("cropping_size_max")) * float(W);
        P.crop_value = range(0.0f, 2.0f);
        random_location(size, 1.0f - size, P.crop_loc);
        P.crop_r = int(size);
    }
    // truncation (:231-250)
    if (apply("truncation_z")) {
        P.trunc = 1;
        float a = one() * 0.5f;
        P.top = int(std::fabs(a * float(D)));
        float b = one() * 0.5f;
        P.bot = int(std::fabs(b * float(D)));
    }
    if (apply("noise")) { P.noise = 1; P.noise_mag = opt("noise_mag"); P.noise_mt = opt("noise_mt19937") != 0.f ? 1 : 0; }
    if (apply("ambient")) { P.ambient = 1; P.ambient_add = range(0.0f, 1.0f) * opt("ambient_mag"); }
    if (apply("diffuse")) {
        P.diffuse = 1;
        float d0 = range(-0.5f, 0.5f), d1 = range(-0.5f, 0.5f), d2 = range(-0.5f, 0.5f);
        const float nrm = float(std::sqrt(double(d0) * d0 + double(d1) * d1 + double(d2) * d2));
        d0 /= nrm; d1 /= nrm; d2 /= nrm;
        const float k = opt("diffuse_mag") / float(maxdim);
        P.diff_f[0] = d0 * k; P.diff_f[1] = d1 * k; P.diff_f[2] = d2 * k;
    }
    if (apply("specular")) {
        P.specular = 1;
        random_location(0.4f, 0.6f, P.spec_loc);
        P.spec_mag = opt("specular_mag");
        P.spec_b = 1.0f - P.spec_mag - P.spec_mag;
        P.spec_freq = float(double(opt("specular_freq")) * (std::acos(-1.0) * 0.5 / maxdim));
    }
    // rigid motion + view port (:280-336)
    {
        const float resolution = range(1.0f / opt("scaling_up"), 1.0f / opt("scaling_down"));
        const float tr = opt("translocation_ratio");
        float t[3], r[3], s[3];
        t[0] = one() * tr * float(W); t[1] = one() * tr * float(H); t[2] = one() * tr * float(D);
        r[0] = one() * opt("rotation_x"); r[1] = one() * opt("rotation_y"); r[2] = one() * opt("rotation_z");
        const float asp = opt("aspect_ratio");
        for (int i = 0; i < 3; ++i) s[i] = resolution * range(1.0f / asp, asp);
        affine_matrix(t, r, s, W, H, D, P.M);
        P.persp[0] = range(-0.5f, 0.5f) * opt("perspective") / float(W);
        P.persp[1] = range(-0.5f, 0.5f) * opt("perspective") / float(H);
        P.persp[2] = range(-0.5f, 0.5f) * opt("perspective") / float(D);
        P.has_persp = opt("perspective") > 0.0f;
        P.use_disp = opt("lens_distortion") > 0.0f;
        if (opt("lens_distortion") != 0.0f) {
            P.has_lens = 1;
            const float lens_mag = range(0.0f, 1.0f) * opt("lens_distortion");
            const float radius = float(maxdim / 2);
            P.lens_k = -(lens_mag / (radius * radius));
        }
        if (apply("distortion")) {
            const int num = int(range(1.0f, opt("distortion_count") + 1.0f));
            for (int i = 0; i < num && P.nfoci < kVpaMaxFoci; ++i) {
                VpaFocus& f = P.foci[P.nfoci++];
                random_location(0.3f, 0.7f, f.loc);
                f.radius = float(W) * range(opt("distortion_radius_min"), opt("distortion_radius_max"));
                f.mag = range(opt("distortion_mag_min"), opt("distortion_mag_max"));
                f.coef = -(f.radius * f.mag);
                f.pir = float(std::acos(-1.0) / double(f.radius));
                f.ri = int(f.radius);
            }
        }
    }
    // background (:345-425)
    if (is_label) {
        if (apply("zero_background")) {
            P.zero_bg = 1;
            return 0;
        }
        if (apply("rubber_stamping")) {
            P.rubber = 1;
            const float pi2 = float(std::acos(-1.0) * 2.0);
            for (int it = 0; it < 5; ++it) {
                float t[3], r[3], s[3];
                t[0] = one() * float(W) * 0.5f; t[1] = one() * float(H) * 0.5f; t[2] = one() * float(D) * 0.5f;
                r[0] = one() * pi2; r[1] = one() * pi2; r[2] = one() * pi2;
                s[0] = range(0.8f, 1.25f); s[1] = range(0.8f, 1.25f); s[2] = range(0.8f, 1.25f);
                affine_matrix(t, r, s, W, H, D, P.rubberM[it]);
            }
            for (int c = 0; c < C; ++c)
                for (int it = 0; it < 5; ++it) P.rubber_upper[c][it] = range(0.0f, 1.0f) * opt("rubber_stamping_mag");
        }
        if (apply("perlin_texture")) {
            P.perlin = 1;
            for (int i = 0; i < 512; ++i) P.perm[i] = i & 255;
            // the reference's own call (visual_perception_augmentation.cpp:392); this host code is built against the same libstdc++
            std::shuffle(P.perm, P.perm + 512, std::mt19937(uint32_t(seed)));
            P.zoom = range(0.005f, 0.05f);
            P.perlin_upper = range(0.0f, 1.0f) * opt("perlin_texture_mag");
        }
        P.final_norm = 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

struct Tri {
    int x0, x1, y0, y1, z0, z1;
    float fx, fy, fz;
    bool valid;
};

__device__ __forceinline__ Tri locate(float px, float py, float pz, int W, int H, int D, bool clamp) {
    Tri t;
    if (clamp) {
        px = fminf(fmaxf(px, 0.f), float(W - 1)); py = fminf(fmaxf(py, 0.f), float(H - 1)); pz = fminf(fmaxf(pz, 0.f), float(D - 1));
        t.valid = true;
    } else {
        t.valid = px >= 0.f && px <= float(W - 1) && py >= 0.f && py <= float(H - 1) && pz >= 0.f && pz <= float(D - 1);
        if (!t.valid) { px = py = pz = 0.f; }
    }
    const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
    t.x0 = int(flx); t.y0 = int(fly); t.z0 = int(flz);
    t.fx = px - flx; t.fy = py - fly; t.fz = pz - flz;
    t.x1 = min(t.x0 + 1, W - 1); t.y1 = min(t.y0 + 1, H - 1); t.z1 = min(t.z0 + 1, D - 1);
    return t;
}

// weights in the oracle's order: (wz*wy)*wx, taps z0y0x0, z0y0x1, z0y1x0, ...
template <typename F>
__device__ __forceinline__ float tri_sample(const Tri& t, int W, int H, F fetch) {
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int z = a ? t.z1 : t.z0;
        const float wz = a ? t.fz : 1.f - t.fz;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int y = b ? t.y1 : t.y0;
            const float wy = b ? t.fy : 1.f - t.fy;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int x = c ? t.x1 : t.x0;
                const float wx = c ? t.fx : 1.f - t.fx;
                acc = acc + fetch((size_t(z) * H + y) * W + x) * (__fmul_rn(__fmul_rn(wz, wy), wx));
            }
        }
    }
    return acc;
}

__global__ void k_scale(const float* __restrict__ src, float* __restrict__ dst, int sW, int sH, int sD, int dW, int dH, int dD) {
    const long long n = 1LL * dW * dH * dD;
    const float rx = float(sW) / float(dW), ry = float(sH) / float(dH), rz = float(sD) / float(dD);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % dW);
        const long long q = i / dW;
        const int y = int(q % dH), z = int(q / dH);
        const Tri t = locate(__fmul_rn(float(x), rx), __fmul_rn(float(y), ry), __fmul_rn(float(z), rz), sW, sH, sD, true);
        dst[i] = tri_sample(t, sW, sH, [&](size_t o) { return src[o]; });
    }
}

// std::mt19937(seed) output words 0..n-1 (tempered), for the bit-exact noise stream of the reference CPU path
// (visual_perception_augmentation.cpp:254-257: ONE uniform_dist drawn voxel after voxel, channel after channel).  The recurrence is
// sequential across 624-word blocks; inside a block word i needs the OLD words i, i+1 and, for i < 227, the old word i+397, else the
// NEW word i-227: three dependent phases [0,227) [227,454) [454,624) of one CTA.  ~1.2 ms for a 160x192x160 volume on the prefetch
// stream; the default noise is the counter-based hash (no sequential dependency).
__global__ void __launch_bounds__(256) k_mt19937_words(uint32_t* __restrict__ out, long long n, uint32_t seed) {
    __shared__ uint32_t mt[624];
    __shared__ uint32_t nw[624];
    if (threadIdx.x == 0) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + uint32_t(i);
    }
    __syncthreads();
    auto twist = [](uint32_t a, uint32_t b) {
        const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
        return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    };
    const int t = threadIdx.x;
    for (long long base = 0; base < n; base += 624) {
        if (t < 227) nw[t] = mt[t + 397] ^ twist(mt[t], mt[t + 1]);
        __syncthreads();
        if (t < 227) nw[227 + t] = nw[t] ^ twist(mt[227 + t], mt[228 + t]);
        __syncthreads();
        if (t < 169) nw[454 + t] = nw[227 + t] ^ twist(mt[454 + t], mt[455 + t]);
        else if (t == 169) nw[623] = nw[396] ^ twist(mt[623], nw[0]);
        __syncthreads();
        for (int i = t; i < 624; i += 256) {
            uint32_t y = nw[i];
            mt[i] = y;
            y ^= y >> 11;
            y ^= (y << 7) & 0x9D2C5680u;
            y ^= (y << 15) & 0xEFC60000u;
            y ^= y >> 18;
            if (base + i < n) out[base + i] = y;
        }
        __syncthreads();
    }
}

// crop / truncation / noise / ambient / diffuse / specular, in the reference's order, in place
__global__ void k_pre(float* __restrict__ img, float* __restrict__ lab, const uint32_t* __restrict__ mt_words, const __grid_constant__ VpaPlan P) {
    const long long V = 1LL * P.W * P.H * P.D;
    const uint32_t key = hash32(P.seed);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % P.W);
        const long long q = i / P.W;
        const int y = int(q % P.H), z = int(q / P.H);
        float l = lab[i];
        bool crop_hit = false;
        if (P.crop && l != 0.f && abs(x - P.crop_loc[0]) <= P.crop_r && abs(y - P.crop_loc[1]) <= P.crop_r &&
            abs(z - P.crop_loc[2]) <= P.crop_r) {
            crop_hit = true;   // only the first channel sees the label before it is cleared (.cpp:228-229)
            l = 0.f;
        }
        const bool trunc = P.trunc && (z >= P.D - P.top || z < P.bot);
        if (trunc) l = 0.f;
        lab[i] = l;
        float g_diff = 1.f, g_spec = 1.f;
        if (P.diffuse) {
            const float d = (float(x) - float(P.W * 0.5)) * P.diff_f[0] + (float(y) - float(P.H * 0.5)) * P.diff_f[1] +
                            (float(z) - float(P.D * 0.5)) * P.diff_f[2];
            g_diff = fmaxf(0.f, 1.f + d);
        }
        if (P.specular) {
            const float ex = float(x) - float(P.spec_loc[0]), ey = float(y) - float(P.spec_loc[1]), ez = float(z) - float(P.spec_loc[2]);
            const float dist = sqrtf(ex * ex + ey * ey + ez * ez);
            g_spec = (cosf(dist * P.spec_freq) + 1.f) * P.spec_mag + P.spec_b;
        }
        for (int c = 0; c < P.C; ++c) {
            float v = img[c * V + i];
            if (c == 0 && crop_hit) v = P.crop_value;
            if (trunc) v = 0.f;
            if (P.noise && P.noise_mt) {
                // libstdc++ uniform_real_distribution<float>(0, mag): u = float(word) / 2^32 clipped below 1, value = mag * u
                float u = __fmul_rn(__uint2float_rn(mt_words[c * V + i]), 2.3283064365386963e-10f);
                if (u >= 1.0f) u = 0.99999994f;
                v = __fadd_rn(v, __fmul_rn(P.noise_mag, u));
            } else if (P.noise) {
                const uint32_t h = hash32(uint32_t(c * V + i) ^ key);
                v += float(h >> 8) * (1.0f / 16777216.0f) * P.noise_mag;
            }
            if (P.ambient) v += P.ambient_add;
            if (P.diffuse) v *= g_diff;
            if (P.specular) v *= g_spec;
            img[c * V + i] = v;
        }
    }
}

__device__ __forceinline__ void affine(const float* M, float& x, float& y, float& z) {
    const float ox = M[0] * x + M[1] * y + M[2] * z + M[3];
    const float oy = M[4] * x + M[5] * y + M[6] * z + M[7];
    const float oz = M[8] * x + M[9] * y + M[10] * z + M[11];
    x = ox; y = oy; z = oz;
}

__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {  // v >= 0
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ float block_max(float v) {
    __shared__ float s[32];
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) r = warp_max(r);
    __syncthreads();
    return r;  // valid in warp 0
}

// the warp: displacement (lens + local foci) -> perspective divide -> affine -> gathers; clamps at 0, tracks channel maxima
__global__ void k_warp(const float* __restrict__ img, const float* __restrict__ lab, float* __restrict__ out,
                       float* __restrict__ out_lab, float* __restrict__ chmax, const __grid_constant__ VpaPlan P) {
    const long long V = 1LL * P.W * P.H * P.D;
    float mx[kVpaMaxC];
#pragma unroll
    for (int c = 0; c < kVpaMaxC; ++c) mx[c] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % P.W);
        const long long q = i / P.W;
        const int y = int(q % P.H), z = int(q / P.H);
        float px = float(x), py = float(y), pz = float(z);
        if (P.use_disp) {
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (P.has_lens) {
                const float ex = px - float(P.W / 2), ey = py - float(P.H / 2), ez = pz - float(P.D / 2);
                const float k = P.lens_k * (ex * ex + ey * ey + ez * ez);
                dx = ex * k; dy = ey * k; dz = ez * k;
            }
            for (int f = 0; f < P.nfoci; ++f) {
                const VpaFocus& F = P.foci[f];
                if (abs(x - F.loc[0]) > F.ri || abs(y - F.loc[1]) > F.ri || abs(z - F.loc[2]) > F.ri) continue;
                const float ex = px - float(F.loc[0]), ey = py - float(F.loc[1]), ez = pz - float(F.loc[2]);
                const float len = sqrtf(ex * ex + ey * ey + ez * ez);
                if (len > F.radius || len <= 0.f) continue;
                const float coef = F.coef * sinf(len * F.pir) / len;
                dx += ex * coef; dy += ey * coef; dz += ez * coef;
            }
            px += dx; py += dy; pz += dz;
        }
        if (P.has_persp) {
            const float den = P.persp[0] * (px - float(P.W / 2.0)) + P.persp[1] * (py - float(P.H / 2.0)) +
                              P.persp[2] * (pz - float(P.D / 2.0)) + 1.f;
            px /= den; py /= den; pz /= den;
        }
        affine(P.M, px, py, pz);
        const Tri t = locate(px, py, pz, P.W, P.H, P.D, false);
        float lv = 0.f;
        if (t.valid) {
            if (P.is_label) {
                // majority: label with the largest summed trilinear weight, first in z,y,x tap order on ties
                float vals[8], wts[8];
                int n = 0;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b)
                        for (int c = 0; c < 2; ++c, ++n) {
                            const int zz = a ? t.z1 : t.z0, yy = b ? t.y1 : t.y0, xx = c ? t.x1 : t.x0;
                            vals[n] = lab[(size_t(zz) * P.H + yy) * P.W + xx];
                            wts[n] = __fmul_rn(__fmul_rn(a ? t.fz : 1.f - t.fz, b ? t.fy : 1.f - t.fy), c ? t.fx : 1.f - t.fx);
                        }
                float best = vals[0], best_w = -1.f;
                for (int a = 0; a < 8; ++a) {
                    float tot = 0.f;
                    for (int b = 0; b < 8; ++b) tot = tot + (vals[b] == vals[a] ? wts[b] : 0.f);
                    if (tot > best_w) { best_w = tot; best = vals[a]; }
                }
                lv = best;
            } else
                lv = tri_sample(t, P.W, P.H, [&](size_t o) { return lab[o]; });
        }
        out_lab[i] = lv;
        for (int c = 0; c < P.C; ++c) {
            float v = 0.f;
            if (t.valid) v = fmaxf(tri_sample(t, P.W, P.H, [&](size_t o) { return img[c * V + o]; }), 0.f);
            out[c * V + i] = v;
            mx[c] = fmaxf(mx[c], v);
        }
    }
    for (int c = 0; c < P.C; ++c) {
        const float r = block_max(mx[c]);
        if (threadIdx.x == 0) atomic_max_pos(chmax + c, r);
    }
}

// out[c] *= upper/max[c] (normalize), optional zero-background (preserve), optional new channel maxima
__global__ void k_normalize(float* __restrict__ out, const float* __restrict__ out_lab, const float* __restrict__ chmax, int C,
                            long long V, int zero_bg) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const bool keep = !zero_bg || out_lab[i] != 0.f;
        for (int c = 0; c < C; ++c) {
            const float m = chmax[c];
            float v = out[c * V + i];
            if (m != 0.f) v *= 1.0f / m;
            out[c * V + i] = keep ? v : 0.f;
        }
    }
}

// rubber stamping: background = resample(image masked where label != 0, T), clamped at 0, with its maximum
__global__ void k_rubber_bg(const float* __restrict__ img, const float* __restrict__ lab, float* __restrict__ bg, float* __restrict__ bgmax,
                            const float* M12, int W, int H, int D) {
    __shared__ float M[12];
    if (threadIdx.x < 12) M[threadIdx.x] = M12[threadIdx.x];
    __syncthreads();
    const long long V = 1LL * W * H * D;
    float mx = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        float px = float(int(i % W));
        const long long q = i / W;
        float py = float(int(q % H)), pz = float(int(q / H));
        affine(M, px, py, pz);
        const Tri t = locate(px, py, pz, W, H, D, false);
        float v = 0.f;
        if (t.valid) v = fmaxf(tri_sample(t, W, H, [&](size_t o) { return lab[o] != 0.f ? 0.f : img[o]; }), 0.f);
        bg[i] = v;
        mx = fmaxf(mx, v);
    }
    const float r = block_max(mx);
    if (threadIdx.x == 0) atomic_max_pos(bgmax, r);
}

__device__ __forceinline__ float fade(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
__device__ __forceinline__ float lerpf(float t, float a, float b) { return a + t * (b - a); }
__device__ __forceinline__ float gradp(int hash, float x, float y, float z) {
    const int h = hash & 15;
    const float u = h < 8 ? x : y;
    const float v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
    return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
__device__ float perlin3(float x, float y, float z, const int* p) {
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const int xi = int(fx0) & 255, yi = int(fy0) & 255, zi = int(fz0) & 255;
    const float xf = x - fx0, yf = y - fy0, zf = z - fz0;
    const float u = fade(xf), v = fade(yf), w = fade(zf);
    const int A = p[xi] + yi, B = p[xi + 1] + yi;
    const int aaa = p[p[A] + zi], aba = p[p[A + 1] + zi], aab = p[p[A] + zi + 1], abb = p[p[A + 1] + zi + 1];
    const int baa = p[p[B] + zi], bba = p[p[B + 1] + zi], bab = p[p[B] + zi + 1], bbb = p[p[B + 1] + zi + 1];
    float x1 = lerpf(u, gradp(aaa, xf, yf, zf), gradp(baa, xf - 1, yf, zf));
    float x2 = lerpf(u, gradp(aba, xf, yf - 1, zf), gradp(bba, xf - 1, yf - 1, zf));
    const float y1 = lerpf(v, x1, x2);
    x1 = lerpf(u, gradp(aab, xf, yf, zf - 1), gradp(bab, xf - 1, yf, zf - 1));
    x2 = lerpf(u, gradp(abb, xf, yf - 1, zf - 1), gradp(bbb, xf - 1, yf - 1, zf - 1));
    const float y2 = lerpf(v, x1, x2);
    return lerpf(w, y1, y2);
}

__global__ void k_perlin_bg(float* __restrict__ bg, float* __restrict__ bgmax, const int* __restrict__ perm, float zoom, int W, int H, int D) {
    __shared__ int p[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) p[i] = perm[i];
    __syncthreads();
    const long long V = 1LL * W * H * D;
    float mx = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const float x = float(int(i % W));
        const long long q = i / W;
        const float y = float(int(q % H)), z = float(int(q / H));
        float acc = 0.f, po = 1.f;
        for (int o = 0; o < 4; ++o, po *= 0.5f) {
            const float sc = __fmul_rn(zoom, po);
            acc = acc + __fmul_rn(perlin3(__fmul_rn(x, sc), __fmul_rn(y, sc), __fmul_rn(z, sc), p), po);
        }
        float v = acc * 2.0f;
        v = v - floorf(v);
        bg[i] = v;
        mx = fmaxf(mx, v);
    }
    const float r = block_max(mx);
    if (threadIdx.x == 0) atomic_max_pos(bgmax, r);
}

// where the warped label is background: out += (bg*upper/bgmax) * max(0.1, 1 - out); optionally track the new maximum
__global__ void k_blend(float* __restrict__ out, const float* __restrict__ out_lab, const float* __restrict__ bg,
                        const float* __restrict__ bgmax, float upper, float* __restrict__ newmax, long long V) {
    const float m = *bgmax;
    const float s = m != 0.f ? upper / m : 1.f;
    float mx = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        float v = out[i];
        if (out_lab[i] == 0.f) {
            const float b = m != 0.f ? bg[i] * s : bg[i];
            v += b * fmaxf(0.1f, 1.0f - v);
            out[i] = v;
        }
        mx = fmaxf(mx, fmaxf(v, 0.f));
    }
    if (newmax != nullptr) {
        const float r = block_max(mx);
        if (threadIdx.x == 0) atomic_max_pos(newmax, r);
    }
}

__global__ void k_chan_max(const float* __restrict__ out, float* __restrict__ chmax, long long V) {
    const int c = blockIdx.y;
    float mx = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x)
        mx = fmaxf(mx, out[c * V + i]);
    const float r = block_max(mx);
    if (threadIdx.x == 0) atomic_max_pos(chmax + c, r);
}

__global__ void k_clamp0(float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fmaxf(out[i], 0.f);
}

inline int vgrid(long long n) {
    long long g = (n + 255) / 256;
    return int(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

size_t vpa_workspace_bytes(int W, int H, int D, int C) {
    const size_t V = size_t(W) * H * D;
    // work image [C][V], work label, out [C][V], out label, bg, low-res temp (<= V), maxima + perm + matrices
    return (size_t(2) * C + 4) * V * 4 + 8192;
}

// image/label: DEVICE fp32 buffers ([C][D][H][W] and [D][H][W]); augmented in place.  Returns the launch count via *launches.
int vpa_run(const VpaPlan& P, float* image, float* label, void* workspace, cudaStream_t s, long long* launches) {
    const long long V = 1LL * P.W * P.H * P.D;
    const int C = P.C;
    float* ws = static_cast<float*>(workspace);
    float* wimg = ws;                 // [C][V]
    float* wlab = wimg + C * V;       // [V]
    float* out = wlab + V;            // [C][V]
    float* olab = out + C * V;        // [V]
    float* bg = olab + V;             // [V]
    float* low = bg + V;              // [V]
    float* small = low + V;           // maxima etc.
    float* chmax = small;             // [8]
    float* chmax2 = small + 8;        // [8]
    float* bgmax = small + 16;        // [8]
    float* mats = small + 32;         // 5 x 12
    int* perm = reinterpret_cast<int*>(small + 128);  // 512
    long long nl = 0;
    U3D_CUDA_CHECK(cudaMemsetAsync(small, 0, 128 * 4, s));
    U3D_CUDA_CHECK(cudaMemcpyAsync(wlab, label, size_t(V) * 4, cudaMemcpyDeviceToDevice, s));
    if (P.ds) {
        for (int c = 0; c < C; ++c) {
            k_scale<<<vgrid(1LL * P.lw * P.lh * P.ld), 256, 0, s>>>(image + c * V, low, P.W, P.H, P.D, P.lw, P.lh, P.ld);
            k_scale<<<vgrid(V), 256, 0, s>>>(low, wimg + c * V, P.lw, P.lh, P.ld, P.W, P.H, P.D);
            nl += 2;
        }
    } else
        U3D_CUDA_CHECK(cudaMemcpyAsync(wimg, image, size_t(C) * V * 4, cudaMemcpyDeviceToDevice, s));
    if (P.crop || P.trunc || P.noise || P.ambient || P.diffuse || P.specular) {
        uint32_t* mt_words = nullptr;
        if (P.noise && P.noise_mt) {   // `out` is not written before k_warp: stage the C*V generator words there
            mt_words = reinterpret_cast<uint32_t*>(out);
            k_mt19937_words<<<1, 256, 0, s>>>(mt_words, C * V, P.seed);
            ++nl;
        }
        k_pre<<<vgrid(V), 256, 0, s>>>(wimg, wlab, mt_words, P);
        ++nl;
    }
    k_warp<<<vgrid(V), 256, 0, s>>>(wimg, wlab, out, olab, chmax, P);
    k_normalize<<<vgrid(V), 256, 0, s>>>(out, olab, chmax, C, V, P.zero_bg);
    nl += 2;
    if (P.is_label && !P.zero_bg) {
        if (P.rubber) {
            U3D_CUDA_CHECK(cudaMemcpyAsync(mats, P.rubberM, sizeof(P.rubberM), cudaMemcpyHostToDevice, s));
            for (int c = 0; c < C; ++c)
                for (int it = 0; it < 5; ++it) {
                    U3D_CUDA_CHECK(cudaMemsetAsync(bgmax, 0, 4, s));
                    k_rubber_bg<<<vgrid(V), 256, 0, s>>>(wimg + c * V, wlab, bg, bgmax, mats + it * 12, P.W, P.H, P.D);
                    k_blend<<<vgrid(V), 256, 0, s>>>(out + c * V, olab, bg, bgmax, P.rubber_upper[c][it], nullptr, V);
                    nl += 2;
                }
        }
        if (P.perlin) {
            U3D_CUDA_CHECK(cudaMemcpyAsync(perm, P.perm, sizeof(P.perm), cudaMemcpyHostToDevice, s));
            U3D_CUDA_CHECK(cudaMemsetAsync(bgmax, 0, 4, s));
            k_perlin_bg<<<vgrid(V), 256, 0, s>>>(bg, bgmax, perm, P.zoom, P.W, P.H, P.D);
            ++nl;
            for (int c = 0; c < C; ++c) {
                k_blend<<<vgrid(V), 256, 0, s>>>(out + c * V, olab, bg, bgmax, P.perlin_upper, nullptr, V);
                ++nl;
            }
        }
        if (P.final_norm) {
            k_clamp0<<<vgrid(C * V), 256, 0, s>>>(out, C * V);
            k_chan_max<<<dim3(vgrid(V), C), 256, 0, s>>>(out, chmax2, V);
            k_normalize<<<vgrid(V), 256, 0, s>>>(out, olab, chmax2, C, V, 0);
            nl += 3;
        }
    }
    U3D_CUDA_CHECK(cudaMemcpyAsync(image, out, size_t(C) * V * 4, cudaMemcpyDeviceToDevice, s));
    U3D_CUDA_CHECK(cudaMemcpyAsync(label, olab, size_t(V) * 4, cudaMemcpyDeviceToDevice, s));
    U3D_CUDA_CHECK(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

}  // namespace u3d
