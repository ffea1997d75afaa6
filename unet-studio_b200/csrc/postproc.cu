// Inference pre/post-processing around the window forward: window extraction from a volume, softmax,
// window re-assembly, create_mask and argmax on the GPU, so that evaluate returns a 1-byte label map (+ foreground probability)
// instead of out_count x 4 bytes of logits per voxel.
//
// Reference: the default postproc string is "softmax+create_mask+argmax" (/root/reference/unet.cpp:112); it is executed by
// tipl::ml3d::evalution_set::run_postproc (evaluate.cpp:274) and the windows (model_io) are cut and re-assembled by
// handle_fov_pre / handle_fov_post (evaluate.cpp:201-204,274) -- all inside TIPL, which is not vendored: PARITY UNPINNED.
// The one piece of that arithmetic visible in the reference is the argmax call of postproc_actions (evaluate.cpp:315-319):
//     label = tipl::argmax(prob4d, shape, mask > threshold)         -> arg-max channel inside the mask, 0 outside.
// Assumed here (restated in oracle/postproc_oracle.py, which the GPU tests compare against):
//   softmax      p_c = exp(l_c - max) / sum over the out_count channels of a voxel (channel 0 = background: training labels are
//                0 = background and out_count = max(label) + 1, train.cpp:1125)
//   create_mask  fg_prob = 1 - p_0 = sum_{c >= 1} p_c
//   argmax       label = fg_prob > threshold ? first arg-max channel of p : 0
//   windows      a volume larger than the model grid is covered by windows of the model grid at a given stride per axis, the last
//                window of an axis shifted inward to end at the border; a smaller volume is zero-padded at the far end
//                ("align_top"); probabilities of overlapping windows are averaged before create_mask / argmax.
// All kernels are bandwidth-bound: one thread per voxel, channel planes read coalesced.
#include <string>
#include <vector>

#include "common.cuh"
#include "postproc.h"

namespace u3d {
namespace {

inline int pgrid(long long n) {
    const long long g = (n + 255) / 256;
    return int(g < 1 ? 1 : (g > 148LL * 32 ? 148LL * 32 : g));
}

// window (ww x wh x wd at origin ox,oy,oz) of every input channel of a volume; outside the volume reads 0
__global__ void k_crop_window(const float* __restrict__ vol, float* __restrict__ win, int C, int vw, int vh, int vd, int ww, int wh, int wd,
                              int ox, int oy, int oz) {
    const long long WV = 1LL * ww * wh * wd, VV = 1LL * vw * vh * vd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < WV; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % ww);
        const long long q = i / ww;
        const int y = int(q % wh), z = int(q / wh);
        const int sx = x + ox, sy = y + oy, sz = z + oz;
        const bool in = sx < vw && sy < vh && sz < vd;
        const long long s = (1LL * sz * vh + sy) * vw + sx;
        for (int c = 0; c < C; ++c) win[c * WV + i] = in ? vol[c * VV + s] : 0.f;
    }
}

// softmax of one window's logits ([C][wd][wh][ww] planar fp32) added into the volume accumulators acc [C][vd][vh][vw]; cnt counts the
// windows that covered a voxel.  Voxels of the window outside the volume (zero padding) are dropped.
template <int MAXC>
__global__ void k_softmax_accumulate(const float* __restrict__ logits, float* __restrict__ acc, float* __restrict__ cnt, int C, int vw, int vh,
                                     int vd, int ww, int wh, int wd, int ox, int oy, int oz) {
    const long long WV = 1LL * ww * wh * wd, VV = 1LL * vw * vh * vd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < WV; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % ww);
        const long long q = i / ww;
        const int y = int(q % wh), z = int(q / wh);
        const int sx = x + ox, sy = y + oy, sz = z + oz;
        if (sx >= vw || sy >= vh || sz >= vd) continue;
        float l[MAXC];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { l[c] = logits[c * WV + i]; mx = fmaxf(mx, l[c]); }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { l[c] = __expf(l[c] - mx); sum += l[c]; }
        const float inv = 1.f / sum;
        const long long s = (1LL * sz * vh + sy) * vw + sx;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) acc[c * VV + s] += l[c] * inv;
        cnt[s] += 1.f;
    }
}

// create_mask + argmax over the assembled probabilities; optionally writes the averaged probabilities back (label_prob)
template <int MAXC>
__global__ void k_mask_argmax(float* __restrict__ acc, const float* __restrict__ cnt, uint8_t* __restrict__ label, float* __restrict__ fg,
                              int C, long long V, float threshold, int write_prob) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const float n = cnt ? cnt[i] : 1.f;
        const float inv = n > 0.f ? 1.f / n : 0.f;
        float best = -1.f, p0 = 0.f;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float p = acc[c * V + i] * inv;
                if (c == 0) p0 = p;
                if (p > best) { best = p; arg = c; }
                if (write_prob) acc[c * V + i] = p;
            }
        const float f = n > 0.f ? 1.f - p0 : 0.f;
        if (fg) fg[i] = f;
        label[i] = (f > threshold) ? uint8_t(arg) : uint8_t(0);
    }
}

// nearest / trilinear resampling of a [C][sd][sh][sw] volume to [C][dd][dh][dw] with tipl::scale-style index mapping
// (destination index * src_dim / dst_dim, clamped): the resample step in front of the windows
__global__ void k_resample(const float* __restrict__ src, float* __restrict__ dst, int C, int sw, int sh, int sd, int dw, int dh, int dd,
                           int nearest) {
    const long long DV = 1LL * dw * dh * dd, SV = 1LL * sw * sh * sd;
    const float rx = float(sw) / float(dw), ry = float(sh) / float(dh), rz = float(sd) / float(dd);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < DV; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % dw);
        const long long q = i / dw;
        const int y = int(q % dh), z = int(q / dh);
        const float fx = fminf(__fmul_rn(float(x), rx), float(sw - 1)), fy = fminf(__fmul_rn(float(y), ry), float(sh - 1)),
                    fz = fminf(__fmul_rn(float(z), rz), float(sd - 1));
        if (nearest) {
            const int ix = int(fx + 0.5f) < sw ? int(fx + 0.5f) : sw - 1, iy = int(fy + 0.5f) < sh ? int(fy + 0.5f) : sh - 1,
                      iz = int(fz + 0.5f) < sd ? int(fz + 0.5f) : sd - 1;
            for (int c = 0; c < C; ++c) dst[c * DV + i] = src[c * SV + (1LL * iz * sh + iy) * sw + ix];
            continue;
        }
        const int x0 = int(fx), y0 = int(fy), z0 = int(fz);
        const int x1 = x0 + 1 < sw ? x0 + 1 : sw - 1, y1 = y0 + 1 < sh ? y0 + 1 : sh - 1, z1 = z0 + 1 < sd ? z0 + 1 : sd - 1;
        const float ax = fx - float(x0), ay = fy - float(y0), az = fz - float(z0);
        for (int c = 0; c < C; ++c) {
            const float* s = src + c * SV;
            auto at = [&](int zz, int yy, int xx) { return s[(1LL * zz * sh + yy) * sw + xx]; };
            const float c00 = at(z0, y0, x0) * (1.f - ax) + at(z0, y0, x1) * ax, c01 = at(z0, y1, x0) * (1.f - ax) + at(z0, y1, x1) * ax;
            const float c10 = at(z1, y0, x0) * (1.f - ax) + at(z1, y0, x1) * ax, c11 = at(z1, y1, x0) * (1.f - ax) + at(z1, y1, x1) * ax;
            dst[c * DV + i] = (c00 * (1.f - ay) + c01 * ay) * (1.f - az) + (c10 * (1.f - ay) + c11 * ay) * az;
        }
    }
}

}  // namespace

std::vector<int> window_origins(int vdim, int wdim, int stride) {
    std::vector<int> o;
    if (vdim <= wdim) { o.push_back(0); return o; }
    if (stride < 1) stride = wdim;
    for (int p = 0; p + wdim < vdim; p += stride) o.push_back(p);
    o.push_back(vdim - wdim);   // the last window ends at the border
    return o;
}

int crop_window_launch(const float* vol, float* win, int C, int vw, int vh, int vd, int ww, int wh, int wd, int ox, int oy, int oz,
                       cudaStream_t s) {
    k_crop_window<<<pgrid(1LL * ww * wh * wd), 256, 0, s>>>(vol, win, C, vw, vh, vd, ww, wh, wd, ox, oy, oz);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int softmax_accumulate_launch(const float* logits, float* acc, float* cnt, int C, int vw, int vh, int vd, int ww, int wh, int wd, int ox,
                              int oy, int oz, cudaStream_t s) {
    if (C < 1 || C > kPostMaxC) { set_error("postproc supports 1.." + std::to_string(kPostMaxC) + " output channels"); return 1; }
    const int g = pgrid(1LL * ww * wh * wd);
    if (C <= 8) k_softmax_accumulate<8><<<g, 256, 0, s>>>(logits, acc, cnt, C, vw, vh, vd, ww, wh, wd, ox, oy, oz);
    else k_softmax_accumulate<kPostMaxC><<<g, 256, 0, s>>>(logits, acc, cnt, C, vw, vh, vd, ww, wh, wd, ox, oy, oz);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int mask_argmax_launch(float* acc, const float* cnt, uint8_t* label, float* fg, int C, long long V, float threshold, int write_prob,
                       cudaStream_t s) {
    if (C < 1 || C > kPostMaxC) { set_error("postproc supports 1.." + std::to_string(kPostMaxC) + " output channels"); return 1; }
    if (C <= 8) k_mask_argmax<8><<<pgrid(V), 256, 0, s>>>(acc, cnt, label, fg, C, V, threshold, write_prob);
    else k_mask_argmax<kPostMaxC><<<pgrid(V), 256, 0, s>>>(acc, cnt, label, fg, C, V, threshold, write_prob);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int resample_launch(const float* src, float* dst, int C, int sw, int sh, int sd, int dw, int dh, int dd, int nearest, cudaStream_t s) {
    k_resample<<<pgrid(1LL * dw * dh * dd), 256, 0, s>>>(src, dst, C, sw, sh, sd, dw, dh, dd, nearest);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
