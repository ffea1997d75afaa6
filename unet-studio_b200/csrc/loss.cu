// Fused loss head: softmax + cross-entropy + soft-Dice + MSE(Brier) with valid mask and optional
// `collapse_before` logsumexp merge, forward and gradient, for one deep-supervision level.
// Restates calc_losses (/root/reference/train.cpp:501-552) and the target down-sampling of
// train.cpp:645-662 (nearest interpolate by exact halving == take voxel (z<<k, y<<k, x<<k)).
// Two passes over the logits because the Dice gradient needs the global per-class sums:
//   pass 1: per-voxel softmax, warp-shuffle + shared reduction of {sum ce*v, sum v, sum mse*v, I_c, K_c}
//   pass 2: recompute softmax, write dL/dlogits (x loss_scale) as fp16 NDHWC for the head's dgrad/wgrad.
#include <string>

#include "common.cuh"
#include "elementwise.h"

namespace u3d {
namespace {

constexpr int kMaxC = 32;
constexpr int kAccN = 3 + 2 * kMaxC;

template <int MAXC>
struct Voxel {
    float s[MAXC];     // softmax over the (collapsed) classes
    float wi[MAXC];    // softmax inside the collapsed group (first `cb` original logits), only if collapse
    int t;             // collapsed target
    float v;           // valid
    float lse;         // log-sum-exp of collapsed logits
    float lt;          // collapsed logit of the target
};

template <int MAXC>
__device__ __forceinline__ void eval_voxel(const LossLevel& L, long long vox, int x, int y, int z, Voxel<MAXC>& o) {
    const long long nv = (long long)L.d * L.h * L.w;
    const int cb = L.collapse_before;
    const int Cc = cb ? L.C - cb + 1 : L.C;
    float l[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) l[c] = 0.f;
    if (cb) {
        float mx = -INFINITY;
        for (int c = 0; c < cb; ++c) mx = fmaxf(mx, L.logits[c * nv + vox]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < cb) {
                o.wi[c] = expf(L.logits[c * nv + vox] - mx);
                sum += o.wi[c];
            }
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < cb) o.wi[c] *= inv;
        l[0] = mx + logf(sum);
#pragma unroll
        for (int c = 1; c < MAXC; ++c)
            if (c < Cc) l[c] = L.logits[(cb + c - 1) * nv + vox];
    } else {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < Cc) l[c] = L.logits[c * nv + vox];
    }
    const long long li = ((long long)(z << L.shift) * L.H0 + (y << L.shift)) * L.W0 + (x << L.shift);
    const long long traw = (long long)L.label[li];
    const bool valid = traw < L.C;
    long long t = traw;
    if (cb) t = t - cb + 1 < 0 ? 0 : t - cb + 1;
    if (!valid || t < 0) t = 0;
    o.t = int(t);
    o.v = valid ? 1.f : 0.f;
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < Cc) mx = fmaxf(mx, l[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        o.s[c] = c < Cc ? expf(l[c] - mx) : 0.f;
        sum += o.s[c];
    }
    const float inv = 1.f / sum;
    o.lt = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        o.s[c] *= inv;
        if (c == o.t) o.lt = l[c];
    }
    o.lse = mx + logf(sum);
}

__device__ __forceinline__ float clampp(float s) { return fminf(fmaxf(s, 1e-6f), 1.0f - 1e-6f); }

template <int MAXC>
__global__ void loss_reduce_kernel(const LossLevel L) {
    const long long nv = (long long)L.d * L.h * L.w;
    const int Cc = L.collapse_before ? L.C - L.collapse_before + 1 : L.C;
    float a_ce = 0.f, a_n = 0.f, a_mse = 0.f;
    float a_i[MAXC], a_k[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) a_i[c] = a_k[c] = 0.f;
    for (long long vox = blockIdx.x * (long long)blockDim.x + threadIdx.x; vox < nv; vox += (long long)gridDim.x * blockDim.x) {
        const int x = int(vox % L.w);
        const long long q = vox / L.w;
        const int y = int(q % L.h), z = int(q / L.h);
        Voxel<MAXC> o;
        eval_voxel<MAXC>(L, vox, x, y, z, o);
        a_n += o.v;
        a_ce += (o.lse - o.lt) * o.v;
        float pp = 0.f, pt = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < Cc) {
                const float p = clampp(o.s[c]);
                pp += p * p;
                if (c == o.t) pt = p;
                if (c >= 1) {
                    const float m = (c == o.t) ? o.v : 0.f;
                    a_i[c] += p * o.v * m;
                    a_k[c] += p * o.v + m;
                }
            }
        a_mse += (pp - 2.f * pt + 1.f) * o.v;
    }
    __shared__ float red[kAccN];
    for (int i = threadIdx.x; i < kAccN; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    a_ce = warp_sum(a_ce); a_n = warp_sum(a_n); a_mse = warp_sum(a_mse);
#pragma unroll
    for (int c = 1; c < MAXC; ++c)
        if (c < Cc) { a_i[c] = warp_sum(a_i[c]); a_k[c] = warp_sum(a_k[c]); }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&red[0], a_ce); atomicAdd(&red[1], a_n); atomicAdd(&red[2], a_mse);
#pragma unroll
        for (int c = 1; c < MAXC; ++c)
            if (c < Cc) { atomicAdd(&red[3 + 2 * c], a_i[c]); atomicAdd(&red[4 + 2 * c], a_k[c]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 + 2 * Cc + 2; i += blockDim.x)
        if (i < kAccN && red[i] != 0.f) atomicAdd(&L.acc[i], double(red[i]));
}

__global__ void loss_finalize_kernel(const LossLevel L) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int Cc = L.collapse_before ? L.C - L.collapse_before + 1 : L.C;
    const double n = L.acc[1] > 1.0 ? L.acc[1] : 1.0;
    double dice_sum = 0;
    const double eps = double(1e-5f);
    for (int c = 1; c < Cc; ++c) dice_sum += (2.0 * L.acc[3 + 2 * c] + eps) / (L.acc[4 + 2 * c] + eps);
    const double Z = double(Cc - 1 > 1 ? Cc - 1 : 1);
    L.out3[0] = float(L.acc[0] / n);
    L.out3[1] = float(1.0 - dice_sum / Z);
    L.out3[2] = float(L.acc[2] / n);
}

template <int MAXC>
__global__ void loss_grad_kernel(const LossLevel L) {
    const long long nv = (long long)L.d * L.h * L.w;
    const int cb = L.collapse_before;
    const int Cc = cb ? L.C - cb + 1 : L.C;
    __shared__ float sI[MAXC], sK[MAXC];
    __shared__ float s_n;
    if (threadIdx.x < MAXC) {
        sI[threadIdx.x] = threadIdx.x < Cc ? float(L.acc[3 + 2 * threadIdx.x]) : 0.f;
        sK[threadIdx.x] = threadIdx.x < Cc ? float(L.acc[4 + 2 * threadIdx.x]) : 0.f;
    }
    if (threadIdx.x == 0) s_n = float(L.acc[1] > 1.0 ? L.acc[1] : 1.0);
    __syncthreads();
    const float inv_n = 1.f / s_n;
    const float eps = 1e-5f;
    const float invZ = 1.f / float(Cc - 1 > 1 ? Cc - 1 : 1);
    __half* out = static_cast<__half*>(L.dlogits);
    for (long long vox = blockIdx.x * (long long)blockDim.x + threadIdx.x; vox < nv; vox += (long long)gridDim.x * blockDim.x) {
        const int x = int(vox % L.w);
        const long long q = vox / L.w;
        const int y = int(q % L.h), z = int(q / L.h);
        Voxel<MAXC> o;
        eval_voxel<MAXC>(L, vox, x, y, z, o);
        float g[MAXC];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            g[c] = 0.f;
            if (c < Cc) {
                const float s = o.s[c];
                const float p = clampp(s);
                const float hit = (c == o.t) ? 1.f : 0.f;
                float G = L.w_mse * o.v * (2.f * p - 2.f * hit) * inv_n;
                if (c >= 1) {
                    const float den = sK[c] + eps;
                    G -= L.w_dice * invZ * o.v * (2.f * hit * den - (2.f * sI[c] + eps)) / (den * den);
                }
                const bool pass = s >= 1e-6f && s <= 1.0f - 1e-6f;  // clamp passes gradient on the closed interval
                g[c] = pass ? G : 0.f;
                dot += g[c] * s;
            }
        }
        float dl[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            dl[c] = 0.f;
            if (c < Cc) {
                const float hit = (c == o.t) ? 1.f : 0.f;
                dl[c] = (o.s[c] * (g[c] - dot) + L.w_ce * o.v * (o.s[c] - hit) * inv_n) * L.loss_scale;
            }
        }
        // un-collapse and store: original channel j
        __half* row = out + vox * L.Cp;
        for (int j = 0; j < L.Cp; ++j) {
            float val = 0.f;
            if (j < L.C) {
                if (cb) {
                    if (j < cb) {
                        float wj = 0.f;
#pragma unroll
                        for (int c = 0; c < MAXC; ++c)
                            if (c == j) wj = o.wi[c];
                        val = dl[0] * wj;
                    } else {
#pragma unroll
                        for (int c = 1; c < MAXC; ++c)
                            if (c == j - cb + 1) val = dl[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < MAXC; ++c)
                        if (c == j) val = dl[c];
                }
            }
            row[j] = __float2half_rn(val);
        }
    }
}

}  // namespace

int loss_level_launch(const LossLevel& L, cudaStream_t s) {
    if (L.C > kMaxC || L.C < 1) {
        set_error("loss head supports 1..32 output channels");
        return 1;
    }
    if (L.collapse_before < 0 || L.collapse_before >= L.C) {
        set_error("invalid collapse_before");  // train.cpp:507-508
        return 1;
    }
    U3D_CUDA_CHECK(cudaMemsetAsync(L.acc, 0, sizeof(double) * kAccN, s));
    const long long nv = (long long)L.d * L.h * L.w;
    long long g = (nv + 255) / 256;
    const int grid = int(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    if (L.C <= 8) loss_reduce_kernel<8><<<grid, 256, 0, s>>>(L);
    else loss_reduce_kernel<kMaxC><<<grid, 256, 0, s>>>(L);
    loss_finalize_kernel<<<1, 32, 0, s>>>(L);
    if (L.dlogits != nullptr) {
        if (L.C <= 8) loss_grad_kernel<8><<<grid, 256, 0, s>>>(L);
        else loss_grad_kernel<kMaxC><<<grid, 256, 0, s>>>(L);
    }
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
