// Bandwidth-bound kernels on NDHWC fp16 tensors (channel pitch Cp, multiple of 16): per-channel statistics,
// InstanceNorm3d / BatchNorm3d apply + ReLU / LeakyReLU(0.01) / ELU forward and backward
// (/root/reference/unet.cpp:74-98), MaxPool3d(2,2) with indices and nearest Upsample x2 (unet.cpp:38-44).
// One thread moves one 16-byte chunk (8 channels of one voxel); reductions keep a fixed channel chunk per
// thread so no atomics are needed until the per-block partial rows.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "elementwise.h"

namespace u3d {
namespace {

__device__ __forceinline__ void load8(const uint4* p, float (&f)[8]) {
    const uint4 q = *p;
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = unpack2<false>(w[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = unpack2<false>(w[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ void store8(uint4* p, const float (&f)[8]) {
    uint4 q;
    q.x = pack2<false>(f[0], f[1]);
    q.y = pack2<false>(f[2], f[3]);
    q.z = pack2<false>(f[4], f[5]);
    q.w = pack2<false>(f[6], f[7]);
    *p = q;
}

__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(z, 0.f);
        case ACT_LEAKY: return z > 0.f ? z : 0.01f * z;
        case ACT_ELU: return z > 0.f ? z : expm1f(z);
        default: return z;
    }
}
__device__ __forceinline__ float act_grad(float z, int act) {
    switch (act) {
        case ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case ACT_LEAKY: return z > 0.f ? 1.f : 0.01f;
        case ACT_ELU: return z > 0.f ? 1.f : expf(z);
        default: return 1.f;
    }
}

// ---------------------------------------------------------------------------------------------
// statistics finalize: partial rows [rows][2][ntot] -> mean / rstd (and BatchNorm running stats)
// ---------------------------------------------------------------------------------------------
// one warp per channel: lanes stride over the partial rows, double accumulation, shuffle reduce
__global__ void finalize_stats_kernel(const float* __restrict__ partials, int rows, int ntot, int C, double count, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd, float* running_mean,
                                      float* running_var, float momentum) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double a = 0, q = 0;
    for (int r = lane; r < rows; r += 32) {
        a += partials[size_t(r) * 2 * ntot + c];
        q += partials[size_t(r) * 2 * ntot + ntot + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane != 0) return;
    const double m = a / count;
    double var = q / count - m * m;
    if (var < 0) var = 0;
    mean[c] = float(m);
    rstd[c] = float(1.0 / sqrt(var + double(eps)));
    if (running_mean != nullptr) {  // torch BatchNorm: running_var uses the unbiased estimate
        const double unb = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * float(m);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(unb);
    }
}

// BatchNorm3d in eval mode (running statistics, eps 0: unet.cpp:80-84): rstd[c] = 1/sqrt(running_var[c])
__global__ void rstd_from_var_kernel(const float* __restrict__ var, float* __restrict__ rstd, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) rstd[c] = rsqrtf(var[c] + eps);
}

// Generic two-value per-channel reduction over a [V][Cp] tensor.  MODE 0: (x, x^2).
// MODE 1 (norm/act backward): dz = dy*act'(z), xhat = (x-mean)*rstd -> (dz, dz*xhat).
struct ReduceArgs {
    const uint4* x;      // raw conv output (MODE 0: the tensor itself)
    const uint4* dy;     // MODE 1 only
    const float* mean;
    const float* rstd;
    const float* gamma;
    const float* beta;
    float* partials;     // [grid][2][Cp]
    long long V;
    int C, Cp, has_norm, act;
    // MODE 1 with counter != nullptr: the block that finishes last also reduces the partial rows (fixed order => deterministic) into
    // sums[2][Cp] and adds them to the gamma / beta gradients -- no separate finalize launch
    unsigned int* counter;
    float* sums;
    float* dgamma;
    float* dbeta;
};

template <int MODE>
__global__ void channel_reduce_kernel(const ReduceArgs a) {
    extern __shared__ float sm[];
    const int nch = a.Cp / 8;
    const int k = blockDim.x / nch;
    const int t = threadIdx.x;
    const int ch = t % nch, vsub = t / nch;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    float sc[8], sh[8], mu[8], rs[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            const bool real = c < a.C;
            mu[j] = (a.has_norm && real) ? a.mean[c] : 0.f;
            rs[j] = (a.has_norm && real) ? a.rstd[c] : 1.f;
            const float g = (a.has_norm && real) ? a.gamma[c] : 1.f;
            const float b = (a.has_norm && real) ? a.beta[c] : 0.f;
            sc[j] = g * rs[j];
            sh[j] = b - mu[j] * sc[j];
        }
    }
    if (vsub < k) {
        const long long vstride = (long long)gridDim.x * k;
        for (long long v = (long long)blockIdx.x * k + vsub; v < a.V; v += 4 * vstride) {
            uint4 rx[4], rd[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long vu = v + u * vstride;
                ok[u] = vu < a.V;
                if (ok[u]) {
                    rx[u] = a.x[vu * nch + ch];
                    if (MODE == 1) rd[u] = a.dy[vu * nch + ch];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                float x[8];
                unpack8(rx[u], x);
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s0[j] += x[j];
                        s1[j] += x[j] * x[j];
                    }
                } else {
                    float d[8];
                    unpack8(rd[u], d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float z = sc[j] * x[j] + sh[j];
                        const float dz = d[j] * act_grad(z, a.act);
                        s0[j] += dz;
                        s1[j] += dz * (x[j] - mu[j]) * rs[j];
                    }
                }
            }
        }
    }
    // block reduce over vsub
    float* row = sm + size_t(vsub) * a.Cp * 2;
    if (vsub < k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            row[ch * 8 + j] = s0[j];
            row[a.Cp + ch * 8 + j] = s1[j];
        }
    }
    __syncthreads();
    for (int i = t; i < 2 * a.Cp; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < k; ++r) acc += sm[size_t(r) * a.Cp * 2 + i];
        a.partials[size_t(blockIdx.x) * 2 * a.Cp + i] = acc;
    }
    if constexpr (MODE == 1) {
        if (a.counter == nullptr) return;
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (t == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const int ncol = 2 * a.Cp, rows = int(gridDim.x);
        double* red = reinterpret_cast<double*>(sm);   // blockDim.x doubles fit: the dynamic buffer holds 16 floats per thread
        auto emit = [&](int col, double tot) {
            a.sums[col] = float(tot);
            const int c = col < a.Cp ? col : col - a.Cp;
            if (c < a.C) {
                if (col < a.Cp) { if (a.dbeta) a.dbeta[c] += float(tot); }
                else if (a.dgamma) a.dgamma[c] += float(tot);
            }
        };
        if (ncol <= int(blockDim.x)) {
            const int G = int(blockDim.x) / ncol;        // row groups; thread (g, col) walks rows g, g+G, ...
            const int col = t % ncol, g = t / ncol;
            double acc = 0;
            if (g < G) {
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;   // four independent chains in a fixed pattern (deterministic)
                int r = g;
                for (; r + 3 * G < rows; r += 4 * G) {
                    a0 += double(__ldcg(a.partials + size_t(r) * ncol + col));
                    a1 += double(__ldcg(a.partials + size_t(r + G) * ncol + col));
                    a2 += double(__ldcg(a.partials + size_t(r + 2 * G) * ncol + col));
                    a3 += double(__ldcg(a.partials + size_t(r + 3 * G) * ncol + col));
                }
                for (; r < rows; r += G) a0 += double(__ldcg(a.partials + size_t(r) * ncol + col));
                acc = (a0 + a1) + (a2 + a3);
            }
            red[t] = acc;
            __syncthreads();
            if (g == 0) {
                double tot = 0;
                for (int q = 0; q < G; ++q) tot += red[q * ncol + col];
                emit(col, tot);
            }
        } else {
            for (int col = t; col < ncol; col += blockDim.x) {
                double tot = 0;
#pragma unroll 4
                for (int r = 0; r < rows; ++r) tot += double(__ldcg(a.partials + size_t(r) * ncol + col));
                emit(col, tot);
            }
        }
        if (t == 0) *a.counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// y = act(scale*x + shift), scale/shift from (mean, rstd, gamma, beta); padded channels stay 0
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
    const uint4* x;
    uint4* y;
    const float* mean;
    const float* rstd;
    const float* gamma;
    const float* beta;
    long long V;
    int C, Cp, has_norm, act;
};

__global__ void norm_act_fwd_kernel(const ApplyArgs a) {
    extern __shared__ float sm[];
    float* sc = sm;
    float* sh = sm + a.Cp;
    for (int c = threadIdx.x; c < a.Cp; c += blockDim.x) {
        float s = 1.f, b = 0.f;
        if (c >= a.C) s = 0.f;
        else if (a.has_norm) {
            const float r = a.rstd ? a.rstd[c] : 1.f;
            const float m = a.mean ? a.mean[c] : 0.f;
            s = a.gamma[c] * r;
            b = a.beta[c] - m * s;
        }
        sc[c] = s;
        sh[c] = b;
    }
    __syncthreads();
    const int nch = a.Cp / 8;
    const long long total = a.V * nch;
    // two 16-byte chunks per thread and iteration (independent loads in flight); the channel group of a chunk is tracked with
    // 32-bit adds instead of a 64-bit modulo per element
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int smod = int(stride % nch);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int ch = int(i % nch);
    if (smod == 0) {
        float csc[8], csh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            csc[j] = c < a.C ? sc[c] : 0.f;
            csh[j] = c < a.C ? sh[c] : 0.f;
        }
        const int act = a.act;
        for (; i < total; i += 2 * stride) {
            const long long i2 = i + stride;
            const bool has2 = i2 < total;
            float x[8], y[8];
            load8(a.x + i, x);
            if (has2) load8(a.x + i2, y);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = act_fwd(csc[j] * x[j] + csh[j], act);
            store8(a.y + i, x);
            if (has2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = act_fwd(csc[j] * y[j] + csh[j], act);
                store8(a.y + i2, y);
            }
        }
        return;
    }
    for (; i < total; i += 2 * stride) {
        const long long i2 = i + stride;
        int ch2 = ch + smod; if (ch2 >= nch) ch2 -= nch;
        const bool has2 = i2 < total;
        float x[8], y[8];
        load8(a.x + i, x);
        if (has2) load8(a.x + i2, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            x[j] = c < a.C ? act_fwd(sc[c] * x[j] + sh[c], a.act) : 0.f;
        }
        store8(a.y + i, x);
        if (has2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ch2 * 8 + j;
                y[j] = c < a.C ? act_fwd(sc[c] * y[j] + sh[c], a.act) : 0.f;
            }
            store8(a.y + i2, y);
        }
        ch = ch2 + smod; if (ch >= nch) ch -= nch;
    }
}

// ---------------------------------------------------------------------------------------------
// backward apply: dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)), dz = dy*act'(z)
// sums [2][Cp] = finalized (sum dz, sum dz*xhat); without norm: dx = dz.
// ---------------------------------------------------------------------------------------------
struct BwdApplyArgs {
    const uint4* x;
    const uint4* dy;
    uint4* dx;
    const float* mean;
    const float* rstd;
    const float* gamma;
    const float* beta;
    const float* sums;   // [2][Cp] totals
    long long V;
    int C, Cp, has_norm, act;
};

template <bool FAST, int NCH, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_act_bwd_apply_kernel(const BwdApplyArgs a) {
    extern __shared__ float sm[];
    float* sc = sm;
    float* sh = sc + a.Cp;
    float* mu = sh + a.Cp;
    float* rs = mu + a.Cp;
    float* m1 = rs + a.Cp;   // mean(dz)
    float* m2 = m1 + a.Cp;   // mean(dz*xhat)
    for (int c = threadIdx.x; c < a.Cp; c += blockDim.x) {
        const bool real = c < a.C;
        const bool nrm = a.has_norm && real;
        mu[c] = nrm ? a.mean[c] : 0.f;
        rs[c] = nrm ? a.rstd[c] : 1.f;
        const float g = nrm ? a.gamma[c] : 1.f;
        const float b = nrm ? a.beta[c] : 0.f;
        sc[c] = g * rs[c];
        sh[c] = b - mu[c] * sc[c];
        m1[c] = nrm ? a.sums[c] / float(a.V) : 0.f;
        m2[c] = nrm ? a.sums[a.Cp + c] / float(a.V) : 0.f;
    }
    __syncthreads();
    const int nch = a.Cp / 8;
    const long long total = a.V * nch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int smod = int(stride % nch);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int ch = int(i % nch);
    if constexpr (FAST) {
        // the grid stride is a multiple of the channel groups: every chunk of this thread has the same 8 channels -> coefficients
        // live in registers (the shared-memory version issued 96 LDS per iteration and was LSU-bound, not HBM-bound)
        // dx = sc*(dz - m1 - xhat*m2) with xhat = (x - mu)*rs  ==  G*dz + P*x + Q, four coefficients per channel instead of six
        // (G = sc with a norm, 1 without, 0 for a padded channel): fewer registers -> three blocks per SM instead of two
        // (sc = gamma*rstd with a norm, 1 without; 0 here for a padded channel, whose output must be 0)
        float csc[8], csh[8], cp[8], cq[8];
        const bool hn = a.has_norm != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            const bool real = c < a.C;
            csc[j] = real ? sc[c] : 0.f; csh[j] = sh[c];
            cp[j] = (real && hn) ? -sc[c] * m2[c] * rs[c] : 0.f;
            cq[j] = (real && hn) ? sc[c] * (m2[c] * rs[c] * mu[c] - m1[c]) : 0.f;
        }
        const int act = a.act;
        // NCH chunks per iteration: the 2*NCH 16-byte loads are issued before the first result is needed
        for (; i < total; i += NCH * stride) {
            uint4 rx[NCH], rd[NCH];
            bool ok[NCH];
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
                const long long iu = i + u * stride;
                ok[u] = iu < total;
                if (ok[u]) { rx[u] = a.x[iu]; rd[u] = a.dy[iu]; }
            }
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
                if (!ok[u]) continue;
                float x[8], d[8];
                unpack8(rx[u], x);
                unpack8(rd[u], d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float z = csc[j] * x[j] + csh[j];
                    const float dz = d[j] * act_grad(z, act);
                    d[j] = csc[j] * dz + (cp[j] * x[j] + cq[j]);
                }
                store8(a.dx + i + u * stride, d);
            }
        }
        return;
    } else {
    auto one = [&](int chq, float (&x)[8], float (&d)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = chq * 8 + j;
            if (c < a.C) {
                const float z = sc[c] * x[j] + sh[c];
                const float dz = d[j] * act_grad(z, a.act);
                const float xh = (x[j] - mu[c]) * rs[c];
                d[j] = a.has_norm ? sc[c] * (dz - m1[c] - xh * m2[c]) : dz;
            } else
                d[j] = 0.f;
        }
    };
    for (; i < total; i += 2 * stride) {
        const long long i2 = i + stride;
        int ch2 = ch + smod; if (ch2 >= nch) ch2 -= nch;
        const bool has2 = i2 < total;
        float x[8], d[8], x2[8], d2[8];
        load8(a.x + i, x);
        load8(a.dy + i, d);
        if (has2) { load8(a.x + i2, x2); load8(a.dy + i2, d2); }
        one(ch, x, d);
        store8(a.dx + i, d);
        if (has2) {
            one(ch2, x2, d2);
            store8(a.dx + i2, d2);
        }
        ch = ch2 + smod; if (ch >= nch) ch -= nch;
    }
    }
}

// sums partial rows into totals [2][Cp]; optionally accumulates into the gamma/beta gradients.  ONE block: thread (g, col) walks the
// rows g, g+G, ... of column col (coalesced 2*Cp-float rows, independent loads in flight), then a fixed-order tree over the G row
// groups in shared memory (deterministic).  The previous warp-per-channel version spent ~15 us per call on dependent strided loads.
__global__ void __launch_bounds__(1024) finalize_bwd_sums_kernel(const float* __restrict__ partials, int rows, int Cp, int C,
                                                                 float* __restrict__ sums, float* dgamma, float* dbeta) {
    __shared__ double red[1024];
    const int ncol = 2 * Cp;                     // <= 512
    const int G = 1024 / ncol;                   // row groups (>= 2)
    const int col = threadIdx.x % ncol, g = threadIdx.x / ncol;
    double a = 0;
    if (g < G) {
#pragma unroll 4
        for (int r = g; r < rows; r += G) a += double(partials[size_t(r) * ncol + col]);
    }
    red[threadIdx.x] = (g < G) ? a : 0.0;
    __syncthreads();
    if (g == 0) {
        double t = 0;
        for (int k = 0; k < G; ++k) t += red[k * ncol + col];
        sums[col] = float(t);
        const int c = col < Cp ? col : col - Cp;
        if (c < C) {
            if (col < Cp) { if (dbeta) dbeta[c] += float(t); }
            else if (dgamma) dgamma[c] += float(t);
        }
    }
}

// per-channel sum of a [V][Cp] tensor added into out[c] (bias gradients of convs without a norm behind them)
__global__ void finalize_colsum_kernel(const float* __restrict__ partials, int rows, int Cp, int C, float* out) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double a = 0;
    for (int r = lane; r < rows; r += 32) a += partials[size_t(r) * 2 * Cp + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[c] += float(a);
}

// ---------------------------------------------------------------------------------------------
// MaxPool3d(2,2) / Upsample(nearest x2)
// ---------------------------------------------------------------------------------------------
// torch tie rule (aten/src/ATen/native/cuda/DilatedMaxPool3d.cu): scan d,h,w in order, take val if (val > max) || isnan(val)
__global__ void maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int* __restrict__ idx, int Cp, int od,
                                   int oh, int ow) {
    const int nch = Cp / 8;
    const long long total = (long long)od * oh * ow * nch;
    const int ih = oh * 2, iw = ow * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % nch);
        long long v = i / nch;
        const int ox = int(v % ow); v /= ow;
        const int oy = int(v % oh);
        const int oz = int(v / oh);
        float best[8];
        int bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = ((oz * 2) * ih + oy * 2) * iw + ox * 2; }
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const int flat = ((oz * 2 + a) * ih + (oy * 2 + b)) * iw + (ox * 2 + c);
                    float f[8];
                    load8(x + (long long)flat * nch + ch, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (f[j] > best[j] || f[j] != f[j]) { best[j] = f[j]; bi[j] = flat; }
                }
        store8(y + i, best);
        int4* ip = reinterpret_cast<int4*>(idx + i * 8);
        ip[0] = make_int4(bi[0], bi[1], bi[2], bi[3]);
        ip[1] = make_int4(bi[4], bi[5], bi[6], bi[7]);
    }
}

__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dy, const int* __restrict__ idx, uint4* __restrict__ dx, int Cp,
                                   int od, int oh, int ow) {
    const int nch = Cp / 8;
    const long long total = (long long)od * oh * ow * nch;
    const int ih = oh * 2, iw = ow * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % nch);
        long long v = i / nch;
        const int ox = int(v % ow); v /= ow;
        const int oy = int(v % oh);
        const int oz = int(v / oh);
        float d[8];
        load8(dy + i, d);
        int bi[8];
        const int4* ip = reinterpret_cast<const int4*>(idx + i * 8);
        const int4 i0 = ip[0], i1 = ip[1];
        bi[0] = i0.x; bi[1] = i0.y; bi[2] = i0.z; bi[3] = i0.w; bi[4] = i1.x; bi[5] = i1.y; bi[6] = i1.z; bi[7] = i1.w;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const int flat = ((oz * 2 + a) * ih + (oy * 2 + b)) * iw + (ox * 2 + c);
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = bi[j] == flat ? d[j] : 0.f;
                    store8(dx + (long long)flat * nch + ch, f);
                }
    }
}

__global__ void upsample_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int Cp, int id, int ih, int iw) {
    const int nch = Cp / 8;
    const int oh = ih * 2, ow = iw * 2;
    const long long total = (long long)id * 2 * oh * ow * nch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % nch);
        long long v = i / nch;
        const int ox = int(v % ow); v /= ow;
        const int oy = int(v % oh);
        const int oz = int(v / oh);
        y[i] = x[((long long)((oz >> 1) * ih + (oy >> 1)) * iw + (ox >> 1)) * nch + ch];
    }
}

__global__ void upsample_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int Cp, int id, int ih, int iw) {
    const int nch = Cp / 8;
    const int oh = ih * 2, ow = iw * 2;
    const long long total = (long long)id * ih * iw * nch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % nch);
        long long v = i / nch;
        const int x = int(v % iw); v /= iw;
        const int y = int(v % ih);
        const int z = int(v / ih);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    float f[8];
                    load8(dy + ((long long)((z * 2 + a) * oh + (y * 2 + b)) * ow + (x * 2 + c)) * nch + ch, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += f[j];
                }
        store8(dx + i, acc);
    }
}

__global__ void add16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        load8(dst + i, a);
        load8(src + i, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += b[j];
        store8(dst + i, a);
    }
}

inline int ew_grid(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 8;
    return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int reduce_rows_max() { return 148 * 2; }   // measured: 2 blocks per SM balance the reduce kernel against its single-block finalize

static int reduce_launch(int mode, const ReduceArgs& a, int* rows, cudaStream_t s) {
    const int nch = a.Cp / 8;
    if (nch > 1024) { set_error("channel count too large"); return 1; }
    const int k = nch >= 256 ? 1 : 256 / nch;
    const int block = nch * k;
    long long want = (a.V + k - 1) / k;
    int grid = int(want < 1 ? 1 : (want > reduce_rows_max() ? reduce_rows_max() : want));
    if (mode == 1 && a.counter != nullptr) {
        // the block that finishes last sums the partial rows: keep its per-thread chain (rows * 2*Cp / block) at ~80 loads.  The wide
        // layers are the small deep-level tensors, which a few dozen blocks stream in a microsecond anyway.
        const int cap = std::max(8, 80 * block / (2 * a.Cp));
        grid = std::min(grid, cap);
    }
    const size_t smem = size_t(k) * a.Cp * 2 * sizeof(float);
    if (mode == 0) channel_reduce_kernel<0><<<grid, block, smem, s>>>(a);
    else channel_reduce_kernel<1><<<grid, block, smem, s>>>(a);
    *rows = grid;
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int channel_stats_launch(const void* x, long long V, int C, int Cp, float* partials, int* rows, cudaStream_t s) {
    ReduceArgs a{};
    a.x = static_cast<const uint4*>(x); a.V = V; a.C = C; a.Cp = Cp; a.partials = partials;
    return reduce_launch(0, a, rows, s);
}

int finalize_stats_launch(const float* partials, int rows, int ntot, int C, double count, float eps, float* mean, float* rstd,
                          float* running_mean, float* running_var, float momentum, cudaStream_t s) {
    finalize_stats_kernel<<<(C * 32 + 127) / 128, 128, 0, s>>>(partials, rows, ntot, C, count, eps, mean, rstd, running_mean,
                                                          running_var, momentum);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int rstd_from_var_launch(const float* var, float* rstd, int C, float eps, cudaStream_t s) {
    rstd_from_var_kernel<<<(C + 127) / 128, 128, 0, s>>>(var, rstd, C, eps);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int norm_act_fwd_launch(const void* x, void* y, long long V, int C, int Cp, int has_norm, int act, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, cudaStream_t s) {
    ApplyArgs a{};
    a.x = static_cast<const uint4*>(x); a.y = static_cast<uint4*>(y); a.V = V; a.C = C; a.Cp = Cp;
    a.has_norm = has_norm; a.act = act; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta;
    norm_act_fwd_kernel<<<ew_grid(V * (Cp / 8), 256), 256, size_t(2) * Cp * sizeof(float), s>>>(a);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int norm_act_bwd_launch(const void* x, const void* dy, void* dx, long long V, int C, int Cp, int has_norm, int act,
                        const float* mean, const float* rstd, const float* gamma, const float* beta, float* partials,
                        float* sums, float* dgamma, float* dbeta, cudaStream_t s, unsigned int* counter) {
    if (has_norm) {
        ReduceArgs r{};
        r.x = static_cast<const uint4*>(x); r.dy = static_cast<const uint4*>(dy); r.V = V; r.C = C; r.Cp = Cp;
        r.has_norm = 1; r.act = act; r.mean = mean; r.rstd = rstd; r.gamma = gamma; r.beta = beta; r.partials = partials;
        r.counter = counter; r.sums = sums; r.dgamma = dgamma; r.dbeta = dbeta;
        int rows = 0;
        if (reduce_launch(1, r, &rows, s)) return 1;
        if (counter == nullptr) {
            if (2 * Cp > 1024) { set_error("norm_act_bwd_launch: more than 512 padded channels"); return 1; }
            finalize_bwd_sums_kernel<<<1, 1024, 0, s>>>(partials, rows, Cp, C, sums, dgamma, dbeta);
            U3D_CUDA_CHECK(cudaGetLastError());
        }
    }
    BwdApplyArgs a{};
    a.x = static_cast<const uint4*>(x); a.dy = static_cast<const uint4*>(dy); a.dx = static_cast<uint4*>(dx);
    a.V = V; a.C = C; a.Cp = Cp; a.has_norm = has_norm; a.act = act;
    a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.sums = sums;
    const int grid = ew_grid(V * (Cp / 8), 256);
    // (measured: 2 / 3 / 4 resident blocks per SM and 2 - 4 chunks in flight per thread give the same 7.86 ms step)
    const size_t sm = size_t(6) * Cp * sizeof(float);
    if ((1LL * grid * 256) % (Cp / 8) != 0) norm_act_bwd_apply_kernel<false, 2, 2><<<grid, 256, sm, s>>>(a);
    else norm_act_bwd_apply_kernel<true, 4, 2><<<grid, 256, sm, s>>>(a);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int colsum_accumulate_launch(const void* x, long long V, int C, int Cp, float* partials, float* out, cudaStream_t s) {
    int rows = 0;
    if (channel_stats_launch(x, V, C, Cp, partials, &rows, s)) return 1;
    finalize_colsum_kernel<<<(C * 32 + 127) / 128, 128, 0, s>>>(partials, rows, Cp, C, out);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int maxpool_fwd_launch(const void* x, void* y, int* idx, int Cp, int od, int oh, int ow, cudaStream_t s) {
    maxpool_fwd_kernel<<<ew_grid(1LL * od * oh * ow * (Cp / 8), 256), 256, 0, s>>>(static_cast<const uint4*>(x),
                                                                                    static_cast<uint4*>(y), idx, Cp, od, oh, ow);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int maxpool_bwd_launch(const void* dy, const int* idx, void* dx, int Cp, int od, int oh, int ow, cudaStream_t s) {
    maxpool_bwd_kernel<<<ew_grid(1LL * od * oh * ow * (Cp / 8), 256), 256, 0, s>>>(static_cast<const uint4*>(dy), idx,
                                                                                    static_cast<uint4*>(dx), Cp, od, oh, ow);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int upsample_fwd_launch(const void* x, void* y, int Cp, int id, int ih, int iw, cudaStream_t s) {
    upsample_fwd_kernel<<<ew_grid(8LL * id * ih * iw * (Cp / 8), 256), 256, 0, s>>>(static_cast<const uint4*>(x),
                                                                                     static_cast<uint4*>(y), Cp, id, ih, iw);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int upsample_bwd_launch(const void* dy, void* dx, int Cp, int id, int ih, int iw, cudaStream_t s) {
    upsample_bwd_kernel<<<ew_grid(1LL * id * ih * iw * (Cp / 8), 256), 256, 0, s>>>(static_cast<const uint4*>(dy),
                                                                                     static_cast<uint4*>(dx), Cp, id, ih, iw);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int add16_launch(void* dst, const void* src, long long n_chunks, cudaStream_t s) {
    add16_kernel<<<ew_grid(n_chunks, 256), 256, 0, s>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), n_chunks);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
