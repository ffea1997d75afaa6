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
#include "xform.cuh"

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
    const long long li = L.shift == 0 ? vox : ((long long)(z << L.shift) * L.H0 + (y << L.shift)) * L.W0 + (x << L.shift);
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
        int x = 0, y = 0, z = 0;
        if (L.shift != 0) {   // level 0 reads label[vox] directly; deeper levels need (x,y,z) (32-bit: a level has < 2^31 voxels)
            const unsigned uv = unsigned(vox);
            x = int(uv % unsigned(L.w));
            const unsigned q = uv / unsigned(L.w);
            y = int(q % unsigned(L.h));
            z = int(q / unsigned(L.h));
        }
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
    // per-warp rows in shared memory, fixed-order sum over the warps, one partial row per block: no atomics, deterministic
    __shared__ float red[8][kAccN];
    a_ce = warp_sum(a_ce); a_n = warp_sum(a_n); a_mse = warp_sum(a_mse);
#pragma unroll
    for (int c = 1; c < MAXC; ++c)
        if (c < Cc) { a_i[c] = warp_sum(a_i[c]); a_k[c] = warp_sum(a_k[c]); }
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[wid][0] = a_ce; red[wid][1] = a_n; red[wid][2] = a_mse;
        red[wid][3] = 0.f; red[wid][4] = 0.f;
#pragma unroll
        for (int c = 1; c < MAXC; ++c)
            if (c < Cc) { red[wid][3 + 2 * c] = a_i[c]; red[wid][4 + 2 * c] = a_k[c]; }
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 3 + 2 * Cc; i += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += red[w][i];
        L.part[size_t(blockIdx.x) * kAccN + i] = t;
    }
}

__global__ void __launch_bounds__(1024) loss_finalize_kernel(const LossLevel L, int rows) {
    const int Cc = L.collapse_before ? L.C - L.collapse_before + 1 : L.C;
    // fixed-order double sum over the per-block partial rows.  The 1024 threads form G groups of NI lanes (NI = the accumulator count
    // rounded up to a power of two: 8 for a binary head, so G = 128): thread (g, i) walks rows g, g+G, ... of accumulator i, then a
    // fixed-shape tree over the groups.  (With a fixed 10 x 96 split only 70 of 960 threads worked: 23 us per level on the critical
    // path between the forward and the backward pass.)
    __shared__ double tot[kAccN];
    __shared__ double grp[1024];
    const int nacc = 3 + 2 * Cc;
    int NI = 8;
    while (NI < nacc) NI <<= 1;          // <= 128 (kAccN = 67)
    const int G = 1024 / NI;
    const int i = threadIdx.x % NI, g = threadIdx.x / NI;
    double t = 0;
    if (i < nacc)
        for (int r = g; r < rows; r += G) t += double(L.part[size_t(r) * kAccN + i]);
    grp[g * NI + i] = t;
    __syncthreads();
    for (int half = G >> 1; half > 0; half >>= 1) {   // G is a power of two
        if (g < half) grp[g * NI + i] += grp[(g + half) * NI + i];
        __syncthreads();
    }
    if (g == 0 && i < nacc) {
        tot[i] = grp[i];
        L.acc[i] = grp[i];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double n = tot[1] > 1.0 ? tot[1] : 1.0;
    double dice_sum = 0;
    const double eps = double(1e-5f);
    for (int c = 1; c < Cc; ++c) dice_sum += (2.0 * tot[3 + 2 * c] + eps) / (tot[4 + 2 * c] + eps);
    const double Z = double(Cc - 1 > 1 ? Cc - 1 : 1);
    L.out3[0] = float(tot[0] / n);
    L.out3[1] = float(1.0 - dice_sum / Z);
    L.out3[2] = float(tot[2] / n);
}

// dL/dlogit (x loss_scale) for the ORIGINAL output channels j < L.C of one voxel (un-collapsed), from the softmax `o`
// and the global Dice sums.  Shared by the stand-alone gradient kernel and the head-fused one.
template <int MAXC>
__device__ __forceinline__ void voxel_grad(const LossLevel& L, const Voxel<MAXC>& o, const float* sI, const float* sK, float inv_n,
                                           float invZ, float (&dlo)[MAXC]) {
    const int cb = L.collapse_before;
    const int Cc = cb ? L.C - cb + 1 : L.C;
    const float eps = 1e-5f;
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
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
        float v = 0.f;
        if (j < L.C) {
            if (cb) {
                if (j < cb) {
                    v = dl[0] * o.wi[j];
                } else {
#pragma unroll
                    for (int c = 1; c < MAXC; ++c)
                        if (c == j - cb + 1) v = dl[c];
                }
            } else {
                v = dl[j];
            }
        }
        dlo[j] = v;
    }
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
    const float invZ = 1.f / float(Cc - 1 > 1 ? Cc - 1 : 1);
    __half* out = static_cast<__half*>(L.dlogits);
    for (long long vox = blockIdx.x * (long long)blockDim.x + threadIdx.x; vox < nv; vox += (long long)gridDim.x * blockDim.x) {
        int x = 0, y = 0, z = 0;
        if (L.shift != 0) {   // level 0 reads label[vox] directly; deeper levels need (x,y,z) (32-bit: a level has < 2^31 voxels)
            const unsigned uv = unsigned(vox);
            x = int(uv % unsigned(L.w));
            const unsigned q = uv / unsigned(L.w);
            y = int(q % unsigned(L.h));
            z = int(q / unsigned(L.h));
        }
        Voxel<MAXC> o;
        eval_voxel<MAXC>(L, vox, x, y, z, o);
        float dlo[MAXC];
        voxel_grad<MAXC>(L, o, sI, sK, inv_n, invZ, dlo);
        // 8 channels = one 16-byte store (adjacent threads = adjacent voxels)
        uint4* row = reinterpret_cast<uint4*>(out + vox * L.Cp);
        for (int j0 = 0; j0 < L.Cp; j0 += 8) {
            float val[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float v = 0.f;
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                    if (c == j0 + jj) v = dlo[c];
                val[jj] = v;
            }
            uint4 qv;
            qv.x = pack2<false>(val[0], val[1]); qv.y = pack2<false>(val[2], val[3]);
            qv.z = pack2<false>(val[4], val[5]); qv.w = pack2<false>(val[6], val[7]);
            row[j0 / 8] = qv;
        }
    }
}

// ---- 1x1 output head fused with the loss gradient (levels whose head input has <= 32 channels: bandwidth-bound) ----
// forward: logits[c][v] = b[c] + sum_k W[c][k] x[v][k]            (unet.cpp:186-187, Conv3d k1 of the output token)
// xf.enabled: x is the RAW output of the last conv; its norm + activation is applied here (and the activated voxels stored to
// xf.writeback for the backward pass) -- see SrcTransform in u3d.h.
// One thread per (voxel, group of 8 input channels): a warp's loads (and write-back stores) are contiguous 16-byte chunks; the
// XCP/8 partial dot products of a voxel are combined by shuffles in a fixed order and lane g writes the classes c = g, g + GP, ...
// ACT = -1: plain input; otherwise the activation of the folded norm step as a compile-time constant
template <int XCP, int ACT>
__global__ void head_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ logits, int C, int xc, long long nv, const SrcTransform xf) {
    constexpr int GP = XCP / 8;
    __shared__ float sw[kMaxC * XCP];
    __shared__ float sb[kMaxC];
    __shared__ float ssc[XCP], ssh[XCP];
    for (int i = threadIdx.x; i < C * XCP; i += blockDim.x) {
        const int c = i / XCP, k = i % XCP;
        sw[i] = k < xc ? w[c * xc + k] : 0.f;
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = b[i];
    if (ACT >= 0)
        for (int i = threadIdx.x; i < XCP; i += blockDim.x) xf_coef1(xf, i, ssc[i], ssh[i]);
    uint4* const wb = ACT >= 0 ? static_cast<uint4*>(xf.writeback) : nullptr;
    __syncthreads();
    const int g = int(threadIdx.x) % GP;
    const long long total = nv * GP;
    // whole warps stay in the loop (the shuffles need all GP lanes of a voxel; GP divides 32 and the stride); two chunks per thread
    // and iteration so that two loads are in flight
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = blockIdx.x * (long long)blockDim.x; base < total; base += 2 * stride) {
        const long long i0 = base + threadIdx.x, i1 = i0 + stride;
        const bool on0 = i0 < total, on1 = i1 < total;
        uint4 q[2];
        q[0] = on0 ? x[i0] : make_uint4(0u, 0u, 0u, 0u);
        q[1] = on1 ? x[i1] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long i = h ? i1 : i0;
            const bool on = h ? on1 : on0;
            if (ACT >= 0) {
                q[h] = xf_apply_c<(ACT < 0 ? 0 : ACT)>(q[h], ssc + g * 8, ssh + g * 8);
                if (wb != nullptr && on) wb[i] = q[h];
            }
            const uint32_t u[4] = {q[h].x, q[h].y, q[h].z, q[h].w};
            float xv[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2<false>(u[j]);
                xv[2 * j] = f.x;
                xv[2 * j + 1] = f.y;
            }
            const long long vox = i / GP;
            for (int c = 0; c < C; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = fmaf(sw[c * XCP + g * 8 + k], xv[k], acc);
#pragma unroll
                for (int o = 1; o < GP; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (on && (c % GP) == g) logits[c * nv + vox] = acc + sb[c];
            }
        }
    }
}

// backward: dlogits stay in registers; dx[v][k] = sum_c dl[c] W[c][k] (fp16, store or accumulate), dW[c][k] += dl[c] x[v][k],
// db[c] += dl[c].  Replaces loss_grad + the head's dgrad / wgrad / bias-gradient launches and the dlogits round trip.
template <int CT, int XCP>
__global__ void __launch_bounds__(128, (CT * XCP <= 32 ? 5 : CT * XCP <= 64 ? 4 : 3)) loss_grad_head_kernel(const LossLevel L, const HeadFuse Hd) {
    const long long nv = (long long)L.d * L.h * L.w;
    const int cb = L.collapse_before;
    const int Cc = cb ? L.C - cb + 1 : L.C;
    __shared__ float sI[CT], sK[CT];
    __shared__ float s_n;
    __shared__ float sw[CT * XCP];
    __shared__ float sacc[CT * XCP + CT];
    for (int i = threadIdx.x; i < CT * XCP; i += blockDim.x) {
        const int c = i / XCP, k = i % XCP;
        sw[i] = (c < L.C && k < Hd.xc) ? Hd.w[c * Hd.xc + k] : 0.f;
    }
    for (int i = threadIdx.x; i < CT * XCP + CT; i += blockDim.x) sacc[i] = 0.f;
    if (threadIdx.x < CT) {
        sI[threadIdx.x] = threadIdx.x < Cc ? float(L.acc[3 + 2 * threadIdx.x]) : 0.f;
        sK[threadIdx.x] = threadIdx.x < Cc ? float(L.acc[4 + 2 * threadIdx.x]) : 0.f;
    }
    if (threadIdx.x == 0) s_n = float(L.acc[1] > 1.0 ? L.acc[1] : 1.0);
    __syncthreads();
    const float inv_n = 1.f / s_n;
    const float invZ = 1.f / float(Cc - 1 > 1 ? Cc - 1 : 1);
    const uint4* xin = static_cast<const uint4*>(Hd.x);
    uint4* dxo = static_cast<uint4*>(Hd.dx);
    // A warp takes 32 voxels per iteration.  Phase A: lane l evaluates the softmax gradient of voxel l (once per voxel; logits and
    // labels are read as contiguous 128-byte rows).  Phase B: the 32*GP 16-byte chunks of those voxels (GP = XCP/8 channel groups)
    // are processed in GP rounds of 32 contiguous chunks; a lane gets the gradient of its round's voxel by shuffle and owns 16 bytes
    // of x / dx, so it carries CT*8 weight-gradient accumulators.  (One thread per voxel needed 167 registers: 19 % occupancy; one
    // thread per chunk with the gradient evaluated redundantly by the GP lanes of a voxel was instruction-bound at 176 us.)
    constexpr int GP = XCP / 8;
    const int lane = int(threadIdx.x) & 31;
    const int g = lane % GP;
    float aw[CT][8], ab[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        ab[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) aw[c][k] = 0.f;
    }
    const long long wstride = (long long)gridDim.x * (blockDim.x / 32) * 32;
    for (long long v0 = (blockIdx.x * (long long)(blockDim.x / 32) + threadIdx.x / 32) * 32; v0 < nv; v0 += wstride) {
        // the x rows of both rounds first: their latency overlaps phase A
        uint4 qx[GP], qd[GP];
#pragma unroll
        for (int r = 0; r < GP; ++r) {
            const long long vox = v0 + r * (32 / GP) + lane / GP;
            qx[r] = qd[r] = make_uint4(0u, 0u, 0u, 0u);
            if (vox < nv) {
                qx[r] = xin[vox * GP + g];
                if (Hd.dx_accum) qd[r] = dxo[vox * GP + g];
            }
        }
        float dlo[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) dlo[c] = 0.f;
        const long long va = v0 + lane;
        if (va < nv) {
            int x = 0, y = 0, z = 0;
            if (L.shift != 0) {   // level 0 reads label[vox] directly; deeper levels need (x,y,z) (32-bit: a level has < 2^31 voxels)
                const unsigned uv = unsigned(va);
                x = int(uv % unsigned(L.w));
                const unsigned q = uv / unsigned(L.w);
                y = int(q % unsigned(L.h));
                z = int(q / unsigned(L.h));
            }
            Voxel<CT> o;
            eval_voxel<CT>(L, va, x, y, z, o);
            voxel_grad<CT>(L, o, sI, sK, inv_n, invZ, dlo);
#pragma unroll
            for (int c = 0; c < CT; ++c) ab[c] += dlo[c];
        }
#pragma unroll
        for (int r = 0; r < GP; ++r) {
            const int vl = r * (32 / GP) + lane / GP;
            const long long vox = v0 + vl;
            float d[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) d[c] = __shfl_sync(0xffffffffu, dlo[c], vl);
            const uint32_t u[4] = {qx[r].x, qx[r].y, qx[r].z, qx[r].w};
            const uint32_t ud[4] = {qd[r].x, qd[r].y, qd[r].z, qd[r].w};
            float xv[8], dv[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2<false>(u[j]);
                xv[2 * j] = f.x;
                xv[2 * j + 1] = f.y;
                const float2 fd = unpack2<false>(ud[j]);
                dv[2 * j] = fd.x;
                dv[2 * j + 1] = fd.y;
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    dv[k] = fmaf(d[c], sw[c * XCP + g * 8 + k], dv[k]);
                    aw[c][k] = fmaf(d[c], xv[k], aw[c][k]);      // x = 0 beyond the volume: no contribution
                }
            }
            if (vox < nv) {
                uint4 qo;
                qo.x = pack2<false>(dv[0], dv[1]); qo.y = pack2<false>(dv[2], dv[3]);
                qo.z = pack2<false>(dv[4], dv[5]); qo.w = pack2<false>(dv[6], dv[7]);
                dxo[vox * GP + g] = qo;
            }
        }
    }
    // lanes with the same channel group: lane % GP == g  ->  butterfly over the lane bits above log2(GP)
#pragma unroll
    for (int c = 0; c < CT; ++c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = aw[c][k];
#pragma unroll
            for (int o = 16; o >= GP; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) < GP) atomicAdd(&sacc[c * XCP + g * 8 + k], v);
        }
        const float vb = warp_sum(ab[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[CT * XCP + c], vb);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CT * XCP; i += blockDim.x) {
        const int c = i / XCP, k = i % XCP;
        if (c < L.C && k < Hd.xc) atomicAdd(Hd.dw + c * Hd.xc + k, sacc[i]);
    }
    for (int i = threadIdx.x; i < CT; i += blockDim.x)
        if (i < L.C) atomicAdd(Hd.db + i, sacc[CT * XCP + i]);
}

}  // namespace

int loss_part_rows() { return 148 * 8; }
int loss_part_cols() { return kAccN; }

bool head_fwd_supported(int C, int xcp) { return C >= 1 && C <= kMaxC && (xcp == 16 || xcp == 32); }
bool head_bwd_supported(int C, int xcp) {
    if (xcp != 16 && xcp != 32) return false;
    const int ct = C <= 2 ? 2 : C <= 4 ? 4 : 8;
    return C <= 8 && ct * xcp <= 128;
}

int head_fwd_launch(const void* x, int xc, int xcp, const float* w, const float* b, float* logits, int C, long long nv, cudaStream_t s,
                    const SrcTransform* xfp) {
    SrcTransform xf{};
    if (xfp) xf = *xfp;
    if (!head_fwd_supported(C, xcp)) { set_error("head_fwd_launch: unsupported shape"); return 1; }
    long long g = (nv * (xcp / 8) + 255) / 256;
    const int grid = int(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    const uint4* xp = static_cast<const uint4*>(x);
    const int act = xf.enabled ? xf.act : -1;
#define U3D_HEAD_FWD(XCP_, ACT_) head_fwd_kernel<XCP_, ACT_><<<grid, 256, 0, s>>>(xp, w, b, logits, C, xc, nv, xf)
    if (xcp == 16) {
        switch (act) {
            case ACT_NONE: U3D_HEAD_FWD(16, ACT_NONE); break;
            case ACT_RELU: U3D_HEAD_FWD(16, ACT_RELU); break;
            case ACT_LEAKY: U3D_HEAD_FWD(16, ACT_LEAKY); break;
            case ACT_ELU: U3D_HEAD_FWD(16, ACT_ELU); break;
            default: U3D_HEAD_FWD(16, -1); break;
        }
    } else {
        switch (act) {
            case ACT_NONE: U3D_HEAD_FWD(32, ACT_NONE); break;
            case ACT_RELU: U3D_HEAD_FWD(32, ACT_RELU); break;
            case ACT_LEAKY: U3D_HEAD_FWD(32, ACT_LEAKY); break;
            case ACT_ELU: U3D_HEAD_FWD(32, ACT_ELU); break;
            default: U3D_HEAD_FWD(32, -1); break;
        }
    }
#undef U3D_HEAD_FWD
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int loss_level_launch(const LossLevel& L, cudaStream_t s) { return loss_level_launch(L, nullptr, s); }

int loss_level_launch(const LossLevel& L, const HeadFuse* Hd, cudaStream_t s) {
    if (L.C > kMaxC || L.C < 1) {
        set_error("loss head supports 1..32 output channels");
        return 1;
    }
    if (L.collapse_before < 0 || L.collapse_before >= L.C) {
        set_error("invalid collapse_before");  // train.cpp:507-508
        return 1;
    }
    if (L.part == nullptr) { set_error("loss_level_launch: partial-sum scratch missing"); return 1; }
    const long long nv = (long long)L.d * L.h * L.w;
    long long g = (nv + 255) / 256;
    const int grid = int(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    // the class loops are unrolled to the template bound: the binary / few-class heads must not pay for 8
    if (L.C <= 2) loss_reduce_kernel<2><<<grid, 256, 0, s>>>(L);
    else if (L.C <= 4) loss_reduce_kernel<4><<<grid, 256, 0, s>>>(L);
    else if (L.C <= 8) loss_reduce_kernel<8><<<grid, 256, 0, s>>>(L);
    else loss_reduce_kernel<kMaxC><<<grid, 256, 0, s>>>(L);
    loss_finalize_kernel<<<1, 1024, 0, s>>>(L, grid);
    if (Hd != nullptr) {
        if (!head_bwd_supported(L.C, Hd->xcp)) { set_error("loss_level_launch: unsupported fused head shape"); return 1; }
        long long gh = (nv + 127) / 128;   // a warp takes 32 voxels per iteration
        const int gridh = int(gh < 1 ? 1 : (gh > 148 * 16 ? 148 * 16 : gh));
        const int ct = L.C <= 2 ? 2 : L.C <= 4 ? 4 : 8;
        if (ct == 2 && Hd->xcp == 16) loss_grad_head_kernel<2, 16><<<gridh, 128, 0, s>>>(L, *Hd);
        else if (ct == 2) loss_grad_head_kernel<2, 32><<<gridh, 128, 0, s>>>(L, *Hd);
        else if (ct == 4 && Hd->xcp == 16) loss_grad_head_kernel<4, 16><<<gridh, 128, 0, s>>>(L, *Hd);
        else if (ct == 4) loss_grad_head_kernel<4, 32><<<gridh, 128, 0, s>>>(L, *Hd);
        else loss_grad_head_kernel<8, 16><<<gridh, 128, 0, s>>>(L, *Hd);
    } else if (L.dlogits != nullptr) {
        if (L.C <= 8) loss_grad_kernel<8><<<grid, 256, 0, s>>>(L);
        else loss_grad_kernel<kMaxC><<<grid, 256, 0, s>>>(L);
    }
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
