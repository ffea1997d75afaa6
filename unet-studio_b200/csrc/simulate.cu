// simulate_modality on the GPU: the per-sample contrast synthesis the reference runs on the CPU right
// before visual_perception_augmentation (/root/reference/train.cpp:43-117 labelled-template overload, :119-180 image-only
// overload; call site train.cpp:459-462).  Random scalars (tissue LUT, the 20 polynomial terms, gamma) are drawn on the host in
// the reference's draw order; the volume work is four streaming kernels over the resident sample:
//   1. tissue = gaussian(LUT[label])  (or gaussian(t1w))        7-point star, LUT lookup fused into the loads
//   2. tissue = gaussian(tissue)
//   3. t1w = pow(sum_t w*x^a*z^b*(1-x)^c*(1-z)^d, gamma) where t1w > 0.02, else 0; min / max over the selected voxels
//   4. t1w = clamp((t1w - min) * (1/(max - min)), 0, 1) when max > min
// Algorithmic bytes per sample: 4 B * V * (1 + 1) + (1 + 1) + (2 or 3 + 1) + (1 + 1) ~ 10-11 float volumes (216 MB at 160x192x160).
// fp32 arithmetic in the reference's evaluation order (no FMA contraction) so that everything before pow() is bit-exact against
// oracle/simulate_oracle.py; the TIPL primitives' assumed semantics are listed in that file's header.
#include <cmath>
#include <cstring>
#include <random>
#include <string>

#include "common.cuh"
#include "simulate.h"

namespace u3d {
namespace {

struct IntDist {   // [TIPL] uniform_dist<int>(seed)(n) = std::uniform_int_distribution<int>(0, n-1) over std::mt19937 (libstdc++ >= 11)
    std::mt19937 g;
    explicit IntDist(uint32_t seed) : g(seed) {}
    uint32_t operator()(uint32_t n) {
        const uint32_t thr = uint32_t(-n) % n;
        for (;;) {
            const uint64_t prod = uint64_t(uint32_t(g())) * n;
            if (uint32_t(prod) >= thr) return uint32_t(prod >> 32);
        }
    }
};

struct UnitDist {   // [TIPL] uniform_dist<float>(0,1,seed)
    std::mt19937 g;
    explicit UnitDist(uint32_t seed) : g(seed) {}
    float operator()() {
        float u = float(uint32_t(g())) / 4294967296.0f;
        if (u >= 1.0f) u = std::nextafterf(1.0f, 0.0f);
        return u;
    }
};

__device__ __forceinline__ uint32_t enc_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t dec_ordered_bits(uint32_t e) { return (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e; }

struct SimLut {
    float v[kSimMaxLut];
};

// 7-point star on linear offsets; additions in the order +1, -1, +W, -W, +WH, -WH after 2*centre, then /8.
template <bool LUT>
__global__ void __launch_bounds__(256) k_sim_star(const float* __restrict__ src, float* __restrict__ dst, long long n, int W, long long WH,
                                                  const __grid_constant__ SimLut lut, int n_lut) {
    __shared__ float slut[LUT ? kSimMaxLut : 1];   // divergent label values: shared memory, not the constant bank
    if (LUT) {
        for (int k = threadIdx.x; k < n_lut; k += blockDim.x) slut[k] = lut.v[k];
        __syncthreads();
    }
    auto at = [&](long long i) -> float {
        const float v = __ldg(src + i);
        if (!LUT) return v;
        int k = int(v);
        k = k < 0 ? 0 : (k >= n_lut ? n_lut - 1 : k);
        return slut[k];
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float d = __fmul_rn(at(i), 2.0f);
        if (i >= 1) d = __fadd_rn(d, at(i - 1));
        if (i + 1 < n) d = __fadd_rn(d, at(i + 1));
        if (i >= W) d = __fadd_rn(d, at(i - W));
        if (i + W < n) d = __fadd_rn(d, at(i + W));
        if (i >= WH) d = __fadd_rn(d, at(i - WH));
        if (i + WH < n) d = __fadd_rn(d, at(i + WH));
        dst[i] = __fmul_rn(d, 0.125f);
    }
}

// Same star, four consecutive voxels per thread with 16-byte loads (W and n multiples of 4): 6 vector + 2 scalar loads per 4 voxels
// instead of 28 scalar ones.  Same additions in the same order.
template <bool LUT>
__global__ void __launch_bounds__(256) k_sim_star4(const float* __restrict__ src, float* __restrict__ dst, long long n4, int W, long long WH,
                                                   const __grid_constant__ SimLut lut, int n_lut) {
    __shared__ float slut[LUT ? kSimMaxLut : 1];
    if (LUT) {
        for (int k = threadIdx.x; k < n_lut; k += blockDim.x) slut[k] = lut.v[k];
        __syncthreads();
    }
    auto conv = [&](float v) -> float {
        if (!LUT) return v;
        int k = int(v);
        k = k < 0 ? 0 : (k >= n_lut ? n_lut - 1 : k);
        return slut[k];
    };
    auto conv4 = [&](float4 v) { return make_float4(conv(v.x), conv(v.y), conv(v.z), conv(v.w)); };
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const long long n = n4 * 4, W4 = W / 4, WH4 = WH / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += stride) {
        const long long i = q * 4;
        const float4 c = conv4(__ldg(s4 + q));
        const bool hl = i >= 1, hr = i + 4 < n, hu = q >= W4, hd = q + W4 < n4, hf = q >= WH4, hb = q + WH4 < n4;
        const float l = hl ? conv(__ldg(src + i - 1)) : 0.f, r = hr ? conv(__ldg(src + i + 4)) : 0.f;
        float4 u, d, f, b;
        if (hu) u = conv4(__ldg(s4 + q - W4));
        if (hd) d = conv4(__ldg(s4 + q + W4));
        if (hf) f = conv4(__ldg(s4 + q - WH4));
        if (hb) b = conv4(__ldg(s4 + q + WH4));
        const float cc[4] = {c.x, c.y, c.z, c.w};
        const float lf[4] = {l, c.x, c.y, c.z}, rt[4] = {c.y, c.z, c.w, r};
        const float uu[4] = {u.x, u.y, u.z, u.w}, dd[4] = {d.x, d.y, d.z, d.w}, ff[4] = {f.x, f.y, f.z, f.w}, bb[4] = {b.x, b.y, b.z, b.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = __fmul_rn(cc[j], 2.0f);
            if (j > 0 || hl) a = __fadd_rn(a, lf[j]);
            if (j < 3 || hr) a = __fadd_rn(a, rt[j]);
            if (hu) a = __fadd_rn(a, uu[j]);
            if (hd) a = __fadd_rn(a, dd[j]);
            if (hf) a = __fadd_rn(a, ff[j]);
            if (hb) a = __fadd_rn(a, bb[j]);
            o[j] = __fmul_rn(a, 0.125f);
        }
        reinterpret_cast<float4*>(dst)[q] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

struct SimTerms {
    uint8_t code[kSimTerms];   // a | b << 2 | c << 4 | d << 6
    float w[kSimTerms];
    float gamma;
};

constexpr int kSimVP = 4;   // voxels per thread and pass: the per-term dispatch below is paid once for four voxels

// One polynomial term in the reference's evaluation order (((w * x^a) * z^b) * (1-x)^c) * (1-z)^d with the exponents as template
// constants, so the power tables stay in registers (run-time exponents cost ~12 selects per term: 240 us per sample measured).
// Two 16-way dispatches (a,b) then (c,d) instead of one 256-way: the 256-case body is 86 KB of code and stalled on instruction
// fetch (ncu: no_instruction 2.6 warps per issue).
template <int AB>
__device__ __forceinline__ void sim_term_ab(float w, const float (&px)[4][kSimVP], const float (&pz)[4][kSimVP], float (&v)[kSimVP]) {
#pragma unroll
    for (int u = 0; u < kSimVP; ++u) v[u] = __fmul_rn(__fmul_rn(w, px[AB & 3][u]), pz[AB >> 2][u]);
}
template <int CD>
__device__ __forceinline__ void sim_term_cd(const float (&qx)[4][kSimVP], const float (&qz)[4][kSimVP], const float (&v)[kSimVP],
                                            float (&s)[kSimVP]) {
#pragma unroll
    for (int u = 0; u < kSimVP; ++u) s[u] = __fadd_rn(s[u], __fmul_rn(__fmul_rn(v[u], qx[CD & 3][u]), qz[CD >> 2][u]));
}

#define U3D_SIM_AB(c) case c: sim_term_ab<c>(w, px, pz, v); break;
#define U3D_SIM_CD(c) case c: sim_term_cd<c>(qx, qz, v, s); break;
#define U3D_SIM_R4(m, b) m(b) m(b + 1) m(b + 2) m(b + 3)
#define U3D_SIM_R16(m) U3D_SIM_R4(m, 0) U3D_SIM_R4(m, 4) U3D_SIM_R4(m, 8) U3D_SIM_R4(m, 12)

__global__ void __launch_bounds__(256, 5) k_sim_poly(float* __restrict__ t1w, const float* __restrict__ tissue, const float* __restrict__ label,
                                                  long long n, const __grid_constant__ SimTerms T, uint32_t* __restrict__ minmax) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n; i0 += kSimVP * stride) {
        float px[4][kSimVP], pz[4][kSimVP], qx[4][kSimVP], qz[4][kSimVP], s[kSimVP], xin[kSimVP], lab[kSimVP];
#pragma unroll
        for (int u = 0; u < kSimVP; ++u) {
            const long long i = i0 + u * stride;
            const bool ok = i < n;
            const float x = ok ? t1w[i] : 0.f, z = ok ? tissue[i] : 0.f;
            lab[u] = (ok && label) ? label[i] : 1.f;   // every load of the pass is in flight before the arithmetic starts
            const float rx = __fsub_rn(1.0f, x), rz = __fsub_rn(1.0f, z);
            xin[u] = x;
            px[0][u] = 1.f; px[1][u] = x; px[2][u] = __fmul_rn(x, x); px[3][u] = __fmul_rn(px[2][u], x);
            pz[0][u] = 1.f; pz[1][u] = z; pz[2][u] = __fmul_rn(z, z); pz[3][u] = __fmul_rn(pz[2][u], z);
            qx[0][u] = 1.f; qx[1][u] = rx; qx[2][u] = __fmul_rn(rx, rx); qx[3][u] = __fmul_rn(qx[2][u], rx);
            qz[0][u] = 1.f; qz[1][u] = rz; qz[2][u] = __fmul_rn(rz, rz); qz[3][u] = __fmul_rn(qz[2][u], rz);
            s[u] = 0.f;
        }
#pragma unroll 1
        for (int t = 0; t < kSimTerms; ++t) {
            const float w = T.w[t];
            const int code = T.code[t];   // uniform over the grid: no divergence
            float v[kSimVP];
            switch (code & 15) { U3D_SIM_R16(U3D_SIM_AB) }
            switch (code >> 4) { U3D_SIM_R16(U3D_SIM_CD) }
        }
#pragma unroll
        for (int u = 0; u < kSimVP; ++u) {
            const long long i = i0 + u * stride;
            if (i >= n) continue;
            if (xin[u] <= 0.02f) {   // train.cpp:87-92
                t1w[i] = 0.f;
                continue;
            }
            const float r = powf(s[u], T.gamma);
            t1w[i] = r;
            if (lab[u] != 0.f && r == r) {
                const uint32_t e = enc_ordered(r);
                lo = min(lo, e);
                hi = max(hi, e);
            }
        }
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    __shared__ uint32_t slo[8], shi[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        if (lo != 0xFFFFFFFFu) {   // min / max are order-independent: the atomics are deterministic
            atomicMin(minmax, lo);
            atomicMax(minmax + 1, hi);
        }
    }
}

__global__ void __launch_bounds__(256) k_sim_renorm(float* __restrict__ t1w, long long n, const uint32_t* __restrict__ minmax) {
    const uint32_t elo = minmax[0], ehi = minmax[1];
    if (elo == 0xFFFFFFFFu) return;   // nothing selected: mn = FLT_MAX, mx = -FLT_MAX -> the reference skips the rescale
    const float mn = __uint_as_float(dec_ordered_bits(elo)), mx = __uint_as_float(dec_ordered_bits(ehi));
    if (!(mx > mn)) return;
    const float inv = __fdiv_rn(1.0f, __fsub_rn(mx, mn));
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = __fmul_rn(__fsub_rn(t1w[i], mn), inv);
        v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
        t1w[i] = v;
    }
}

}  // namespace

int simulate_make_plan(int labelled, unsigned max_label, unsigned seed, int W, int H, int D, SimPlan& plan) {
    if (W <= 0 || H <= 0 || D <= 0) { set_error("simulate_modality: bad shape"); return 1; }
    if (labelled && max_label + 1 > unsigned(kSimMaxLut)) {
        set_error("simulate_modality: max_label " + std::to_string(max_label) + " exceeds the " + std::to_string(kSimMaxLut - 1) + " supported");
        return 1;
    }
    std::memset(&plan, 0, sizeof(plan));
    plan.W = W; plan.H = H; plan.D = D;
    plan.labelled = labelled ? 1 : 0;
    IntDist rand_int(seed);
    UnitDist rand_float(seed + 1u);
    if (labelled) {
        plan.n_lut = int(max_label) + 1;
        for (int i = 0; i < plan.n_lut; ++i) {
            const float r = rand_float() * 0.2f;
            plan.lut[i] = 0.4f + r;
        }
    }
    for (int t = 0; t < kSimTerms; ++t) {
        uint32_t a, b;
        do {
            a = rand_int(4);
            b = rand_int(4);
        } while (a + b == 0);
        plan.a[t] = uint8_t(a);
        plan.b[t] = uint8_t(b);
        plan.c[t] = uint8_t(rand_int(4));
        plan.d[t] = uint8_t(rand_int(4));
        plan.w[t] = rand_float();
    }
    const float g = 1.2f * rand_float();
    plan.gamma = 0.6f + g;
    return 0;
}

size_t simulate_workspace_bytes(int W, int H, int D) { return 2 * size_t(W) * H * D * sizeof(float) + 256; }

int simulate_run(const SimPlan& plan, float* t1w, const float* label, void* workspace, cudaStream_t s, long long* launches) {
    const long long n = 1LL * plan.W * plan.H * plan.D;
    if (plan.labelled && !label) { set_error("simulate_modality: the labelled overload needs a label volume"); return 1; }
    uint32_t* minmax = static_cast<uint32_t*>(workspace);
    float* ta = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 256);
    float* tb = ta + n;
    U3D_CUDA_CHECK(cudaMemsetAsync(minmax, 0xFF, 4, s));
    U3D_CUDA_CHECK(cudaMemsetAsync(minmax + 1, 0, 4, s));
    const int grid = int(std::min<long long>((n + 255) / 256, 148LL * 8));
    SimLut lut;
    std::memcpy(lut.v, plan.lut, sizeof(lut.v));
    const long long WH = 1LL * plan.W * plan.H;
    const bool aligned = ((reinterpret_cast<uintptr_t>(t1w) | reinterpret_cast<uintptr_t>(label)) & 15u) == 0;
    if (plan.W % 4 == 0 && aligned) {   // n is then a multiple of 4 as well and every row start is 16-byte aligned
        const int g4 = int(std::min<long long>((n / 4 + 255) / 256, 148LL * 8));
        if (plan.labelled)
            k_sim_star4<true><<<g4, 256, 0, s>>>(label, ta, n / 4, plan.W, WH, lut, plan.n_lut);
        else
            k_sim_star4<false><<<g4, 256, 0, s>>>(t1w, ta, n / 4, plan.W, WH, lut, 0);
        k_sim_star4<false><<<g4, 256, 0, s>>>(ta, tb, n / 4, plan.W, WH, lut, 0);
    } else {
        if (plan.labelled)
            k_sim_star<true><<<grid, 256, 0, s>>>(label, ta, n, plan.W, WH, lut, plan.n_lut);
        else
            k_sim_star<false><<<grid, 256, 0, s>>>(t1w, ta, n, plan.W, WH, lut, 0);
        k_sim_star<false><<<grid, 256, 0, s>>>(ta, tb, n, plan.W, WH, lut, 0);
    }
    SimTerms T;
    for (int t = 0; t < kSimTerms; ++t) T.code[t] = uint8_t(plan.a[t] | plan.b[t] << 2 | plan.c[t] << 4 | plan.d[t] << 6);
    std::memcpy(T.w, plan.w, sizeof(T.w));
    T.gamma = plan.gamma;
    k_sim_poly<<<grid, 256, 0, s>>>(t1w, tb, plan.labelled ? label : nullptr, n, T, minmax);
    k_sim_renorm<<<grid, 256, 0, s>>>(t1w, n, minmax);
    U3D_CUDA_CHECK(cudaGetLastError());
    if (launches) *launches += 4;
    return 0;
}

}  // namespace u3d
