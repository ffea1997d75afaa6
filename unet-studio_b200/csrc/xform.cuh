// Device side of SrcTransform (u3d.h): per-thread coefficients of one 8-channel group and the transform of one 16-byte chunk.
// The arithmetic is norm_act_fwd_kernel's (elementwise.cu), operation for operation -- fp32 scale*x + shift, activation, round to
// fp16 -- so a consumer that transforms on the fly reads bit-identical operands to one that reads the materialised tensor.
#pragma once
#include "common.cuh"
#include "elementwise.h"
#include "u3d.h"

namespace u3d {

template <int A>
struct ActC { static constexpr int value = A; };

struct XfCoef {
    float sc[8], sh[8];
    int act;
};

// coefficients of channels [8*cgl, 8*cgl + 8) of a source
__device__ __forceinline__ void xf_coefs(const SrcTransform& x, int cgl, XfCoef& k) {
    k.act = x.act;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cgl * 8 + j;
        float s = 1.f, b = 0.f;
        if (c >= x.C) s = 0.f;
        else if (x.has_norm) {
            const float r = x.rstd ? x.rstd[c] : 1.f;
            const float m = x.mean ? x.mean[c] : 0.f;
            s = x.gamma[c] * r;
            b = x.beta[c] - m * s;
        }
        k.sc[j] = s;
        k.sh[j] = b;
    }
}

__device__ __forceinline__ float xf_act(float z, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(z, 0.f);
        case ACT_LEAKY: return z > 0.f ? z : 0.01f * z;
        case ACT_ELU: return z > 0.f ? z : expm1f(z);
        default: return z;
    }
}

// activation known at compile time: branch-free, 1-2 instructions per element.  (The transform runs in the producer warps of a
// tensor-core kernel, one plane per ~5000 clk: with the run-time switch above it was ~370 instructions per chunk with the ELU code
// inline and made the layer 2.5x slower.)  Same results as xf_act: max(z, 0.01 z) == (z > 0 ? z : 0.01 z) for every z incl. -0 / NaN.
template <int ACT>
__device__ __forceinline__ float xf_act_c(float z) {
    if constexpr (ACT == ACT_RELU) return fmaxf(z, 0.f);
    else if constexpr (ACT == ACT_LEAKY) return fmaxf(z, 0.01f * z);
    else if constexpr (ACT == ACT_ELU) return z > 0.f ? z : expm1f(z);
    else return z;
}
template <int ACT>
__device__ __forceinline__ uint4 xf_apply_c(const uint4& q, const float* sc, const float* sh) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = unpack2<false>(w[j]);
        o[j] = pack2<false>(xf_act_c<ACT>(sc[2 * j] * t.x + sh[2 * j]), xf_act_c<ACT>(sc[2 * j + 1] * t.y + sh[2 * j + 1]));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
// calls f(std::integral_constant<int, act>) with the run-time activation as a compile-time constant
template <class F>
__device__ __forceinline__ void xf_dispatch_act(int act, F&& f) {
    switch (act) {
        case ACT_RELU: f(ActC<ACT_RELU>{}); break;
        case ACT_LEAKY: f(ActC<ACT_LEAKY>{}); break;
        case ACT_ELU: f(ActC<ACT_ELU>{}); break;
        default: f(ActC<ACT_NONE>{}); break;
    }
}

// sc / sh: the 8 coefficients of the chunk's channel group (registers or shared memory)
__device__ __forceinline__ uint4 xf_apply(const uint4& q, const float* sc, const float* sh, int act) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = unpack2<false>(w[j]);
        o[j] = pack2<false>(xf_act(sc[2 * j] * t.x + sh[2 * j], act), xf_act(sc[2 * j + 1] * t.y + sh[2 * j + 1], act));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint4 xf_apply(const uint4& q, const XfCoef& k) { return xf_apply(q, k.sc, k.sh, k.act); }

// one coefficient of channel c (the prologue of a kernel that keeps all channels in shared memory)
__device__ __forceinline__ void xf_coef1(const SrcTransform& x, int c, float& s, float& b) {
    s = 1.f; b = 0.f;
    if (c >= x.C) s = 0.f;
    else if (x.has_norm) {
        const float r = x.rstd ? x.rstd[c] : 1.f;
        const float m = x.mean ? x.mean[c] : 0.f;
        s = x.gamma[c] * r;
        b = x.beta[c] - m * s;
    }
}

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 lds16(uint32_t src) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src) : "memory");
    return v;
}
// predicated forms (no branch around the access: the four chunks of a batch stay one straight-line block)
__device__ __forceinline__ uint4 lds16_if(uint32_t src, bool on) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"(src), "r"(uint32_t(on)));
    return v;
}
__device__ __forceinline__ void sts16_if(uint32_t dst, const uint4& v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(uint32_t(on)) : "memory");
}
__device__ __forceinline__ void stg16_if(void* dst, const uint4& v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(uint32_t(on)) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t dst, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace u3d
