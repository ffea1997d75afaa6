// Device side of SrcTransform (u3d.h): per-channel coefficients and the transform of one 16-byte chunk (8 channels of a voxel).
// The arithmetic is norm_act_fwd_kernel's (elementwise.cu), operation for operation -- fp32 scale*x + shift, activation, round to
// fp16 -- so a consumer that transforms on the fly computes on bit-identical operands to one that reads the materialised tensor.
#pragma once
#include "common.cuh"
#include "elementwise.h"
#include "u3d.h"

namespace u3d {

// activation known at compile time: branch-free, 1-2 instructions per element (with a run-time switch inside an unrolled chunk loop
// the compiler inlines the ELU path 8 times: ~370 instructions per 16-byte chunk).  Same results as act_fwd of elementwise.cu:
// max(z, 0.01 z) == (z > 0 ? z : 0.01 z) for every z incl. -0 / NaN.
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

}  // namespace u3d
