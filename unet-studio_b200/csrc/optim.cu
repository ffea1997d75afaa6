// Optimizer step on the flat fp32 parameter / gradient / momentum buffers (reference tensorN order):
//   grad /= batch_size; clip_grad_norm_(12.0); SGD(momentum 0.99, Nesterov, dampening 0,
//   weight_decay 3e-5 for conv weights, 0 for biases and norm affine); zero_grad
// (/root/reference/train.cpp:759-766, unet.cpp:246-277).  The gradients arrive multiplied by the
// loss scale of the fp16 backward path; a non-finite gradient (fp16 overflow) skips the update.
#include <string>

#include "common.cuh"
#include "elementwise.h"

namespace u3d {
namespace {

__global__ void grad_sumsq_kernel(const float* __restrict__ g, long long n, float inv, SgdStatus* st) {
    double acc = 0;
    int bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g[i] * inv;
        if (!isfinite(v)) bad = 1;
        acc += double(v) * double(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ double sacc[32];
    __shared__ int sbad[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sacc[w] = acc; sbad[w] = bad; }
    __syncthreads();
    if (w == 0) {
        acc = l < (blockDim.x >> 5) ? sacc[l] : 0.0;
        bad = l < (blockDim.x >> 5) ? sbad[l] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (l == 0) {
            atomicAdd(&st->sumsq, acc);
            if (bad) atomicOr(&st->nonfinite, 1);
        }
    }
}

__global__ void sgd_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, const SgdChunk* __restrict__ chunks,
                           float inv, float lr, float mu, float max_norm, int first_step, const SgdStatus* st) {
    const SgdChunk ck = chunks[blockIdx.x];
    const bool skip = st->nonfinite != 0 || !isfinite(st->sumsq);
    const float total_norm = float(sqrt(st->sumsq));
    const float clip = fminf(max_norm / (total_norm + 1e-6f), 1.0f);   // torch clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    const float scale = inv * clip;
    for (int i = threadIdx.x; i < ck.count; i += blockDim.x) {
        const long long j = ck.offset + i;
        if (!skip) {
            const float w = p[j];
            float gr = g[j] * scale;
            if (ck.weight_decay != 0.f) gr += ck.weight_decay * w;
            const float buf = first_step ? gr : mu * m[j] + gr;
            m[j] = buf;
            gr += mu * buf;          // Nesterov
            p[j] = w - lr * gr;
        }
        g[j] = 0.f;                  // optimizer->zero_grad()
    }
}

}  // namespace

int sgd_step_launch(float* params, float* grads, float* momentum, long long n, const SgdChunk* chunks_dev, int nchunks,
                    float inv_scale_batch, float lr, float mu, float max_norm, int first_step, SgdStatus* status_dev,
                    cudaStream_t s) {
    U3D_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, sizeof(SgdStatus), s));
    grad_sumsq_kernel<<<148 * 4, 256, 0, s>>>(grads, n, inv_scale_batch, status_dev);
    sgd_kernel<<<nchunks, 256, 0, s>>>(params, grads, momentum, chunks_dev, inv_scale_batch, lr, mu, max_norm, first_step, status_dev);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
