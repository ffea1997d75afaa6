// Microbenchmark: tcgen05.mma.cta_group::2 (one instruction drives the tensor cores of a CTA pair: M = 256 = 128 rows per CTA, each CTA
// supplies its own A rows and half of the B columns from its own shared memory) against cta_group::1 — is the ~45 clk per-instruction
// floor of tools/mma_ts_bench.cu paid once per pair?  All-ones operands, so D = 16 * (number of accumulating MMAs) in both CTAs.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_2cta_bench tools/mma_2cta_bench.cu
#include <cstdio>
#include <cstdint>
#include "../unet-studio_b200/csrc/common.cuh"
using namespace u3d;

constexpr int kKS = 6, kAcc = 3;
constexpr uint32_t kABytes = 128 * 16 * kKS * 2;   // 24 KB: [12 chunks of 16 B][128 rows][16 B]
constexpr uint32_t kBMax = 128 * 16 * 2;           // one K step of B at N = 128

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void umma2_f16_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma2_f16_first(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}

// n = N of the pair instruction (each CTA holds n/2 columns of B)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2(int iters, int n, long long* clk, float* probe) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    const uint32_t a0 = sb, b0 = sb + kABytes;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < int(kABytes + kKS * kAcc * kBMax) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = tptr;
    const int half = n / 2;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(256, n, 0, 0, 0, 0);
        const uint32_t bstep = uint32_t(half) * 32u, dstep = uint32_t(n < 64 ? n : 64);
        uint64_t ad[kKS], bd[kKS * kAcc];
        uint32_t dd[kAcc];
#pragma unroll
        for (int ks = 0; ks < kKS; ++ks) {
            ad[ks] = umma_smem_desc(a0 + ks * 4096u, 2048u, 128u);
#pragma unroll
            for (int acc = 0; acc < kAcc; ++acc) bd[ks * kAcc + acc] = umma_smem_desc(b0 + (ks * kAcc + acc) * bstep, uint32_t(half) * 16u, 128u);
        }
#pragma unroll
        for (int acc = 0; acc < kAcc; ++acc) dd[acc] = tm + acc * dstep;
#pragma unroll
        for (int acc = 0; acc < kAcc; ++acc) umma2_f16_first(dd[acc], ad[0], bd[acc], idesc);
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < kKS; ++ks) {
#pragma unroll
                for (int acc = 0; acc < kAcc; ++acc) umma2_f16_acc(dd[acc], ad[ks], bd[ks * kAcc + acc], idesc);
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        clk[0] = t1 - t0; clk[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer must stay resident (its shared memory and TMEM are operands) until the leader has seen completion
    tc_fence_after();
    if (threadIdx.x < 32) {
        float v[16];
        tmem_ld16(tm, v);
        if (threadIdx.x == 0) { probe[2 * rank] = v[0]; probe[2 * rank + 1] = v[15]; }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    long long* dC; float* dP;
    cudaMalloc(&dC, 16); cudaMalloc(&dP, 16);
    const size_t smem = kABytes + kKS * kAcc * kBMax + 1024;
    cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    for (int n : {32, 64, 96, 128, 192, 256}) {
        const int iters = 400;
        cudaMemset(dP, 0, 16);
        bench2<<<2, 128, smem>>>(iters, n, dC, dP);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
        long long c[2]; float p[4];
        cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(p, dP, 16, cudaMemcpyDeviceToHost);
        printf("cta_group::2 M=256 N=%3d: issue %.1f clk, complete %.1f clk per pair instruction (= %.1f clk per 128-row MMA); "
               "D[0] cta0 %.0f cta1 %.0f (expected %d)\n", n, double(c[0]) / (iters * 18), double(c[1]) / (iters * 18),
               double(c[1]) / (iters * 18) / 2, p[0], p[2], 16 * (iters * kKS + 1));
    }
    return 0;
}
