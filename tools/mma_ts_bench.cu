// Microbenchmark: tcgen05.mma (kind::f16, M = 128, N = 64, K = 16) with the A operand in TENSOR MEMORY (staged by tcgen05.cp
// 128x256b from the same SWIZZLE_NONE K-major shared-memory tile conv_band uses) against the A-from-shared-memory form.
// Pattern = conv_band's inner loop: per (plane, ky) 6 K steps x 3 accumulators = 18 MMAs that share one 128 x 96 A tile.
//   1. correctness: D(ts) == D(ss) bit for bit on small-integer operands
//   2. clk per MMA for both forms, one issuing thread, fully unrolled
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_ts_bench tools/mma_ts_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../unet-studio_b200/csrc/common.cuh"
using namespace u3d;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_f16_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc)
        : "memory");
}

constexpr int kN = 64, kKS = 6, kAcc = 3;
constexpr uint32_t kABytes = 128 * 16 * kKS * 2;      // [kchunk16B = 12][128 rows][16 B]: 24 KB per A tile
constexpr uint32_t kBBytes = kN * 16 * 2;             // one K step of B: [2 chunks][64][16 B] = 2 KB
constexpr uint32_t kBMax = 128 * 16 * 2;             // timing runs go up to N = 128
constexpr int kATiles = 4;

// mode 0: SS (A from shared memory); 1: TS (tcgen05.cp then A from TMEM); 2: TS without the copies (pure MMA rate)
__global__ void __launch_bounds__(128, 1) bench(int mode, int iters, int n, const __half* gA, const __half* gB, float* out, long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    const uint32_t a0 = sb, b0 = sb + kATiles * kABytes;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    // A tile t: element (row r, k) at t*kABytes + (k/8)*2048 + r*16 + (k%8)*2 ; B (tap j = kstep*3 + acc): element (n, k) at
    // j*kBBytes + (k/8)*(kN*16) + n*16 + (k%8)*2
    for (int i = threadIdx.x; i < kATiles * 128 * 96; i += blockDim.x) {
        const int t = i / (128 * 96), r = (i / 96) % 128, k = i % 96;
        *reinterpret_cast<__half*>(smem + t * kABytes + (k / 8) * 2048 + r * 16 + (k % 8) * 2) = gA[i];
    }
    for (int i = threadIdx.x; i < kKS * kAcc * kN * 16; i += blockDim.x) {
        const int j = i / (kN * 16), n = (i / 16) % kN, k = i % 16;
        *reinterpret_cast<__half*>(smem + kATiles * kABytes + j * kBBytes + (k / 8) * (kN * 16) + n * 16 + (k % 8) * 2) = gB[i];
    }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    const uint32_t tmA = tm + 256;   // A staging buffers: 4 x 48 columns behind the accumulators
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        const uint32_t bstep = uint32_t(n) * 32u, dstep = uint32_t(n < 64 ? n : 64);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t at = a0 + (it % kATiles) * kABytes;
            const uint32_t ta = tmA + (it % 4) * 48;
            if (mode == 1 || (mode == 2 && it < 4)) {
#pragma unroll
                for (int ks = 0; ks < kKS; ++ks) tmem_cp_128x256b(ta + ks * 8, umma_smem_desc(at + ks * 4096u, 2048u, 128u));
            }
#pragma unroll
            for (int ks = 0; ks < kKS; ++ks) {
#pragma unroll
                for (int acc = 0; acc < kAcc; ++acc) {
                    const uint64_t bd = umma_smem_desc(b0 + (ks * kAcc + acc) * bstep, uint32_t(n) * 16u, 128u);
                    const uint32_t accum = (it > 0 || ks > 0) ? 1u : 0u;
                    if (mode == 0) umma_f16(tm + acc * dstep, umma_smem_desc(at + ks * 4096u, 2048u, 128u), bd, idesc, accum);
                    else umma_f16_ts(tm + acc * dstep, ta + ks * 8, bd, idesc, accum);
                }
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        clk[0] = t1 - t0; clk[1] = t2 - t0;
    }
    __syncthreads();
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int acc = 0; acc < kAcc; ++acc)
        for (int c = 0; c < kN; c += 16) {
            float v[16];
            tmem_ld16(tm + (uint32_t(warp * 32) << 16) + acc * kN + c, v);
            for (int j = 0; j < 16; ++j) out[(size_t(acc) * 128 + warp * 32 + lane) * kN + c + j] = v[j];
        }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// Lean issue loop (what a production issuer looks like): every descriptor precomputed, constant accumulate predicate, the loop body is
// nothing but the 18 (+6) tcgen05 instructions.  mode as above.
template <int mode>
__global__ void __launch_bounds__(128, 1) bench_lean(int iters, int n, long long* clk, uint32_t a_lbo = 2048u, uint32_t a_step = 4096u, uint32_t a_off = 0u, int same_acc = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    const uint32_t a0 = sb, b0 = sb + kATiles * kABytes;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < int(kATiles * kABytes + kKS * kAcc * kBMax) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        const uint32_t bstep = n > 128 ? 0u : uint32_t(n) * 32u, dstep = uint32_t(n < 64 ? n : 64);   // N > 128: every MMA reads the same B tile
        uint64_t ad[kKS], bd[kKS * kAcc];
        uint32_t ta[2][kKS], dd[kAcc];
#pragma unroll
        for (int ks = 0; ks < kKS; ++ks) {
            ad[ks] = umma_smem_desc(a0 + a_off + ks * a_step, a_lbo, 128u);
            ta[0][ks] = tm + 256 + ks * 8;
            ta[1][ks] = tm + 256 + 48 + ks * 8;
#pragma unroll
            for (int acc = 0; acc < kAcc; ++acc) bd[ks * kAcc + acc] = umma_smem_desc(b0 + (ks * kAcc + acc) * bstep, uint32_t(n) * 16u, 128u);
        }
#pragma unroll
        for (int acc = 0; acc < kAcc; ++acc) dd[acc] = tm + (same_acc ? 0u : acc * dstep);
        if (mode == 2)
            for (int h = 0; h < 2; ++h)
                for (int ks = 0; ks < kKS; ++ks) tmem_cp_128x256b(ta[h][ks], ad[ks]);
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; it += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (mode == 1) {
#pragma unroll
                    for (int ks = 0; ks < kKS; ++ks) tmem_cp_128x256b(ta[h][ks], ad[ks]);
                }
#pragma unroll
                for (int ks = 0; ks < kKS; ++ks) {
#pragma unroll
                    for (int acc = 0; acc < kAcc; ++acc) {
                        if (mode == 0) umma_f16_acc(dd[acc], ad[ks], bd[ks * kAcc + acc], idesc);
                        else umma_f16_ts_acc(dd[acc], ta[h][ks], bd[ks * kAcc + acc], idesc);
                    }
                }
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        clk[0] = t1 - t0; clk[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// NT issuing threads (lane 0 of warps 0..NT-1), SS mode, disjoint accumulators: is the ~45 clk floor per thread or per SM?
template <int NT>
__global__ void __launch_bounds__(128, 1) bench_multi(int iters, int n, long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    const uint32_t a0 = sb, b0 = sb + kATiles * kABytes;
    if (threadIdx.x == 0) { for (int t = 0; t < 4; ++t) mbar_init(smem_u32(&bar[t]), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < int(kATiles * kABytes + kKS * kAcc * kBMax) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && warp < NT) {
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        const uint32_t bstep = uint32_t(n) * 32u;
        uint64_t ad[kKS], bd[kKS * kAcc];
        uint32_t dd[kAcc];
#pragma unroll
        for (int ks = 0; ks < kKS; ++ks) {
            ad[ks] = umma_smem_desc(a0 + warp * kABytes + ks * 4096u, 2048u, 128u);
#pragma unroll
            for (int acc = 0; acc < kAcc; ++acc) bd[ks * kAcc + acc] = umma_smem_desc(b0 + (ks * kAcc + acc) * bstep, uint32_t(n) * 16u, 128u);
        }
#pragma unroll
        for (int acc = 0; acc < kAcc; ++acc) dd[acc] = tm + uint32_t(warp) * 128u + acc * uint32_t(n < 32 ? n : 32);
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < kKS; ++ks) {
#pragma unroll
                for (int acc = 0; acc < kAcc; ++acc) umma_f16_acc(dd[acc], ad[ks], bd[ks * kAcc + acc], idesc);
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0, 0xF00);
        long long t2 = clock64();
        clk[2 * warp] = t1 - t0; clk[2 * warp + 1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    std::vector<__half> hA(kATiles * 128 * 96), hB(kKS * kAcc * kN * 16);
    srand(1);
    for (auto& v : hA) v = __float2half(float(rand() % 5 - 2));
    for (auto& v : hB) v = __float2half(float(rand() % 5 - 2));
    __half *dA, *dB; float* dO; long long* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, kAcc * 128 * kN * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = kATiles * kABytes + kKS * kAcc * kBMax + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(bench_lean<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(bench_lean<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(bench_lean<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    std::vector<float> ref(kAcc * 128 * kN), got(ref.size());
    // correctness: 4 iterations (each A tile once)
    for (int mode = 0; mode < 2; ++mode) {
        bench<<<1, 128, smem>>>(mode, 4, kN, dA, dB, dO, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(mode == 0 ? ref.data() : got.data(), dO, ref.size() * 4, cudaMemcpyDeviceToHost);
    }
    // host check of the SS result itself
    double herr = 0;
    for (int acc = 0; acc < kAcc; ++acc)
        for (int r = 0; r < 128; r += 17)
            for (int n = 0; n < kN; n += 7) {
                double s = 0;
                for (int t = 0; t < 4; ++t)
                    for (int k = 0; k < 96; ++k)
                        s += double(__half2float(hA[(t * 128 + r) * 96 + k])) *
                             double(__half2float(hB[(((k / 16) * kAcc + acc) * kN + n) * 16 + k % 16]));
                herr = fmax(herr, fabs(s - ref[(size_t(acc) * 128 + r) * kN + n]));
            }
    size_t bad = 0;
    for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
    printf("SS vs host max abs err %.3g ; TS vs SS mismatches %zu of %zu\n", herr, bad, ref.size());
    const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int n : ns)
        for (int mode = 0; mode < 3; ++mode) {
            const int iters = 400;
            if (mode == 0) bench_lean<0><<<1, 128, smem>>>(iters, n, dC);
            else if (mode == 1) bench_lean<1><<<1, 128, smem>>>(iters, n, dC);
            else bench_lean<2><<<1, 128, smem>>>(iters, n, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            long long c[2];
            cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
            printf("N=%3d mode %d (%s): issue %.1f clk/MMA, complete %.1f clk/MMA\n", n, mode,
                   mode == 0 ? "A smem" : mode == 1 ? "A tmem + 6 cp per 18 MMAs" : "A tmem, no cp", double(c[0]) / (iters * 18),
                   double(c[1]) / (iters * 18));
        }
    for (int n : {64, 128, 192, 256}) {
        bench_lean<0><<<1, 128, smem>>>(400, n, dC, 2048u, 4096u, 0u, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("same acc: %s\n", cudaGetErrorString(e)); return 1; }
        long long c[2];
        cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
        printf("N=%3d every MMA into the SAME accumulator columns: %.1f clk per MMA\n", n, double(c[1]) / (400 * 18));
    }
    // conv_band's A operand: 16-byte rows, K chunks 10432 B apart, start addresses at arbitrary 16-byte offsets
    for (int n : {48, 64, 192})
        for (int v = 0; v < 4; ++v) {
            const uint32_t lbo = v == 0 ? 2048u : 10432u, step = v <= 1 ? 4096u : 144u, off = v == 3 ? 16u : 0u;
            bench_lean<0><<<1, 128, smem>>>(400, n, dC, lbo, step, off);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("layout %d: %s\n", v, cudaGetErrorString(e)); return 1; }
            long long c[2];
            cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
            printf("N=%3d A layout: LBO %5u, start step %4u B, offset %2u B: %.1f clk per MMA\n", n, lbo, step, off, double(c[1]) / (400 * 18));
        }
    cudaFuncSetAttribute(bench_multi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(bench_multi<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(bench_multi<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    for (int n : {16, 32, 64})
        for (int nt : {1, 2, 4}) {
            const int iters = 400;
            if (nt == 1) bench_multi<1><<<1, 128, smem>>>(iters, n, dC);
            else if (nt == 2) bench_multi<2><<<1, 128, smem>>>(iters, n, dC);
            else bench_multi<4><<<1, 128, smem>>>(iters, n, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("multi %d: %s\n", nt, cudaGetErrorString(e)); return 1; }
            long long c[8];
            cudaMemcpy(c, dC, 64, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int t = 0; t < nt; ++t) mx = c[2 * t + 1] > mx ? c[2 * t + 1] : mx;
            printf("N=%3d, %d issuing threads (SS): %.1f clk per MMA per thread, %.1f clk per MMA aggregate\n", n, nt,
                   double(mx) / (iters * 18), double(mx) / (iters * 18 * nt));
        }
    return 0;
}
