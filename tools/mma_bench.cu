// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) issued back to back by one thread, as a function of N,
// of the shared-memory layout (SWIZZLE_NONE interleaved vs SWIZZLE_128B) and of the A-operand strides.
#include <cstdio>
#include <cstdint>
#include "../unet-studio_b200/csrc/common.cuh"
using namespace u3d;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t(1) << 16;                 // LBO (ignored for swizzled K-major)
    d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;                 // SWIZZLE_128B
    return d;
}

// mode 0: SWIZZLE_NONE, A LBO=2048 SBO=128 (gather kernel); 1: SWIZZLE_NONE, A LBO=big (halo kernel); 2: SWIZZLE_128B
__global__ void __launch_bounds__(128, 1) bench(int n, int mode, int iters, int nacc, long long* out, int a_lbo_big) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        const uint32_t a0 = sb, b0 = sb + 128 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint64_t ad, bd;
            const uint32_t shift = (i % 27) * 16u;   // different start address per "tap" like the halo kernel
            if (mode == 0) { ad = umma_smem_desc(a0 + (i % 8) * 4096u, 2048u, 128u); bd = umma_smem_desc(b0, n * 16u, 128u); }
            else if (mode == 1) { ad = umma_smem_desc(a0 + shift, a_lbo_big, 128u); bd = umma_smem_desc(b0, n * 16u, 128u); }
            else { ad = desc_sw128(a0 + (i % 4) * 32u, 1024u); bd = desc_sw128(b0, 1024u); }
            umma_f16(tm + (i % nacc) * n, ad, bd, idesc, i >= nacc ? 1u : 0u);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// lean issue loop: 27 'taps' fully unrolled, descriptors = base + register offset
__global__ void __launch_bounds__(128, 1) bench_lean(int n, int iters, long long* out, int lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        int toff[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) toff[k] = ((k / 9) * 10 + (k / 3) % 3) * 34 + k % 3;
        const uint64_t a_base = umma_smem_desc(sb + 4096, lbo, 128u);
        const uint64_t b_base = umma_smem_desc(sb + 128 * 1024, n * 16u, 128u);
        const uint32_t bstep = (n * 32u) >> 4;
        long long t0 = clock64();
        for (int i = 0; i < iters / 27; ++i) {
            const uint64_t a_mt = a_base + uint64_t(i & 7) * 128u;
            uint64_t bd = b_base;
            umma_f16_first(tm + (i & 1) * n, a_mt + toff[0], bd, idesc);
#pragma unroll
            for (int k = 1; k < 27; ++k) {
                bd += bstep;
                umma_f16_acc(tm + (i & 1) * n, a_mt + toff[k], bd, idesc);
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// warp-uniform issue loop, MMA guarded by elect.sync (CUTLASS idiom)
__global__ void __launch_bounds__(128, 1) bench_elect(int n, int iters, long long* out, int lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x < 32) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        int toff[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) toff[k] = ((k / 9) * 10 + (k / 3) % 3) * 34 + k % 3;
        const uint64_t a_base = umma_smem_desc(sb + 4096, lbo, 128u);
        const uint64_t b_base = umma_smem_desc(sb + 128 * 1024, n * 16u, 128u);
        const uint32_t bstep = (n * 32u) >> 4;
        long long t0 = clock64();
        for (int i = 0; i < iters / 27; ++i) {
            const uint64_t a_mt = a_base + uint64_t(i & 7) * 128u;
            uint64_t bd = b_base;
            if (elect_one()) umma_f16_first(tm + (i & 1) * n, a_mt + toff[0], bd, idesc);
#pragma unroll
            for (int k = 1; k < 27; ++k) {
                bd += bstep;
                if (elect_one()) umma_f16_acc(tm + (i & 1) * n, a_mt + toff[k], bd, idesc);
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// several warps issue MMAs concurrently into different TMEM accumulators
__global__ void __launch_bounds__(256, 1) bench_multi(int n, int iters, int nwarps, long long* out, int lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    long long t0 = clock64();
    if (warp < nwarps && lane == 0) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 0, 0);
        int toff[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) toff[k] = ((k / 9) * 10 + (k / 3) % 3) * 34 + k % 3;
        const uint64_t a_base = umma_smem_desc(sb + 4096, lbo, 128u);
        const uint64_t b_base = umma_smem_desc(sb + 128 * 1024, n * 16u, 128u);
        const uint32_t bstep = (n * 32u) >> 4;
        const uint32_t d = tm + warp * 2 * n;
        for (int i = 0; i < iters / 27; ++i) {
            const uint64_t a_mt = a_base + uint64_t((i + warp) & 7) * 128u;
            uint64_t bd = b_base;
            umma_f16_first(d + (i & 1) * n, a_mt + toff[0], bd, idesc);
#pragma unroll
            for (int k = 1; k < 27; ++k) {
                bd += bstep;
                umma_f16_acc(d + (i & 1) * n, a_mt + toff[k], bd, idesc);
            }
        }
        umma_commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0, 0xF00);
    }
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t2 - t0; out[1] = t2 - t0; }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// MN-major operands with the strides of conv_wgrad_rows.cu (LBO 128 B, SBO 544 B), 9 accumulators, runtime accumulate flag
__global__ void __launch_bounds__(128, 1) bench_mn(int n, int iters, long long* out, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t idesc = (variant >= 2) ? umma_idesc(128, n, 0, 0, 0, 0) : umma_idesc(128, n, 0, 0, 1, 1);
        long long aoff[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) aoff[k] = (long long)((k / 3 - 1) * 10 * 2) * 34 + (k % 3 - 1);
        const uint64_t a_base = umma_smem_desc(sb + 32768, 128u, 544u);
        const uint64_t b_base = umma_smem_desc(sb + 150 * 1024, 128u, 512u);
        bool first = true;
        long long t0 = clock64();
        for (int i = 0; i < iters / 18; ++i) {
            const uint64_t a_row = a_base + uint64_t(i & 15) * 68u;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint64_t ad = a_row + uint64_t(aoff[k]) + uint64_t(ks * 16);
                    const uint64_t bd = b_base + uint64_t(ks * 16);
                    if (variant == 0) { if (first) umma_f16_first(tm + k * n, ad, bd, idesc); else umma_f16_acc(tm + k * n, ad, bd, idesc); }
                    else if (variant == 3) umma_f16_acc(tm + k * n, ad, bd, idesc);
                    else umma_f16(tm + k * n, ad, bd, idesc, first ? 0u : 1u);
                }
                first = false;
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// MN-major operands with 128-byte aligned chunk strides (wgrad_band layout): A chunks sbo_a apart, B chunks sbo_b apart, LBO = 128
__global__ void __launch_bounds__(128, 1) bench_mn2(int n, int iters, long long* out, int sbo_a, int sbo_b, int nacc, int lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(128, n, 0, 0, 1, 1);
        const uint64_t a_base = umma_smem_desc(sb, lbo, sbo_a);
        const uint64_t b_base = umma_smem_desc(sb + 100 * 1024, lbo, sbo_b);
        long long t0 = clock64();
        for (int i = 0; i < iters / 6; ++i) {
            const uint64_t a_row = a_base + uint64_t(i & 7) * uint64_t(2 * 128 >> 4);
            const uint64_t b_row = b_base + uint64_t(i & 3) * uint64_t(6 * 128 >> 4);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const uint64_t ad = a_row + uint64_t((k / 3) * (2 * lbo >> 4));
                const uint64_t bd = b_row + uint64_t((k / 3) * (2 * lbo >> 4));
                if (i == 0) umma_f16_first(tm + (k % nacc) * n, ad, bd, idesc); else umma_f16_acc(tm + (k % nacc) * n, ad, bd, idesc);
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// conv_band issue pattern: 6 MMAs per (dz,dy) with N = 16,32,48,48,32,16 at D column offsets 0,0,0,16,32,48 (mode 0), or the
// same A/B/D addresses with one uniform N (mode 1: N = 64 at offset 0; mode 2: N = 48 at offsets 0/16)
__global__ void __launch_bounds__(128, 1) bench_mixed(int iters, long long* out, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        fence_proxy_async();
        uint32_t idesc[6], doff[6], boff[6], aoff[6];
        const int nn[6] = {16, 32, 48, 48, 32, 16};
        const int dd[6] = {0, 0, 0, 16, 32, 48};
        const int bb[6] = {32, 16, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            idesc[k] = umma_idesc(128, mode == 0 ? nn[k] : (mode == 1 ? 64 : 48), 0, 0, 0, 0);
            doff[k] = mode == 0 ? dd[k] : (mode == 1 ? 0 : (k & 1) * 16);
            boff[k] = mode == 0 ? bb[k] : 0;
            aoff[k] = (k % 4) * 152 + k / 4;
        }
        const uint64_t a_base = umma_smem_desc(sb + 4096, 9728u, 128u);
        const uint64_t b_base = umma_smem_desc(sb + 128 * 1024, 1024u, 128u);
        long long t0 = clock64();
        for (int i = 0; i < iters / 54; ++i) {
            const uint32_t d = tm + (i & 1) * 64;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
                const uint64_t ar = a_base + uint64_t((t9 % 3) * 9 + (t9 / 3) * 1216);
                const uint64_t br = b_base + uint64_t(t9 * 128);
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if (t9 == 0 && k == 0) umma_f16_first(d + doff[k], ar + aoff[k], br + boff[k], idesc[k]);
                    else umma_f16_acc(d + doff[k], ar + aoff[k], br + boff[k], idesc[k]);
                }
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    cudaFuncSetAttribute(bench_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 1, 2}) {
        bench_mixed<<<1, 128, 200 * 1024>>>(54 * 100, d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mixed mode %d : issue %.1f cyc/mma, complete %.1f (%s)\n", mode, double(h[0]) / 5400, double(h[1]) / 5400, cudaGetErrorString(e));
    }
    cudaFuncSetAttribute(bench_mn2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {16, 48, 64, 96, 128, 192})
        for (int sa : {128, 512})
            for (int sbo_b : {128, 512}) {
                const int lbo = (sa == 128 || sbo_b == 128) ? 4096 : 128;
                const int nacc = 3 * n <= 512 ? 3 : 1;
                bench_mn2<<<1, 128, 200 * 1024>>>(n, 6 * 400, d, sa == 128 ? 128 : (lbo == 4096 ? 256 : sa), sbo_b == 128 ? 128 : (lbo == 4096 ? 256 : sbo_b), nacc, lbo);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("mn2 N %3d sbo_a %4d sbo_b %4d : issue %.1f cyc/mma, complete %.1f (%s)\n", n, sa, sbo_b, double(h[0]) / (6 * 400), double(h[1]) / (6 * 400), cudaGetErrorString(e));
            }
    for (int mode = 0; mode < 0; ++mode)
        for (int n : {16, 32, 64, 128, 256})
            for (int nacc : {1, 2}) {
                if (nacc * n > 512) continue;
                bench<<<1, 128, 200 * 1024>>>(n, mode, iters, nacc, d, 35200);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("mode %d N %3d nacc %d : issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", mode, n, nacc, double(h[0]) / iters,
                       double(h[1]) / iters, cudaGetErrorString(e));
            }
    cudaFuncSetAttribute(bench_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {16, 64}) {
        bench_lean<<<1, 128, 200 * 1024>>>(n, 27 * 150, d, 35200);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("lean N %3d : issue %.1f cyc/mma, complete %.1f cyc/mma (%s)\n", n, double(h[0]) / (27 * 150), double(h[1]) / (27 * 150), cudaGetErrorString(e));
    }
    cudaFuncSetAttribute(bench_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {16, 32})
        for (int variant : {1, 2, 3}) {
            bench_mn<<<1, 128, 200 * 1024>>>(n, 18 * 200, d, variant);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mn-major N %3d variant %d : issue %.1f cyc/mma, complete %.1f (%s)\n", n, variant, double(h[0]) / (18 * 200), double(h[1]) / (18 * 200), cudaGetErrorString(e));
        }
    cudaFuncSetAttribute(bench_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {16, 32})
        for (int nw : {1, 2, 4, 8}) {
            bench_multi<<<1, 256, 200 * 1024>>>(n, 27 * 150, nw, d, 35200);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("multi N %3d warps %d : %.1f cyc per MMA aggregate (%s)\n", n, nw, double(h[0]) / (27 * 150 * nw), cudaGetErrorString(e));
        }
    cudaFuncSetAttribute(bench_elect, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {16, 32, 64}) {
        bench_elect<<<1, 128, 200 * 1024>>>(n, 27 * 150, d, 35200);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("elect N %3d : issue %.1f cyc/mma, complete %.1f cyc/mma (%s)\n", n, double(h[0]) / (27 * 150), double(h[1]) / (27 * 150), cudaGetErrorString(e));
    }
    return 0;
}
