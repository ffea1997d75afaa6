// Microbenchmark + layout check: tcgen05.mma kind::f16 with BOTH operands MN-major in the hardware swizzle layouts
// (SWIZZLE_32B / 64B / 128B = a dense NDHWC row of 16 / 32 / 64 fp16 channels per voxel, K = consecutive voxels), as a function of
// M, N and the operand strides.  Question behind it: the weight-gradient GEMM multiplies two NDHWC tensors over the
// voxel axis, i.e. both operands are MN-major; with SWIZZLE_NONE core matrices an MN-major MMA costs 122 clk for every N
// (tools/mma_bench.cu).  Does the swizzled form run at the K-major rate, and may the N-side atoms OVERLAP (LBO = one voxel) so that
// the three dx taps come from one copy of a dy row?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_mn_bench tools/mma_mn_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../unet-studio_b200/csrc/common.cuh"
using namespace u3d;

// layout_type: 0 none, 6 = 32B, 4 = 64B, 2 = 128B
__device__ __forceinline__ uint64_t desc_any(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(layout_type) << 61;
    return d;
}

struct Cfg {
    int m, n;            // MMA shape (K = 16)
    int cw;              // channels per voxel row of the layout: 16 (SW32), 32 (SW64), 64 (SW128)
    int a_lbo, b_lbo;    // byte stride between consecutive MN atoms (cw channels each) of A / B
    int iters;
    int verify;
};

// smem: region A at 0, region B at 96 KB.  Voxel row v of a region = cw fp16 at byte v*cw*2, 16-byte chunks XOR-swizzled by the
// hardware pattern of the mode (chunk ^= (row_in_atom) for 128B: bits [4,7) ^= bits [7,10); 64B: bits [4,6) ^= [7,9); 32B: bit 4 ^= bit 7).
__device__ __forceinline__ uint32_t swz(uint32_t byte_off, int cw) {
    if (cw == 64) return byte_off ^ (((byte_off >> 7) & 7) << 4);
    if (cw == 32) return byte_off ^ (((byte_off >> 7) & 3) << 4);
    return byte_off ^ (((byte_off >> 7) & 1) << 4);
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out, float* dout, const __half* ain, const __half* bin, int nvox) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    // fill: voxel rows from global (A region then B region), swizzled
    const uint32_t breg = 96 * 1024;
    for (int i = threadIdx.x; i < nvox * c.cw; i += blockDim.x) {
        const int v = i / c.cw, ch = i % c.cw;
        const uint32_t off = uint32_t(v) * c.cw * 2 + ch * 2;
        *reinterpret_cast<__half*>(smem + swz(off, c.cw)) = ain[i];
        *reinterpret_cast<__half*>(smem + breg + swz(off, c.cw)) = bin[i];
    }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tptr;
    const uint32_t lt = c.cw == 64 ? 2u : c.cw == 32 ? 4u : 6u;
    const uint32_t sbo = 8u * c.cw * 2;    // next 8 voxels
    if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t idesc = umma_idesc(c.m, c.n, 0, 0, 1, 1);
        long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) {
            // different start voxel per MMA like a tap loop (verification: only i = 0 counts, start 0)
            const uint32_t sh = c.verify ? 0u : uint32_t(i % 7) * c.cw * 2;
            const uint64_t ad = desc_any(sb + sh, c.a_lbo, sbo, lt);
            const uint64_t bd = desc_any(sb + breg + sh, c.b_lbo, sbo, lt);
            umma_f16(tm + (c.verify ? 0 : (i & 1) * 256), ad, bd, idesc, (i >= 2 && !c.verify) ? 1u : 0u);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, 0xF00);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (c.verify) {
        const int warp = threadIdx.x / 32;
        for (int c0 = 0; c0 < c.n; c0 += 16) {
            float v[16];
            tmem_ld16(tm + (uint32_t(warp * 32) << 16) + c0, v);
            if (int(threadIdx.x) < c.m)
                for (int j = 0; j < 16; ++j) dout[threadIdx.x * c.n + c0 + j] = v[j];
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    float* dout; cudaMalloc(&dout, 128 * 256 * 4);
    const int nvox = 512;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int cw : {16, 32, 64}) {
        std::vector<__half> ha(nvox * cw), hb(nvox * cw);
        for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2half(float(int(i * 7 % 13) - 6) * 0.25f); hb[i] = __float2half(float(int(i * 5 % 11) - 5) * 0.5f); }
        __half *da, *db;
        cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2);
        cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
        const int row = cw * 2;   // bytes per voxel
        // ---- verification: A atoms = rows of 40 voxels apart (a "tap row" stride), B atoms = OVERLAPPING, one voxel apart
        {
            Cfg c{128, 3 * cw <= 256 ? 3 * cw : cw, cw, 40 * row, row, 1, 1};
            if (cw == 64) { c.m = 128; c.n = 192; }
            bench<<<1, 128, 200 * 1024>>>(c, d, dout, da, db, nvox);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> ho(128 * 256);
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < c.m; ++m)
                for (int n = 0; n < c.n; ++n) {
                    const int ga = m / cw, ca = m % cw, gb = n / cw, cb = n % cw;
                    double ref = 0;
                    for (int k = 0; k < 16; ++k) {
                        const int va = ga * 40 + k, vb = gb * 1 + k;
                        ref += double(__half2float(ha[va * cw + ca])) * double(__half2float(hb[vb * cw + cb]));
                    }
                    maxerr = std::max(maxerr, std::fabs(ref - ho[m * c.n + n]));
                }
            printf("verify cw %2d (SW%dB) M %d N %d, A atoms 40 voxels apart, B atoms overlapping by one voxel: max |err| = %g (%s)\n", cw, row, c.m, c.n,
                   maxerr, cudaGetErrorString(e));
        }
        // ---- timing
        for (int m : {128})
            for (int n : {16, 32, 48, 64, 96, 128, 144, 192, 256}) {
                if (n % cw && cw > 16 && n % 16) continue;
                if (n < cw && n % 16) continue;
                for (int overlap : {0, 1}) {
                    Cfg c{m, n, cw, 40 * row, overlap ? row : 24 * row, 2048, 0};
                    bench<<<1, 128, 200 * 1024>>>(c, d, dout, da, db, nvox);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("cw %2d M %3d N %3d %s : issue %.1f clk/mma, complete %.1f clk/mma -> %.0f FLOP/clk (%s)\n", cw, m, n,
                           overlap ? "B overlapping atoms" : "B separate atoms   ", double(h[0]) / c.iters, double(h[1]) / c.iters,
                           2.0 * m * n * 16 / (double(h[1]) / c.iters), cudaGetErrorString(e));
                    if (e != cudaSuccess) return 1;
                }
            }
        // ---- atom strides that are NOT multiples of 128 bytes (are the 114 clk a bank-conflict artefact of the strides above?)
        for (int pad : {0, 32, 64, 96})
            for (int n : {48, 192}) {
                Cfg c{128, n, cw, 40 * row + pad, 24 * row + pad, 2048, 0};
                bench<<<1, 128, 200 * 1024>>>(c, d, dout, da, db, nvox);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("cw %2d M 128 N %3d atom strides %d / %d bytes (pad %d): %.1f clk/mma (%s)\n", cw, n, c.a_lbo, c.b_lbo, pad,
                       double(h[1]) / c.iters, cudaGetErrorString(e));
            }
        cudaFree(da); cudaFree(db);
    }
    return 0;
}
