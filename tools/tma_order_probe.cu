// Probe: does cuTensorMapEncodeTiled accept a tensor map whose dimensions are NOT ordered by stride -- (C, Y, Z, X) over an NDHWC
// tensor, i.e. strides (W*32, H*W*32, 32) bytes -- and does the copy engine then write a box (16 ch, 2 y, 2 z, XT x) as
// smem[x][z][y][c] = one 128-byte line per x position with SWIZZLE_128B applied on the linear shared-memory address?
// That layout makes the four g rows of conv_wgrad_quad one MN-major SWIZZLE_128B operand whose dx taps are start-address shifts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tma_order_probe tools/tma_order_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap map, __half* out, int xt, int cy, int cz, int cx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = uint32_t(xt) * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(b), "r"(0), "r"(cy), "r"(cz), "r"(cx) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < xt * 64; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

int main() {
    const int W = 40, H = 12, D = 6, C = 16, XT = 16;
    std::vector<__half> h(size_t(W) * H * D * C);
    for (int z = 0; z < D; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < C; ++c) h[((size_t(z) * H + y) * W + x) * C + c] = __float2half(float(z * 1000 + y * 100 + x) + c / 32.0f);
    __half *d, *o;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&o, XT * 64 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const cuuint64_t gdim[4] = {cuuint64_t(C), cuuint64_t(H), cuuint64_t(D), cuuint64_t(W)};
    const cuuint64_t gstr[3] = {cuuint64_t(W) * C * 2, cuuint64_t(H) * W * C * 2, cuuint64_t(C) * 2};
    const cuuint32_t box[4] = {16, 2, 2, cuuint32_t(XT)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    cuInit(0);
    const CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (C,Y,Z,X) strides (%llu, %llu, %llu): CUresult %d\n", (unsigned long long)gstr[0], (unsigned long long)gstr[1],
           (unsigned long long)gstr[2], int(r));
    if (r != CUDA_SUCCESS) return 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int trial = 0; trial < 2; ++trial) {
        const int cy = trial ? -1 : 3, cz = trial ? 5 : 2, cx = trial ? 30 : 7;   // second box hangs over y < 0, z >= D and x >= W
        probe<<<1, 128, 64 * 1024>>>(map, o, XT, cy, cz, cx);
        const cudaError_t e = cudaDeviceSynchronize();
        std::vector<__half> got(XT * 64);
        cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int x = 0; x < XT; ++x)
            for (int zl = 0; zl < 2; ++zl)
                for (int yl = 0; yl < 2; ++yl)
                    for (int c = 0; c < 16; ++c) {
                        const int gz = cz + zl, gy = cy + yl, gx = cx + x;
                        const bool in = gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W;
                        const float want = in ? __half2float(h[((size_t(gz) * H + gy) * W + gx) * C + c]) : 0.f;
                        uint32_t off = uint32_t(x) * 128u + uint32_t(zl * 2 + yl) * 32u + uint32_t(c) * 2u;   // smem[x][z][y][c]
                        off ^= ((off >> 7) & 7u) << 4;                                                    // SWIZZLE_128B on the linear address
                        if (__half2float(got[off / 2]) != want) ++bad;
                    }
        printf("box at (y %d, z %d, x %d): %s, %d mismatches of %d (layout smem[x][z][y][c] + 128B swizzle)\n", cy, cz, cx, cudaGetErrorString(e), bad,
               XT * 64);
        if (trial == 0)
            for (int i = 0; i < 1024; i += 8) {   // one line per 16-byte chunk: what sits there
                const float v = __half2float(got[i]);
                const int iv = int(v);
                printf("  chunk %3d (byte %4d): z %d y %2d x %2d c %2d\n", i / 8, i * 2, iv / 1000, (iv / 100) % 10, iv % 100, int((v - iv) * 32 + 0.5f));
            }
    }
    return 0;
}
