// Weight gradient of the 16-channel 3x3x3 stride-1 convolutions of the full-resolution level ("quad" kernel: 2 x 2 output rows
// against 4 x 4 input rows per MMA pair).
//
//   dW[co][ci][dz,dy,dx] = sum_v x[v + (dz,dy,dx)][ci] * g[v][co]          (g = gradient of the conv output)
//
// Both operands are NDHWC rows, i.e. MN-major with the reduction index (voxels along x) strided by one voxel row of 32 bytes: exactly
// the SWIZZLE_32B MN-major canonical layout (8 voxels x 32 B per atom, next 8 voxels 256 B further).  Measured
// (tools/mma_mn_bench.cu, profiles/r02_mma_mn_bench.txt): an MN-major tcgen05.mma M = 128, K = 16 costs 114 clk for EVERY N <= 192
// and 128 clk at N = 256, in every swizzle mode, so the only lever is useful work per instruction.  conv_wgrad_band.cu stacks
// (8 x rows) x (3 g rows x 3 dx copies): 9 of its 24 (row, row) pairs are taps, 1 MMA per g row and 16 voxels.  Here
//   M = 8 atoms = (ay' in {0,1}) x (az in 0..3) x 16 ci     x rows of FOUR planes z0-1..z0+2, two y rows  -> two accumulators
//                                                           D1 (y rows y-1, y) and D2 (y rows y+1, y+2)
//   N = 12 atoms = (by in {0,1}) x (bz in {0,1}) x (3 copies shifted by dx) x 16 co = 192
//   tap (dz, dy) = (az - bz - 1, ay - by - 1): 36 of the 64 (x row, g row) pairs are taps (each of the 9 (dz,dy) taps four times,
//   once per g row), 2 MMAs per FOUR g rows and 16 voxels = half the instructions of the band kernel at the same 114 clk.
// The CTA owns a column (xt voxels in x, 2 planes in z) and marches along y two rows at a time through a ring of x row pairs
// (each pair = 4 planes x 2 y rows = ONE tensor-map box, so the 8 atoms of an operand are one row apart) and a ring of g slots
// (3 boxes of 2 planes x 2 rows, fetched at x - dx).  Everything arrives by TMA (cp.async.bulk.tensor with the 32-byte hardware
// swizzle): the first version fed the rings with 16-byte cp.async like conv_wgrad_band and both kernels stopped at ~12.5 bytes per
// clock and SM of shared-memory fill (profiles/r02_ncu_wgrad_quad_cpasync.md), far below what the MMAs consume.  Accumulators live
// in TMEM for the CTA's whole life; one epilogue adds the 36 useful blocks into the reference-layout gradient with fp32 atomics.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kQThreads = 32 * 6;      // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer
constexpr int kQIssuerWarp = 5;
constexpr int kQXSlots = 4, kQYSlots = 3;
// tile width along x (xt voxels = xt/16 K steps per ring slot) is chosen per problem: a slot must carry enough MMA time (114 clk per
// instruction) to cover the L2 latency of the slot being refilled -- with 32 voxels (456 clk per step) the ring ran dry

struct alignas(64) WQMaps {
    CUtensorMap x, g, gh;      // dims (16 ch, W, H, D) over the channel slice of each tensor, SWIZZLE_32B; boxes (16, xt, 2, 4) / (16, xt, 2, 2)
                               // and the one-voxel halo columns of g (16, 1, 2, 2)
};

__device__ __forceinline__ void tma_box_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct WQParams {
    WgradProblem P;
    int tiles_x, zpairs, ychunks, ylen, total_items;
    int xt;                    // voxels per tile along x (multiple of 16)
    uint32_t row_b;            // bytes of one atom row = xt * 32
    uint32_t xslot_b, yslot_b, off_y, off_bars;
    float* scratch;            // != nullptr: the epilogue stores the CTA's gradient block [16 co][16 ci][27] here (wgrad_block_sum_launch adds them)
};

__device__ __forceinline__ uint64_t desc_sw32_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(6) << 61;   // SWIZZLE_32B
    return d;
}
// SWIZZLE_32B: the two 16-byte halves of a 32-byte row are exchanged in rows 4..7 of every 8-row (256-byte) atom
__device__ __forceinline__ uint32_t sw32(uint32_t addr) { return addr ^ (((addr >> 7) & 1u) << 4); }

__global__ void __launch_bounds__(kQThreads, 1) conv_wgrad_quad_kernel(const __grid_constant__ WQParams p, const __grid_constant__ WQMaps maps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t bars = sbase + p.off_bars;
    const uint32_t kQRow = p.row_b, kQXSlotB = p.xslot_b, kQYSlotB = p.yslot_b, kQOffY = p.off_y;
    const int kQXT = p.xt;
    auto xfull = [&](int s) { return bars + 8u * s; };
    auto xempty = [&](int s) { return bars + 8u * (kQXSlots + s); };
    auto yfull = [&](int s) { return bars + 8u * (2 * kQXSlots + s); };
    auto yempty = [&](int s) { return bars + 8u * (2 * kQXSlots + kQYSlots + s); };
    auto ycopy = [&](int s) { return bars + 8u * (2 * kQXSlots + 2 * kQYSlots + s); };   // the shifted copies of a g slot are in place
    const uint32_t done_bar = bars + 8u * (2 * kQXSlots + 3 * kQYSlots);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * kQXSlots + 3 * kQYSlots + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kQXSlots; ++s) { mbar_init(xfull(s), 1); mbar_init(xempty(s), 1); }
        for (int s = 0; s < kQYSlots; ++s) { mbar_init(yfull(s), 1); mbar_init(yempty(s), 1); mbar_init(ycopy(s), 128); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == kQIssuerWarp) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const WgradProblem& P = p.P;
    const int D = P.t_d, H = P.t_h, W = P.t_w;
    const bool has_work = int(blockIdx.x) < p.total_items;

    if (warp == 4) {
        // ===================================== TMA producer (one thread) =====================
        if (lane == 0) {
            uint32_t xcnt = 0, ycnt = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int rem = item;
                const int yc = rem % p.ychunks; rem /= p.ychunks;
                const int tx = rem % p.tiles_x;
                const int zp = rem / p.tiles_x;
                const int x0 = tx * kQXT, z0 = zp * 2, y0 = yc * p.ylen;
                const int y1 = min(H, y0 + p.ylen);
                const int ns = (y1 - y0 + 1) / 2;
                for (int pr = 0; pr <= ns; ++pr) {
                    {   // x pair pr: planes z0-1..z0+2, rows y0 + 2 pr - 1, y0 + 2 pr; row index in the slot = az*2 + ay'.  Out of range = 0.
                        const int slot = xcnt % kQXSlots;
                        mbar_wait(xempty(slot), ((xcnt / kQXSlots) & 1) ^ 1, 0x3500u | slot);
                        mbar_arrive_expect_tx(xfull(slot), kQXSlotB);
                        tma_box_4d(sbase + slot * kQXSlotB, &maps.x, xfull(slot), 0, x0, y0 + 2 * pr - 1, z0 - 1);
                        ++xcnt;
                    }
                    if (pr < ns) {   // g slot pr: rows y0 + 2 pr + by of the planes z0 + bz (row bz*2 + by) -> middle copy (dx = 0) + halo columns
                        const int slot = ycnt % kQYSlots;
                        mbar_wait(yempty(slot), ((ycnt / kQYSlots) & 1) ^ 1, 0x3600u | slot);
                        mbar_arrive_expect_tx(yfull(slot), 4u * kQRow + 256u);
                        const uint32_t blk = sbase + kQOffY + slot * kQYSlotB;
                        tma_box_4d(blk + 4u * kQRow, &maps.g, yfull(slot), 0, x0, y0 + 2 * pr, z0);
                        tma_box_4d(blk + 12u * kQRow, &maps.gh, yfull(slot), 0, x0 - 1, y0 + 2 * pr, z0);
                        tma_box_4d(blk + 12u * kQRow + 128u, &maps.gh, yfull(slot), 0, x0 + kQXT, y0 + 2 * pr, z0);
                        ++ycnt;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kQIssuerWarp) {
        // ===================================== MMA issuer =====================================
        if (has_work) {
            const uint32_t idesc = umma_idesc(128, 192, 0, 0, 1, 1);    // both operands MN-major
            const uint32_t d1 = tmem_base, d2 = tmem_base + 256u;
            uint32_t xcnt = 0, ycnt = 0;
            bool first = true;
            const uint32_t ksteps = uint32_t(kQXT / 16);
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int yc = item % p.ychunks;
                const int y0 = yc * p.ylen, y1 = min(H, y0 + p.ylen);
                const int ns = (y1 - y0 + 1) / 2;
#pragma unroll 1
                for (int j = 0; j < ns; ++j, ++ycnt) {
                    const uint32_t c1 = xcnt + j, c2 = c1 + 1;
                    mbar_wait(xfull(c1 % kQXSlots), (c1 / kQXSlots) & 1, 0x3701u);
                    mbar_wait(xfull(c2 % kQXSlots), (c2 / kQXSlots) & 1, 0x3702u);
                    mbar_wait(ycopy(ycnt % kQYSlots), (ycnt / kQYSlots) & 1, 0x3703u);
                    tc_fence_after();
                    const uint64_t a1 = desc_sw32_mn(sbase + (c1 % kQXSlots) * kQXSlotB, kQRow, 256u);
                    const uint64_t a2 = desc_sw32_mn(sbase + (c2 % kQXSlots) * kQXSlotB, kQRow, 256u);
                    const uint64_t b = desc_sw32_mn(sbase + kQOffY + (ycnt % kQYSlots) * kQYSlotB, kQRow, 256u);
                    const bool f = first;
                    first = false;
                    if (elect_one()) {
                        if (f) { umma_f16_first(d1, a1, b, idesc); umma_f16_first(d2, a2, b, idesc); }
                        else { umma_f16_acc(d1, a1, b, idesc); umma_f16_acc(d2, a2, b, idesc); }
#pragma unroll 1
                        for (uint32_t ks = 1; ks < ksteps; ++ks) {       // next K step: 16 voxels = 512 bytes further
                            umma_f16_acc(d1, a1 + 32u * ks, b + 32u * ks, idesc);
                            umma_f16_acc(d2, a2 + 32u * ks, b + 32u * ks, idesc);
                        }
                        umma_commit(yempty(ycnt % kQYSlots));
                        umma_commit(xempty(c1 % kQXSlots));              // pair j: last used here (as the D1 operand)
                        if (j == ns - 1) umma_commit(xempty(c2 % kQXSlots));
                    }
                    __syncwarp();
                }
                xcnt += uint32_t(ns + 1);
            }
            if (elect_one()) umma_commit(done_bar);
        }
        __syncwarp();
    } else if (has_work) {
        // ============ g replication during the march: the copy engine delivers every g row ONCE (the dx = 0 copy + two halo columns);
        // the copies shifted by dx = -1 / +1 are made here, shared memory to shared memory, by the warps that otherwise only wait for
        // the epilogue.  The fill rate of the rings (~15 bytes per clock and SM with 32-byte rows, TMA or cp.async alike), not the
        // tensor pipe, paces this kernel: three fetched copies were 60 % of the fill.
        {
            uint32_t ycnt = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int yc = item % p.ychunks;
                const int y0 = yc * p.ylen, y1 = min(H, y0 + p.ylen);
                const int ns = (y1 - y0 + 1) / 2;
#pragma unroll 1
                for (int j = 0; j < ns; ++j, ++ycnt) {
                    const int slot = ycnt % kQYSlots;
                    mbar_wait(yfull(slot), (ycnt / kQYSlots) & 1, 0x3900u | slot);
                    const uint32_t blk = sbase + kQOffY + slot * kQYSlotB;
                    const uint32_t mid = blk + uint32_t(4 + warp) * kQRow;        // warp w copies g row w
                    const uint32_t lo = blk + uint32_t(warp) * kQRow;             // dxc = 0: g[x' + 1]
                    const uint32_t hi = blk + uint32_t(8 + warp) * kQRow;         // dxc = 2: g[x' - 1]
                    const uint32_t halo = blk + 12u * kQRow + uint32_t(warp) * 32u;   // left column; right column 128 bytes further
                    for (int idx = lane; idx < 2 * kQXT; idx += 32) {
                        const uint32_t cgo = uint32_t(idx & 1) * 16u;
                        const int xv = idx >> 1;
                        uint4 v;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw32(mid + uint32_t(idx) * 16u)));
                        // this voxel is g[x'] for x' = xv - 1 of the lo copy and x' = xv + 1 of the hi copy
                        if (xv >= 1) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sw32(lo + uint32_t(idx - 2) * 16u)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                        if (xv + 1 < kQXT) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sw32(hi + uint32_t(idx + 2) * 16u)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                        if (xv == 0) {   // hi copy voxel 0 = left halo column
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw32(halo + cgo)));
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sw32(hi + cgo)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                        }
                        if (xv == kQXT - 1) {   // lo copy last voxel = right halo column
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw32(halo + 128u + cgo)));
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sw32(lo + uint32_t(idx) * 16u)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                        }
                    }
                    fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's operand reads
                    mbar_arrive(ycopy(slot));
                }
            }
        }
        // ===================================== epilogue (once) ================================
        mbar_wait(done_bar, 0, 0x3800u);
        tc_fence_after();
        const int m = threadIdx.x;                 // accumulator row = TMEM lane
        const int g = m >> 4, ci = m & 15;
        const int az = g >> 1, ayp = g & 1;        // atom g = az * 2 + ay' (box order: x fastest, then y, then z)
        const size_t nstride = size_t(P.w_mtot) * P.w_ktaps;
        const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
        const bool ci_ok = ci < P.t_creal;
        // partial-block mode (see conv_wgrad_band.cu): the 36 useful (x row, g row) blocks are summed per tap in shared memory (the rings
        // are idle once done_bar has fired) in the gradient's order [co][ci][27] and stored to this CTA's slot of the scratch buffer
        float* const stage = reinterpret_cast<float*>(smem);
        if (p.scratch != nullptr) {
            for (int i = m; i < 16 * 16 * 27; i += 128) stage[i] = 0.f;
            asm volatile("bar.sync 2, 128;" ::: "memory");
        }
#pragma unroll 1
        for (int acc = 0; acc < 2; ++acc) {
            const int ay = acc * 2 + ayp;
#pragma unroll 1
            for (int rw = 0; rw < 4; ++rw) {       // g row rw = bz * 2 + by
                const int bz = rw >> 1, by = rw & 1;
                const int dzi = az - bz, dyi = ay - by;    // = dz + 1, dy + 1
                const bool ok = ci_ok && dzi >= 0 && dzi <= 2 && dyi >= 0 && dyi <= 2;
#pragma unroll 1
                for (int dxc = 0; dxc < 3; ++dxc) {
                    float v[16];
                    tmem_ld16(t_row + uint32_t(acc * 256 + (dxc * 4 + rw) * 16), v);   // warp-collective: every lane takes part
                    if (ok) {
                        const int tap = (dzi * 3 + dyi) * 3 + dxc;
                        if (p.scratch != nullptr) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) atomicAdd(&stage[(j * 16 + ci) * 27 + tap], v[j]);
                        } else {
                            float* dwrow = P.dw + size_t(P.w_moff + ci) * P.w_ktaps + tap;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (j < P.u_creal) atomicAdd(dwrow + size_t(P.w_noff + j) * nstride, v[j]);
                        }
                    }
                }
            }
        }
        if (p.scratch != nullptr) {
            asm volatile("bar.sync 2, 128;" ::: "memory");
            float4* const dst = reinterpret_cast<float4*>(p.scratch + size_t(blockIdx.x) * (16 * 16 * 27));
            const float4* const src = reinterpret_cast<const float4*>(stage);
            for (int i = m; i < 16 * 16 * 27 / 4; i += 128) dst[i] = src[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kQIssuerWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn quad_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}
int encode_rows(CUtensorMap* m, const void* base, int coff, int cp, int W, int H, int D, int xt, int bz) {   // box (16, xt, 2, bz)
    const cuuint64_t gdim[4] = {16, cuuint64_t(W), cuuint64_t(H), cuuint64_t(D)};
    const cuuint64_t gstr[3] = {cuuint64_t(cp) * 2, cuuint64_t(cp) * 2 * W, cuuint64_t(cp) * 2 * W * H};
    const cuuint32_t box[4] = {16, cuuint32_t(xt), 2, cuuint32_t(bz)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    void* p = const_cast<uint8_t*>(static_cast<const uint8_t*>(base) + size_t(coff) * 2);
    const CUresult r = quad_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_wgrad_quad_launch: cuTensorMapEncodeTiled failed with code " + std::to_string(int(r))); return 1; }
    return 0;
}
}  // namespace

unsigned int read_device_error_wquad() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

bool conv_wgrad_quad_eligible(const WgradProblem& P) {
    static const bool disabled = std::getenv("U3D_NO_WQUAD") != nullptr || std::getenv("U3D_NO_WBAND") != nullptr;
    if (disabled || quad_encode_fn() == nullptr) return false;
    if (P.ntaps != 27 || P.tstride != 1 || P.w_ktaps != 27) return false;
    if (P.t_c != 16 || P.u_c != 16) return false;
    if (P.t_d != P.ld || P.t_h != P.lh || P.t_w != P.lw) return false;
    if (1LL * P.ld * P.lh * P.lw < 32768) return false;
    if ((P.t_cp * 2) % 32 || (P.u_cp * 2) % 32 || (P.t_coff * 2) % 32 || (P.u_coff * 2) % 32) return false;   // 16-byte chunks of a 32-byte channel row
    if ((reinterpret_cast<uintptr_t>(P.T) & 15) || (reinterpret_cast<uintptr_t>(P.U) & 15)) return false;
    for (int t = 0; t < 27; ++t)   // forward tap order (kz,ky,kx) with offsets k-1 and identity tap_ref
        if (P.taps[t].dz != t / 9 - 1 || P.taps[t].dy != (t / 3) % 3 - 1 || P.taps[t].dx != t % 3 - 1 || P.tap_ref[t] != t) return false;
    return true;
}

int conv_wgrad_quad_launch(const WgradProblem& P, cudaStream_t stream, float* partial_scratch, size_t partial_scratch_bytes) {
    WQParams wp;
    std::memset(&wp, 0, sizeof(wp));
    wp.P = P;
    {   // widest tile that fits (4 x slots of 8 rows + 3 g slots of 12 rows), least padding of the last tile first
        int best = 32;
        double best_cost = 1e30;
        for (int xt = 32; xt <= 96; xt += 16) {
            const size_t need = size_t(kQXSlots * 8 + kQYSlots * 12) * xt * 32 + kQYSlots * 256 + 2048;
            if (need > 225 * 1024) break;
            const int tiles = (P.lw + xt - 1) / xt;
            const double cost = double(tiles) * xt / double(P.lw) + 4.0 / xt;   // padded work + a penalty for short slots
            if (cost < best_cost - 1e-9) { best_cost = cost; best = xt; }
        }
        wp.xt = best;
    }
    const int kQXT = wp.xt;
    wp.row_b = uint32_t(kQXT) * 32u;
    wp.xslot_b = 8 * wp.row_b;
    wp.yslot_b = 12 * wp.row_b + 256;   // + the two halo columns (4 rows x 32 bytes each)
    wp.off_y = kQXSlots * wp.xslot_b;
    wp.off_bars = wp.off_y + kQYSlots * wp.yslot_b;
    const size_t kQSmem = wp.off_bars + 8 * (2 * kQXSlots + 3 * kQYSlots + 1) + 16 + 1024;   // + slack for the 1024-byte alignment
    wp.tiles_x = (P.lw + kQXT - 1) / kQXT;
    wp.zpairs = (P.ld + 1) / 2;
    const int sms = device_sm_count();
    const long long cols = 1LL * wp.tiles_x * wp.zpairs;
    // y chunk length: even, balances (items per wave) against the one extra x pair every chunk loads
    int best_yc = 1;
    double best_eff = -1;
    for (int yc = 1; yc <= std::max(1, P.lh / 8); ++yc) {
        int yl = (P.lh + yc - 1) / yc;
        yl += yl & 1;
        const int yc_eff = (P.lh + yl - 1) / yl;
        const long long items = cols * yc_eff;
        const long long waves = (items + sms - 1) / sms;
        const double eff = double(items) / double(waves * sms) * double(yl) / double(yl + 2);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_yc = yc_eff; }
    }
    int yl = (P.lh + best_yc - 1) / best_yc;
    yl += yl & 1;
    wp.ylen = yl;
    wp.ychunks = (P.lh + yl - 1) / yl;
    wp.total_items = int(cols * wp.ychunks);
    const int grid = std::max(1, std::min(wp.total_items, sms));
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    WQMaps maps;
    if (encode_rows(&maps.x, P.T, P.t_coff, P.t_cp, P.t_w, P.t_h, P.t_d, kQXT, 4)) return 1;
    if (encode_rows(&maps.g, P.U, P.u_coff, P.u_cp, P.lw, P.lh, P.ld, kQXT, 2)) return 1;
    if (encode_rows(&maps.gh, P.U, P.u_coff, P.u_cp, P.lw, P.lh, P.ld, 1, 2)) return 1;
    static const bool no_partial = std::getenv("U3D_WBAND_ATOMICS") != nullptr;
    if (!no_partial && partial_scratch != nullptr && size_t(grid) * 16 * 16 * 27 * 4 <= partial_scratch_bytes) wp.scratch = partial_scratch;
    conv_wgrad_quad_kernel<<<grid, kQThreads, kQSmem, stream>>>(wp, maps);
    U3D_CUDA_CHECK(cudaGetLastError());
    if (wp.scratch != nullptr) return wgrad_block_sum_launch(wp.scratch, P, 1, 1, grid, 16, 16, stream);
    return 0;
}

}  // namespace u3d
