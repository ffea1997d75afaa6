// N-stacked weight gradient for the 3x3x3 stride-1 convolutions of the large-volume levels (16 / 32 channels).
//
//   dW[co][ci][dz,dy,dx] = sum_v x[v + (dz,dy,dx)][ci] * dy[v][co]
//
// Both operands come straight from the NDHWC tensors, i.e. MN-major (channels contiguous, the reduction index = voxels along
// x strided).  Measured (tools/mma_bench.cu): an MN-major tcgen05.mma M = 128, K = 16 costs ~122 clk whatever N is (16..192), so
// the kernel is organised to make N as wide as TMEM allows:
//   M = (a: 16/ncg consecutive x-rows hy0+a) x Cin            (one descriptor: the row/channel-group runs are 512 B apart)
//   N = (b: R consecutive dy-rows ly0+b) x (3 copies of the row shifted by dx = -1,0,+1) x Cout      (R*3*Cout <= 144 columns)
//   K = 16 of the 32 INPUT x positions of the tile (the dy copies carry the +-1 halo, x does not)
//   D[(a,ci)][(b,dx,co)] += sum_x' x[z+dz][hy0+a][x'][ci] * dy[z][ly0+b][x'-dx][co]      -> tap (dz, dy = a-b-1, dx) if |dy| <= 1
// One accumulator per dz (3 x N columns of TMEM) lives for the CTA's whole life; per R rows x 32 voxels the CTA issues
// 3 (dz) x 2 (K steps) MMAs instead of the 18 of conv_wgrad_rows.cu.  The CTA marches along z through a ring of x planes
// (4 slots: z-1, z, z+1 in use, one loading) and dy planes (2 slots), so every plane is fetched once per (x,y) tile column.
// One epilogue at the end adds the useful (a,b) blocks into the reference-layout gradient with fp32 atomics.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kWThreads = 32 * 15;   // warps 0-3 epilogue, 4-11 producers, 12-14 MMA issuers (one per dz)
constexpr int kWProducers = 256;
constexpr int kXSlots = 4, kYSlots = 3;   // dy ring of 3: with 2 the producers block on it and the x planes behind it arrive late (ncu)
constexpr int kRun = 33;             // a run of 32 voxels (512 B) + 16 B pad: the channel-group runs a warp writes fall into different banks
constexpr uint32_t kRunB = kRun * 16u;

struct WBParams {
    WgradProblem P;
    int TY, HYA;               // rows per tile, allocated x rows per plane (TY + 2 halo + junk rows the M = 128 descriptor runs into)
    int tiles_x, tiles_y, zchunks, zlen, total_items;
    int ngo, npairs, cpp;      // channel-group pairs (gi, go) of a wide layer: CTA b owns pair b % npairs and every cpp-th tile of it
    uint32_t x_slot_bytes, y_slot_bytes, off_y, off_bars;
    float* scratch;            // != nullptr: the epilogue stores the CTA's gradient block [CO][NCG*8][27] here (wgrad_band_sum_kernel adds them up)
};

template <int NCG, int CO>
__global__ void __launch_bounds__(kWThreads, 1) conv_wgrad_band_kernel(const __grid_constant__ WBParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NCGY = CO / 8;
    constexpr int AROWS = 16 / NCG;                         // x rows covered by one M = 128 operand
    constexpr int R = (CO == 16 ? 3 : 1) < AROWS - 2 ? (CO == 16 ? 3 : 1) : AROWS - 2;   // dy rows per MMA
    constexpr int N = R * 3 * CO;
    constexpr int TMEM_COLS = 3 * N <= 256 ? 256 : 512;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + p.off_bars;
    auto xfull = [&](int s) { return bars + 8u * s; };
    auto xempty = [&](int s) { return bars + 8u * (kXSlots + s); };
    auto yfull = [&](int s) { return bars + 8u * (2 * kXSlots + s); };
    auto yempty = [&](int s) { return bars + 8u * (2 * kXSlots + kYSlots + s); };
    const uint32_t done_bar = bars + 8u * (2 * kXSlots + 2 * kYSlots);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * kXSlots + 2 * kYSlots + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXSlots; ++s) { mbar_init(xfull(s), kWProducers); mbar_init(xempty(s), 3); }
        for (int s = 0; s < kYSlots; ++s) { mbar_init(yfull(s), kWProducers); mbar_init(yempty(s), 3); }
        mbar_init(done_bar, 3);
        fence_barrier_init();
    }
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const WgradProblem& P = p.P;
    const int D = P.t_d, H = P.t_h, W = P.t_w;
    // wide layers (Cin, Cout > 32) are cut into channel-group pairs: each CTA accumulates ONE (gi, go) block of the gradient
    const int pair = int(blockIdx.x) % p.npairs, rank = int(blockIdx.x) / p.npairs;
    const int gi = pair / p.ngo, go = pair % p.ngo;
    const bool has_work = rank < p.total_items;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        const uint8_t* const xsrc = static_cast<const uint8_t*>(P.T) + (P.t_coff + gi * NCG * 8) * 2;
        const uint8_t* const ysrc = static_cast<const uint8_t*>(P.U) + (P.u_coff + go * CO) * 2;
        const uint32_t xpitch = uint32_t(P.t_cp) * 2u, ypitch = uint32_t(P.u_cp) * 2u;
        const int TY = p.TY;
        const int xtotal = (TY + 2) * 32 * NCG;
        const int ytotal = TY * 32 * NCGY;
        uint32_t xcnt = 0, ycnt = 0;
        for (int item = rank; item < p.total_items; item += p.cpp) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int x0 = tx * 32, y0 = ty * TY;
            const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
            for (int gz = z0 - 1; gz <= z1; ++gz) {
                {   // x plane gz (rows y0-1 .. y0+TY), no halo along x
                    const int slot = xcnt % kXSlots;
                    mbar_wait(xempty(slot), ((xcnt / kXSlots) & 1) ^ 1, 0x3100u | slot);
                    const uint32_t blk = sbase + slot * p.x_slot_bytes;
                    const bool zok = (unsigned)gz < (unsigned)D;
                    const uint8_t* const xpl = xsrc + (((long long)(zok ? gz : 0) * H + (y0 - 1)) * W + x0) * (long long)xpitch;
#pragma unroll 2
                    for (int idx = t; idx < xtotal; idx += kWProducers) {
                        const int cg = idx % NCG;
                        const int q = idx / NCG;
                        const int lx = q % 32, hy = q / 32;
                        const bool ok = zok && x0 + lx < W && (unsigned)(y0 + hy - 1) < (unsigned)H;
                        const uint8_t* src = ok ? xpl + (long long)(hy * W + lx) * xpitch + cg * 16 : xsrc;
                        cp_async16_ca(blk + uint32_t((hy * NCG + cg) * kRun + lx) * 16u, src, ok ? 16u : 0u);
                    }
                    cp_async_mbar_arrive(xfull(slot));
                    ++xcnt;
                }
                if (gz >= z0 && gz < z1) {   // dy plane gz: per row three copies shifted by dx = -1, 0, +1 (real neighbours, zero outside)
                    const int slot = ycnt % kYSlots;
                    mbar_wait(yempty(slot), ((ycnt / kYSlots) & 1) ^ 1, 0x3200u | slot);
                    const uint32_t blk = sbase + p.off_y + slot * p.y_slot_bytes;
                    const uint8_t* const ypl = ysrc + (((long long)gz * H + y0) * W + x0) * (long long)ypitch;
#pragma unroll 2
                    for (int idx = t; idx < ytotal; idx += kWProducers) {
                        const int cg = idx % NCGY;
                        const int q = idx / NCGY;
                        const int lx = q % 32, ly = q / 32;
                        const bool yok = y0 + ly < H;
                        const uint8_t* const s = ypl + (long long)(ly * W + lx) * ypitch + cg * 16;
                        const uint32_t d0 = blk + uint32_t((ly * 3 * NCGY + cg) * kRun + lx) * 16u;
#pragma unroll
                        for (int dxc = 0; dxc < 3; ++dxc) {   // copy dxc holds dy[x' - dx], dx = dxc - 1
                            const int gx = x0 + lx - (dxc - 1);
                            const bool ok = yok && (unsigned)gx < (unsigned)W;
                            cp_async16_ca(d0 + uint32_t(dxc * NCGY * kRun) * 16u, ok ? s - (long long)(dxc - 1) * ypitch : ysrc, ok ? 16u : 0u);
                        }
                    }
                    cp_async_mbar_arrive(yfull(slot));
                    ++ycnt;
                }
            }
        }
        cp_async_wait<0>();
    } else if (warp >= 12) {
        // ===================================== MMA issuers ===================================
        // one issuing thread per dz (own accumulator): a single thread's issue loop, not the tensor pipe, limits the MMA rate
        // The whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane issues.
        const int dz = warp - 12;
        if (has_work) {
            const uint32_t idesc = umma_idesc(128, N, 0, 0, 1, 1);    // both operands MN-major
            const uint64_t a_rows_u = uint64_t((R * NCG * kRunB) >> 4);       // R x rows further
            const uint64_t b_rows_u = uint64_t((R * 3 * NCGY * kRunB) >> 4);  // R dy rows further
            const uint32_t d_tmem = tmem_base + uint32_t(dz * N);
            const int nrg = p.TY / R;
            uint32_t xcnt = 0, ycnt = 0;
            bool first = true;
            for (int item = rank; item < p.total_items; item += p.cpp) {
                const int zc = item % p.zchunks;
                const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
                const int nz = z1 - z0;
                // x planes this issuer never reads (relative q < dz and q > nz-1+dz): release them once they are resident
                for (int q = 0; q < dz; ++q) {
                    const uint32_t c = xcnt + q;
                    mbar_wait(xfull(c % kXSlots), (c / kXSlots) & 1, 0x3300u);
                    if (lane == 0) mbar_arrive(xempty(c % kXSlots));
                    __syncwarp();
                }
#pragma unroll 1
                for (int j = 0; j < nz; ++j, ++ycnt) {
                    const uint32_t cq = xcnt + j + dz;     // x plane z + dz - 1
                    mbar_wait(xfull(cq % kXSlots), (cq / kXSlots) & 1, 0x3302u);
                    mbar_wait(yfull(ycnt % kYSlots), (ycnt / kYSlots) & 1, 0x3303u);
                    fence_proxy_async();
                    tc_fence_after();
                    uint64_t ad = umma_smem_desc(sbase + (cq % kXSlots) * p.x_slot_bytes, 128u, kRunB);
                    uint64_t bd = umma_smem_desc(sbase + p.off_y + (ycnt % kYSlots) * p.y_slot_bytes, 128u, kRunB);
                    const bool f = first;
                    first = false;
                    if (elect_one()) {
#pragma unroll 1
                        for (int rg = 0; rg < nrg; ++rg, ad += a_rows_u, bd += b_rows_u) {
                            if (f && rg == 0) umma_f16_first(d_tmem, ad, bd, idesc);
                            else umma_f16_acc(d_tmem, ad, bd, idesc);
                            umma_f16_acc(d_tmem, ad + 16u, bd + 16u, idesc);
                        }
                        umma_commit(yempty(ycnt % kYSlots));
                        umma_commit(xempty(cq % kXSlots));
                    }
                    __syncwarp();
                }
                for (int q = nz + dz; q <= nz + 1; ++q) {
                    const uint32_t c = xcnt + q;
                    mbar_wait(xfull(c % kXSlots), (c / kXSlots) & 1, 0x3304u);
                    if (lane == 0) mbar_arrive(xempty(c % kXSlots));
                    __syncwarp();
                }
                xcnt += uint32_t(nz + 2);
            }
            if (elect_one()) umma_commit(done_bar);
        }
        __syncwarp();
    } else if (has_work) {
        // ===================================== epilogue (once) ================================
        const int r = threadIdx.x;
        mbar_wait(done_bar, 0, 0x3400u);
        tc_fence_after();
        constexpr int CIG = NCG * 8;
        constexpr int RUN = CIG * 27;                 // floats per output channel of this CTA's block
        const int a = r / CIG;
        const int cil = r % CIG;
        const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
        if (p.scratch != nullptr) {
            // Partial-block mode.  Every launch ends in ~4 M fp32 atomics (27*Cin*Cout elements x the CTAs that share them), and the L2
            // atomic units make that the ~50 us floor of the deep-level launches, however the atomics are arranged (scattered from the
            // registers, or staged and issued as contiguous 128-byte rows: measured slower, the rows serialise on one L2 slice).
            // Here the block is laid out in shared memory (the rings are free once done_bar has fired) in the gradient's own order
            // [co][ci][27 taps] and stored, coalesced, to this CTA's slot of the scratch buffer; wgrad_band_sum_kernel adds the slots.
            float* const stage = reinterpret_cast<float*>(smem);
            if constexpr (R > 1) {
                for (int i = r; i < CO * RUN; i += 128) stage[i] = 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
#pragma unroll 1
            for (int dz = 0; dz < 3; ++dz) {
#pragma unroll 1
                for (int b = 0; b < R; ++b) {
                    const int dyi = a - b;   // = dy + 1
                    const bool ok = dyi >= 0 && dyi <= 2;
#pragma unroll 1
                    for (int dxc = 0; dxc < 3; ++dxc) {
                        const int tap = (dz * 3 + dyi) * 3 + dxc;
#pragma unroll 1
                        for (int c0 = 0; c0 < CO; c0 += 16) {
                            float v[16];
                            tmem_ld16(t_row + uint32_t(dz * N + (b * 3 + dxc) * CO + c0), v);
                            if (ok) {
                                // (lanes = consecutive ci: stride 27 floats, conflict-free); several (a, b) pairs share a tap when R > 1
                                float* const sp = &stage[c0 * RUN + cil * 27 + tap];
                                if constexpr (R == 1) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sp[j * RUN] = v[j];
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) atomicAdd(sp + j * RUN, v[j]);
                                }
                            }
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float4* const dst = reinterpret_cast<float4*>(p.scratch + size_t(blockIdx.x) * (CO * RUN));
            const float4* const src = reinterpret_cast<const float4*>(stage);
            for (int i = r; i < CO * RUN / 4; i += 128) dst[i] = src[i];
        } else {
            const int ci = gi * CIG + cil;
            const size_t nstride = size_t(P.w_mtot) * P.w_ktaps;
            const bool ci_ok = ci < P.t_creal;
#pragma unroll 1
            for (int dz = 0; dz < 3; ++dz) {
#pragma unroll 1
                for (int b = 0; b < R; ++b) {
                    const int dyi = a - b;   // = dy + 1
                    const bool ok = ci_ok && dyi >= 0 && dyi <= 2;
#pragma unroll 1
                    for (int dxc = 0; dxc < 3; ++dxc) {
                        const int tap = (dz * 3 + dyi) * 3 + dxc;
                        float* dwrow = P.dw + size_t(P.w_moff + ci) * P.w_ktaps + tap;
#pragma unroll 1
                        for (int c0 = 0; c0 < CO; c0 += 16) {
                            float v[16];
                            tmem_ld16(t_row + uint32_t(dz * N + (b * 3 + dxc) * CO + c0), v);
                            if (ok) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (go * CO + c0 + j < P.u_creal) atomicAdd(dwrow + size_t(P.w_noff + go * CO + c0 + j) * nstride, v[j]);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Adds the per-CTA blocks of partial-block mode into the gradient.  CTA index of the main kernel = rank * npairs + pair; block
// (pair, co) of this kernel owns the contiguous run [ci][27 taps] of output channel co of channel-group pair `pair`; blockIdx.y cuts
// the ranks into slices of 16 (fixed-order sums inside a slice; with more than one slice the slices meet in fp32 atomics -- 16x fewer
// than before -- otherwise the add is a plain read-modify-write and the layer's weight gradient is deterministic).
__global__ void wgrad_band_sum_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int npairs, int ngo, int nranks, int co_grp,
                                      int ci_grp, int t_creal, int u_creal, int w_mtot, int w_moff, int w_noff) {
    const int pair = blockIdx.x / co_grp, co = blockIdx.x % co_grp;
    const int gi = pair / ngo, go = pair % ngo;
    const int run_all = ci_grp * 27;
    const int ci_real = min(ci_grp, t_creal - gi * ci_grp);
    const int run = ci_real > 0 ? ci_real * 27 : 0;
    if (go * co_grp + co >= u_creal) return;
    const int r0 = blockIdx.y * 16, r1 = min(nranks, r0 + 16);
    const size_t slot = size_t(co_grp) * run_all;
    float* const dst = dw + (size_t(w_noff + go * co_grp + co) * w_mtot + w_moff + gi * ci_grp) * 27;
    for (int i = threadIdx.x; i < run; i += blockDim.x) {
        // all loads of the slice are issued before the first add (a dependent load-add chain made this kernel 15 - 34 us per launch)
        float v[16];
        const float* const src = scratch + size_t(pair) * slot + size_t(co) * run_all + i;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = r0 + k < r1 ? __ldcs(src + size_t(r0 + k) * npairs * slot) : 0.f;
        const float old = gridDim.y > 1 ? 0.f : dst[i];
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += v[k];
        if (gridDim.y > 1) atomicAdd(dst + i, acc);
        else dst[i] = old + acc;
    }
}

template <int NCG, int CO>
int launch_wband_t(const WBParams& wp, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_band_kernel<NCG, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_wgrad_band_kernel<NCG, CO><<<grid, kWThreads, smem, stream>>>(wp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace

unsigned int read_device_error_wband() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

bool conv_wgrad_band_eligible(const WgradProblem& P) {
    static const bool disabled = std::getenv("U3D_NO_WBAND") != nullptr || std::getenv("U3D_NO_HALO") != nullptr;
    if (disabled) return false;
    if (P.ntaps != 27 || P.tstride != 1 || P.w_ktaps != 27) return false;
    if (P.t_c % 16 || P.u_c % 16 || P.t_c < 16 || P.u_c < 16) return false;
    if (P.t_d != P.ld || P.t_h != P.lh || P.t_w != P.lw) return false;
    // small-channel layers need a big volume to fill the machine; wide layers bring (Cin/32)*(Cout/32) independent pairs
    const long long pairs = (P.t_c > 32 || P.u_c > 32) ? 1LL * (P.t_c / 16) * (P.u_c / 16) / 4 : 1;
    if (1LL * P.ld * P.lh * P.lw * pairs < 32768 || 1LL * P.ld * P.lh * P.lw < 1024 || pairs > 64) return false;
    for (int t = 0; t < 27; ++t)   // forward tap order (kz,ky,kx) with offsets k-1 and identity tap_ref
        if (P.taps[t].dz != t / 9 - 1 || P.taps[t].dy != (t / 3) % 3 - 1 || P.taps[t].dx != t % 3 - 1 || P.tap_ref[t] != t) return false;
    return true;
}

size_t conv_wgrad_band_scratch_bytes() { return size_t(device_sm_count()) * 32 * 32 * 27 * 4; }

int conv_wgrad_band_launch(const WgradProblem& P, cudaStream_t stream, float* partial_scratch, size_t partial_scratch_bytes) {
    WBParams wp;
    std::memset(&wp, 0, sizeof(wp));
    wp.P = P;
    // channel groups per CTA: 32 x 32 when both sides allow it, else 16 x 16 (also the native shapes of the 16/32-channel layers)
    const int tcg = (P.t_c % 32 == 0) ? 32 : 16, ucg = (P.u_c % 32 == 0) ? 32 : 16;
    const int ncg = tcg / 8, ncgy = ucg / 8;
    const int arows = 16 / ncg;
    const int R = std::min(ucg == 16 ? 3 : 1, arows - 2);
    const int ngi = P.t_c / tcg, ngo = P.u_c / ucg;
    // rows per tile: multiple of R, as many as fit next to the rings
    int TY = 0;
    for (int ty = R; ty <= 24; ty += R) {
        const int hya = std::max(ty + 2, ty - R + arows);
        const size_t need = size_t(kXSlots) * hya * ncg * kRunB + size_t(kYSlots) * ty * 3 * ncgy * kRunB + 1024;
        if (need <= 200 * 1024 && ty <= std::max(R, P.lh)) TY = ty;
    }
    if (TY == 0) { set_error("conv_wgrad_band_launch: tile does not fit in shared memory"); return 1; }
    {   // do not pay for rows past the volume: even out the tiles
        const int tiles = (P.lh + TY - 1) / TY;
        int even = (P.lh + tiles - 1) / tiles;
        even = (even + R - 1) / R * R;
        TY = std::min(TY, even);
    }
    wp.TY = TY;
    wp.HYA = std::max(TY + 2, TY - R + arows);
    wp.tiles_x = (P.lw + 31) / 32;
    wp.tiles_y = (P.lh + TY - 1) / TY;
    const int sms0 = device_sm_count();
    wp.ngo = ngo;
    wp.npairs = ngi * ngo;
    const int sms = std::max(1, sms0 / wp.npairs);    // CTAs per pair
    const int cols = wp.tiles_x * wp.tiles_y;
    int best_zc = 1;
    double best_eff = -1;
    for (int zc = 1; zc <= std::max(1, P.ld / 4); ++zc) {
        const int zl = (P.ld + zc - 1) / zc;
        const int zc_eff = (P.ld + zl - 1) / zl;
        const long long items = 1LL * cols * zc_eff;
        const long long waves = (items + sms - 1) / sms;
        const double eff = double(items) / double(waves * sms) * double(zl) / double(zl + 2);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_zc = zc_eff; }
    }
    wp.zlen = (P.ld + best_zc - 1) / best_zc;
    wp.zchunks = (P.ld + wp.zlen - 1) / wp.zlen;
    wp.total_items = cols * wp.zchunks;
    wp.x_slot_bytes = uint32_t(wp.HYA * ncg * kRunB + 127) & ~127u;
    wp.y_slot_bytes = uint32_t(TY * 3 * ncgy * kRunB + 127) & ~127u;
    wp.off_y = kXSlots * wp.x_slot_bytes;
    wp.off_bars = wp.off_y + kYSlots * wp.y_slot_bytes;
    {   // the epilogue stages the CTA's gradient block [ucg][tcg][27] fp32 over the (then idle) rings
        const uint32_t stage_bytes = uint32_t(ucg) * uint32_t(tcg) * 27u * 4u;
        if (wp.off_bars < stage_bytes) wp.off_bars = (stage_bytes + 127u) & ~127u;
    }
    const size_t smem = wp.off_bars + 8 * (2 * kXSlots + 2 * kYSlots + 1) + 16;
    if (smem > 227 * 1024) { set_error("conv_wgrad_band_launch: tile does not fit in shared memory"); return 1; }
    wp.cpp = std::max(1, std::min(wp.total_items, sms));
    const int grid = wp.cpp * wp.npairs;
    static const bool no_partial = std::getenv("U3D_WBAND_ATOMICS") != nullptr;
    const size_t slot_bytes = size_t(ucg) * tcg * 27 * 4;
    if (!no_partial && partial_scratch != nullptr && size_t(grid) * slot_bytes <= partial_scratch_bytes) wp.scratch = partial_scratch;
    int rc;
    if (ncg == 2 && ucg == 16) rc = launch_wband_t<2, 16>(wp, grid, smem, stream);
    else if (ncg == 2) rc = launch_wband_t<2, 32>(wp, grid, smem, stream);
    else if (ucg == 16) rc = launch_wband_t<4, 16>(wp, grid, smem, stream);
    else rc = launch_wband_t<4, 32>(wp, grid, smem, stream);
    if (rc || wp.scratch == nullptr) return rc;
    const int nranks = std::min(wp.cpp, wp.total_items);   // CTAs of rank >= total_items had no work and wrote nothing
    return wgrad_block_sum_launch(wp.scratch, P, wp.npairs, ngo, nranks, ucg, tcg, stream);
}

int wgrad_block_sum_launch(const float* scratch, const WgradProblem& P, int npairs, int ngo, int nranks, int co_grp, int ci_grp, cudaStream_t stream) {
    const dim3 sgrid(unsigned(npairs * co_grp), unsigned((nranks + 15) / 16));
    wgrad_band_sum_kernel<<<sgrid, 256, 0, stream>>>(scratch, P.dw, npairs, ngo, nranks, co_grp, ci_grp, P.t_creal, P.u_creal, P.w_mtot, P.w_moff,
                                                     P.w_noff);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
