// Row-stacked halo weight gradient for k3 s1 p1 convolutions with 16 or 32 input channels per source and <= 32 output
// channels on large volumes (the layers that dominate the backward pass).
//
//   dW[co][ci][dz,dy,dx] = sum_v x[v + (dz,dy,dx)][ci] * dy[v][co]
//
// The generic kernel (conv_wgrad.cu) re-gathers x from L2 once per tap: measured 9.7 GB through the crossbar for one
// 16->16 layer at 160x192x160, L2 at 53 % of peak, MMA thread starved.  Here a CTA loads the x tile plus halo ONCE and
// the dy tile once per output tile, into shared memory laid out
//     xs [hz][hy][cg][hx][8 ch]         dys[tz][ty][cgy][tx][8 ch]
// i.e. inside a (z,y) row the channel groups are separate runs of HX (TX) voxels.  For an MN-major UMMA operand
// (K = voxels along x, 8 voxels x 16 B = one core matrix, LBO = 128 B) the M chunks are then at ONE uniform stride
// SBO = HX*16 B: chunk m = (row hy + m / ncg, channel group m % ncg).  A single M=128 MMA therefore covers the dy = -1,0,+1
// taps (and 16/ncg - 3 junk rows that are never read back) for all input channels, and the remaining taps (dz,dx) are
// 9 start addresses.  Per output row of 32 voxels: 9 accumulators x 2 K-steps = 18 MMAs, no per-tap traffic at all.
// All 9 accumulators (9 x N fp32 columns) stay in TMEM for the CTA's whole life; one epilogue at the end adds them into
// the reference-layout gradient with fp32 atomics (148 CTAs x 27 x Cin x Cout adds).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kRIssuers = 3;     // one MMA-issuing warp per dz plane: a single thread's issue loop (~80-120 clk per MMA, measured)
                                 // is the bound, the hardware accepts one M=128 MMA per ~39 clk from any mix of warps
constexpr int kRThreads = 32 * (12 + kRIssuers);   // warps 0-3 epilogue, 4-11 producers, 12.. MMA issuers
constexpr int kRProducers = 256;

struct RParams {
    WgradProblem P;
    int tiles_x, tiles_y, tiles_z, total_tiles;
    int TX, TY, TZ, HX, HY, HZ;
    int ncg, ncgy;        // channel groups of x (per source) and of dy
    int n;                // padded Cout (16 or 32)
    int nbuf;
    uint32_t x_bytes, dy_bytes, buf_bytes, off_bars;
    int tmem_cols;
};

__global__ void __launch_bounds__(kRThreads, 1) conv_wgrad_rows_kernel(const __grid_constant__ RParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + p.off_bars;
    auto full_bar = [&](int b) { return bars + 8u * b; };
    auto empty_bar = [&](int b) { return bars + 8u * (2 + b); };
    const uint32_t done_bar = bars + 8u * 4;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * 5);

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(full_bar(b), kRProducers);
            mbar_init(empty_bar(b), kRIssuers);
        }
        mbar_init(done_bar, kRIssuers);
        fence_barrier_init();
    }
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const WgradProblem& P = p.P;
    const bool has_work = int(blockIdx.x) < p.total_tiles;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        const int D = P.t_d, H = P.t_h, W = P.t_w;
        const int ncg = p.ncg, ncgy = p.ncgy;
        const uint8_t* const xsrc = static_cast<const uint8_t*>(P.T) + P.t_coff * 2;
        const uint8_t* const ysrc = static_cast<const uint8_t*>(P.U) + P.u_coff * 2;
        const uint32_t xpitch = uint32_t(P.t_cp) * 2u, ypitch = uint32_t(P.u_cp) * 2u;
        uint32_t cnt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++cnt) {
            int rem = tile;
            const int tx = rem % p.tiles_x; rem /= p.tiles_x;
            const int ty = rem % p.tiles_y;
            const int tz = rem / p.tiles_y;
            const int x0 = tx * p.TX, y0 = ty * p.TY, z0 = tz * p.TZ;
            const int buf = cnt % p.nbuf;
            mbar_wait(empty_bar(buf), ((cnt / p.nbuf) & 1) ^ 1, 0xD00u | buf);
            const uint32_t xs = sbase + buf * p.buf_bytes;
            const uint32_t ds = xs + p.x_bytes;
            // x halo block: consecutive lanes = the channel groups of one voxel, then the next voxel along x
            const int xtotal = p.HZ * p.HY * p.HX * ncg;
#pragma unroll 4
            for (int idx = t; idx < xtotal; idx += kRProducers) {
                const int cg = idx % ncg;
                int q = idx / ncg;
                const int hx = q % p.HX; q /= p.HX;
                const int hy = q % p.HY;
                const int hz = q / p.HY;
                const int gx = x0 + hx - 1, gy = y0 + hy - 1, gz = z0 + hz - 1;
                const bool ok = (unsigned)gx < (unsigned)W && (unsigned)gy < (unsigned)H && (unsigned)gz < (unsigned)D;
                const uint8_t* src = ok ? xsrc + ((size_t(gz) * H + gy) * W + gx) * xpitch + cg * 16 : xsrc;
                cp_async16(xs + (uint32_t((hz * p.HY + hy) * ncg + cg) * p.HX + hx) * 16u, src, ok ? 16u : 0u);
            }
            // dy tile: zero outside the volume so ragged tiles contribute nothing
            const int ytotal = p.TZ * p.TY * p.TX * ncgy;
#pragma unroll 4
            for (int idx = t; idx < ytotal; idx += kRProducers) {
                const int cg = idx % ncgy;
                int q = idx / ncgy;
                const int lx = q % p.TX; q /= p.TX;
                const int ly = q % p.TY;
                const int lz = q / p.TY;
                const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
                const bool ok = gx < W && gy < H && gz < D;
                const uint8_t* src = ok ? ysrc + ((size_t(gz) * H + gy) * W + gx) * ypitch + cg * 16 : ysrc;
                cp_async16(ds + (uint32_t((lz * p.TY + ly) * ncgy + cg) * p.TX + lx) * 16u, src, ok ? 16u : 0u);
            }
            cp_async_mbar_arrive(full_bar(buf));
        }
        cp_async_wait<0>();
    } else if (warp >= 12) {
        // ===================================== MMA issuers ===================================
        const int wi = warp - 12;   // this warp owns the three accumulators of dz = wi - 1
        if (lane == 0 && has_work) {
            const int n = p.n, ncg = p.ncg, ncgy = p.ncgy;
            const uint32_t idesc = umma_idesc(128, n, 0, 0, 1, 1);          // both operands MN-major
            const uint32_t sbo_a = uint32_t(p.HX) * 16u, sbo_b = uint32_t(p.TX) * 16u;
            // offsets in 16-byte units; acc index = (dz+1)*3 + (dx+1)
            long long aoff[3];
            uint32_t dacc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int dz = wi - 1, dx = k - 1;
                aoff[k] = (long long)(dz * p.HY * ncg) * p.HX + dx;
                dacc[k] = tmem_base + uint32_t((wi * 3 + k) * n);
            }
            const uint64_t a_row_u = uint64_t(ncg) * p.HX;                   // one hy row
            const uint64_t a_plane_u = a_row_u * p.HY;                       // one hz plane
            const uint64_t b_row_u = uint64_t(ncgy) * p.TX;
            uint32_t cnt = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++cnt) {
                const int buf = cnt % p.nbuf;
                mbar_wait(full_bar(buf), (cnt / p.nbuf) & 1, 0xE00u | buf);
                fence_proxy_async();
                tc_fence_after();
                const uint32_t xs = sbase + buf * p.buf_bytes;
                // A base: row hy-1 of plane hz (chunk 0 = dy -1), channel group 0, voxel hx = 1 (+dx via aoff)
                const uint64_t a_tile = umma_smem_desc(xs + 16u, 128u, sbo_a);
                const uint64_t b_tile = umma_smem_desc(xs + p.x_bytes, 128u, sbo_b);
#pragma unroll 1
                for (int lz = 0; lz < p.TZ; ++lz) {
#pragma unroll 1
                    for (int ly = 0; ly < p.TY; ++ly) {
                        const uint64_t a_row = a_tile + uint64_t(lz + 1) * a_plane_u + uint64_t(ly) * a_row_u;
                        const uint64_t b_row = b_tile + uint64_t(lz * p.TY + ly) * b_row_u;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const uint64_t ad = a_row + uint64_t(aoff[k]) + uint64_t(ks * 16);
                                const uint64_t bd = b_row + uint64_t(ks * 16);
                                umma_f16(dacc[k], ad, bd, idesc, first ? 0u : 1u);
                            }
                            first = false;
                        }
                    }
                }
                umma_commit(empty_bar(buf));
            }
            umma_commit(done_bar);
        }
        __syncwarp();
    } else if (has_work) {
        // ===================================== epilogue (once) ================================
        const int r = threadIdx.x;
        mbar_wait(done_bar, 0, 0xF10u);
        tc_fence_after();
        const int ncg = p.ncg;
        const int chunk = r >> 3;
        const int dyi = chunk / ncg;
        const int ci = (chunk % ncg) * 8 + (r & 7);
        const bool rv = dyi < 3 && ci < P.t_creal;
        const size_t nstride = size_t(P.w_mtot) * P.w_ktaps;
        const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            const int dzi = k / 3, dxi = k % 3;
            const int tap = (dzi * 3 + dyi) * 3 + dxi;
            float* dwrow = P.dw + size_t(P.w_moff + ci) * P.w_ktaps + tap;
#pragma unroll 1
            for (int c0 = 0; c0 < p.n; c0 += 16) {
                float v[16];
                tmem_ld16(t_row + uint32_t(k * p.n + c0), v);
                if (rv) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < P.u_creal) atomicAdd(dwrow + size_t(P.w_noff + c0 + j) * nstride, v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

unsigned int read_device_error_rows() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

bool conv_wgrad_rows_eligible(const WgradProblem& P) {
    static const bool disabled = std::getenv("U3D_NO_HALO") != nullptr;
    if (disabled) return false;
    if (P.ntaps != 27 || P.tstride != 1 || P.w_ktaps != 27) return false;
    if (P.t_c != 16 && P.t_c != 32) return false;
    if (P.u_c != 16 && P.u_c != 32) return false;
    if (P.t_d != P.ld || P.t_h != P.lh || P.t_w != P.lw) return false;
    if (1LL * P.ld * P.lh * P.lw < 32768) return false;
    for (int t = 0; t < 27; ++t) {   // forward tap order (kz,ky,kx) with offsets k-1 and identity tap_ref
        if (P.taps[t].dz != t / 9 - 1 || P.taps[t].dy != (t / 3) % 3 - 1 || P.taps[t].dx != t % 3 - 1 || P.tap_ref[t] != t) return false;
    }
    return true;
}

int conv_wgrad_rows_launch(const WgradProblem& P, cudaStream_t stream) {
    RParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.P = P;
    rp.TX = 32; rp.TY = 8; rp.TZ = 4;
    rp.HX = rp.TX + 2; rp.HY = rp.TY + 2; rp.HZ = rp.TZ + 2;
    rp.ncg = P.t_c / 8;
    rp.ncgy = P.u_c / 8;
    rp.n = P.u_c;
    rp.tiles_x = (P.lw + rp.TX - 1) / rp.TX;
    rp.tiles_y = (P.lh + rp.TY - 1) / rp.TY;
    rp.tiles_z = (P.ld + rp.TZ - 1) / rp.TZ;
    rp.total_tiles = rp.tiles_x * rp.tiles_y * rp.tiles_z;
    // the M = 128 MMA reads 16 chunks = 16/ncg rows starting at row hy-1: up to 16/ncg - 3 rows past the halo.  Pad the x block
    // so those (never used) reads stay inside the buffer.
    const int extra_rows = 16 / rp.ncg;
    rp.x_bytes = uint32_t(((rp.HZ * rp.HY + extra_rows) * rp.ncg * rp.HX + 8) * 16);
    rp.x_bytes = (rp.x_bytes + 127u) & ~127u;
    rp.dy_bytes = uint32_t((rp.TZ * rp.TY * rp.ncgy * rp.TX) * 16);
    rp.dy_bytes = (rp.dy_bytes + 127u) & ~127u;
    rp.buf_bytes = rp.x_bytes + rp.dy_bytes;
    rp.nbuf = (size_t(2) * rp.buf_bytes + 1024 <= 220 * 1024) ? 2 : 1;
    rp.off_bars = rp.nbuf * rp.buf_bytes;
    const size_t smem = rp.off_bars + 8 * 5 + 16;
    if (smem > 227 * 1024) { set_error("conv_wgrad_rows_launch: tile does not fit in shared memory"); return 1; }
    int cols = 32;
    while (cols < 9 * rp.n) cols <<= 1;
    rp.tmem_cols = cols;
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = std::max(1, std::min(rp.total_tiles, device_sm_count()));
    conv_wgrad_rows_kernel<<<grid, kRThreads, smem, stream>>>(rp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Dispatcher used by the model and the op-level API: row-stacked halo kernel per eligible problem, generic kernel for the rest.
int conv_wgrad_dispatch(const std::vector<WgradProblem>& probs, const WgradLaunch& cfg, cudaStream_t stream, int* launches) {
    std::vector<WgradProblem> rest;
    int n = 0;
    for (const auto& P : probs) {
        if (conv_wgrad_band_eligible(P)) {
            if (conv_wgrad_band_launch(P, stream)) return 1;
            ++n;
        } else if (conv_wgrad_rows_eligible(P)) {
            if (conv_wgrad_rows_launch(P, stream)) return 1;
            ++n;
        } else
            rest.push_back(P);
    }
    if (!rest.empty()) {
        if (conv_wgrad_launch(rest, cfg, nullptr, stream)) return 1;
        ++n;
    }
    if (launches) *launches = n;
    return 0;
}

}  // namespace u3d
