// Halo-resident 3x3x3 STRIDE-2 convolution (forward of the down-sampling layer of level 0->1: 16 input channels on a large
// volume) on tcgen05 / TMEM.
//
// The TMA kernel fetches this layer as 27 strided tensor-map boxes of 32-byte elements per tile (measured 172 us for the
// 16->32 layer at 160x192x160: the copy engine is element-rate bound).  Here the input z-planes of a tile are loaded ONCE into
// shared memory, split into the four (y, x) parity planes:   xs[cg][hy&1][hx&1][(hy>>1)*HQ + (hx>>1)][8 ch]
// Input voxel of output (oy, ox) for tap (ky, kx) is halo position (2*oy + ky, 2*ox + kx), i.e. parity plane (ky&1, kx&1) at
// row (oy + (ky>>1))*HQ + ox + (kx>>1): for a fixed tap the rows of consecutive outputs are consecutive, so every tap is a
// start address of a canonical SWIZZLE_NONE K-major operand (as in conv_band.cu) and the 27 taps cost no global traffic.
// M row = one output voxel (p = oy*HQ + ox; rows with ox == OTX are discarded), N = Cout.  The CTA marches along z: output
// plane j needs input planes 2j, 2j+1, 2j+2 of its chunk (ring of 5 slots).
//
// CTA = 416 threads: warps 0-3 epilogue, 4-11 producers (cp.async.ca, zero fill = padding), warp 12 MMA issuer.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kSThreads = 32 * 13;
constexpr int kSProducers = 256;
constexpr int kSMaxSlots = 8;   // ring of input planes: 3 in use per output plane, the rest in flight (the kernel is DRAM-latency bound)

struct SParams {
    ConvProblem P;
    int OTX, OTY, HQ, ROWS, HX, HY;   // output tile, smem row pitch, rows per parity plane, input halo extent per plane
    int tiles_x, tiles_y, zchunks, zlen, total_items;
    int ncg, n;                        // input channel groups of 8, padded Cout
    int nslots;
    uint32_t slot_bytes, w_bytes, off_w, off_stats, off_bars;
    float* stats;
};

template <int HALF, int BIT>
__device__ __forceinline__ void halve_step_s(float (&a)[16], float (&q)[16], int lane) {
    const bool hi = (lane & BIT) != 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        const float sa = hi ? a[j] : a[j + HALF];
        const float ka = hi ? a[j + HALF] : a[j];
        a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, BIT);
        const float sq = hi ? q[j] : q[j + HALF];
        const float kq = hi ? q[j + HALF] : q[j];
        q[j] = kq + __shfl_xor_sync(0xffffffffu, sq, BIT);
    }
}

template <int KS, int NCH>   // K chunks of 16 input channels, output column chunks of 16
__global__ void __launch_bounds__(kSThreads, 1) conv_s2_kernel(const __grid_constant__ SParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int N = NCH * 16;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase + p.off_w;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sbase + p.off_bars;
    const uint32_t kSSlots = uint32_t(p.nslots);
    auto full_bar = [&](uint32_t s) { return bars + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bars + 8u * (kSMaxSlots + s); };
    auto tfull_bar = [&](uint32_t a) { return bars + 8u * (2 * kSMaxSlots + a); };
    auto tempty_bar = [&](uint32_t a) { return bars + 8u * (2 * kSMaxSlots + 2 + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * kSMaxSlots + 4);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * kSMaxSlots + 5));
    constexpr uint32_t TCOLS = 2 * N < 32 ? 32 : 2 * N;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kSSlots; ++s) { mbar_init(full_bar(s), kSProducers); mbar_init(empty_bar(s), 1); }
        for (uint32_t a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        mbar_init(wfull_bar, kSProducers);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 9 * N; i += kSThreads) sstats[i] = 0.f;
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), TCOLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const ConvProblem& P = p.P;
    const int D = P.in_d, H = P.in_h, W = P.in_w;          // input extent
    const int OD = P.od;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        {
            const uint8_t* wsrc = static_cast<const uint8_t*>(P.wpack);
            for (uint32_t o = t * 16u; o < p.w_bytes; o += kSProducers * 16u) cp_async16(sW + o, wsrc + o, 16u);
            cp_async_mbar_arrive(wfull_bar);
        }
        const int ncg = p.ncg;
        const uint8_t* const s0 = static_cast<const uint8_t*>(P.src0);
        const uint32_t pitch0 = uint32_t(P.c0p) * 2u;
        const int HX = p.HX, HY = p.HY, HQ = p.HQ, ROWS = p.ROWS;
        const int per_plane = HY * HX * ncg;
        const uint32_t inv_hx = (1u << 20) / uint32_t(HX) + 1u;
        const int cg_shift = ncg == 2 ? 1 : 2;
        uint32_t cnt = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int x0 = 2 * tx * p.OTX - 1, y0 = 2 * ty * p.OTY - 1;     // input coordinate of halo position (0, 0)
            const int z0 = zc * p.zlen, z1 = min(OD, z0 + p.zlen);
            for (int gz = 2 * z0 - 1; gz <= 2 * (z1 - 1) + 1; ++gz, ++cnt) {
                const uint32_t slot = cnt % kSSlots;
                mbar_wait(empty_bar(slot), ((cnt / kSSlots) & 1) ^ 1, 0x4100u | slot);
                const uint32_t blk = sbase + slot * p.slot_bytes;
                const bool zok = (unsigned)gz < (unsigned)D;
                const long long vox0 = ((long long)(zok ? gz : 0) * H + y0) * W + x0;
                const uint8_t* const p0 = s0 + vox0 * (long long)pitch0;
#pragma unroll 2
                for (int idx = t; idx < per_plane; idx += kSProducers) {
                    const int cg = idx & (ncg - 1);
                    const uint32_t pos = uint32_t(idx) >> cg_shift;
                    const int hy = int((pos * inv_hx) >> 20);
                    const int hx = int(pos) - hy * HX;
                    const bool ok = zok && (unsigned)(x0 + hx) < (unsigned)W && (unsigned)(y0 + hy) < (unsigned)H;
                    const uint8_t* src = ok ? p0 + (long long)(hy * W + hx) * pitch0 + cg * 16 : s0;
                    const int plane = (cg * 2 + (hy & 1)) * 2 + (hx & 1);
                    cp_async16_ca(blk + uint32_t(plane * ROWS + (hy >> 1) * HQ + (hx >> 1)) * 16u, src, ok ? 16u : 0u);
                }
                cp_async_mbar_arrive(full_bar(slot));
            }
        }
        cp_async_wait<0>();
    } else if (warp == 12) {
        // ===================================== MMA issuer ====================================
        // the whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane issues
        {
            const int HQ = p.HQ, ROWS = p.ROWS;
            const uint32_t idesc = umma_idesc(128, N, 0, 0, 0, 0);
            const uint32_t lbo_a = uint32_t(4 * ROWS) * 16u;                 // next channel group of 8
            const uint64_t a_ks_u = uint64_t((2u * lbo_a) >> 4);
            const uint64_t b_base = umma_smem_desc(sW, uint32_t(N) * 16u, 128u);
            const uint64_t b_step = uint64_t((uint32_t(N) * 32u) >> 4);      // one (tap, K chunk) slice
            long long aoff[9];   // (ky, kx) -> parity plane + row offset, 16-byte units
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int ky = k / 3, kx = k % 3;
                aoff[k] = (long long)(((ky & 1) * 2 + (kx & 1)) * ROWS + (ky >> 1) * HQ + (kx >> 1));
            }
            mbar_wait(wfull_bar, 0, 0x4200u);
            fence_proxy_async();
            uint32_t cnt = 0, acc_cnt = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int zc = item % p.zchunks;
                const int z0 = zc * p.zlen, z1 = min(OD, z0 + p.zlen);
                const int nz = z1 - z0;
                mbar_wait(full_bar(cnt % kSSlots), (cnt / kSSlots) & 1, 0x4300u);
#pragma unroll 1
                for (int j = 0; j < nz; ++j, ++acc_cnt) {
                    const uint32_t c0 = cnt + 2 * j;
                    mbar_wait(full_bar((c0 + 1) % kSSlots), ((c0 + 1) / kSSlots) & 1, 0x4301u);
                    mbar_wait(full_bar((c0 + 2) % kSSlots), ((c0 + 2) / kSSlots) & 1, 0x4302u);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint32_t acc = acc_cnt & 1u;
                    mbar_wait(tempty_bar(acc), ((acc_cnt >> 1) & 1) ^ 1, 0x4400u | acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * uint32_t(N);
                    uint64_t bd = b_base;
                    if (elect_one()) {
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz) {
                        const uint64_t a_pl = umma_smem_desc(sbase + ((c0 + kz) % kSSlots) * p.slot_bytes, lbo_a, 128u);
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint64_t ad = a_pl + uint64_t(aoff[k]) + uint64_t(ks) * a_ks_u;
                                if (kz == 0 && k == 0 && ks == 0) umma_f16_first(d_tmem, ad, bd, idesc);
                                else umma_f16_acc(d_tmem, ad, bd, idesc);
                                bd += b_step;
                            }
                        }
                    }
                    umma_commit(tfull_bar(acc));
                    umma_commit(empty_bar(c0 % kSSlots));          // planes 2j and 2j+1 are not needed by the next output plane
                    umma_commit(empty_bar((c0 + 1) % kSSlots));
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(empty_bar((cnt + 2 * nz) % kSSlots));  // the last plane of the chunk
                __syncwarp();
                cnt += uint32_t(2 * nz + 1);
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        const int r = threadIdx.x;
        const int HQ = p.HQ;
        uint32_t acc_cnt = 0;
        float ssum[N], ssq[N];
#pragma unroll
        for (int j = 0; j < N; ++j) ssum[j] = ssq[j] = 0.f;
        float* sbias = sstats + 8 * N;
        for (int j = r; j < N; j += 128) sbias[j] = (P.bias != nullptr && j < P.n_real) ? __ldg(P.bias + j) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool want_stats = p.stats != nullptr;
        const int loy = r / HQ, lox = r % HQ;
        const bool row_in_tile = loy < p.OTY && lox < p.OTX;
        uint8_t* const dst = static_cast<uint8_t*>(P.dst) + P.dst_coff * 2;
        const uint32_t dst_pitch = uint32_t(P.dst_cp) * 2u;
        const int OH = P.oh, OW = P.ow;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int ox = tx * p.OTX + lox, oy = ty * p.OTY + loy;
            const bool rv = row_in_tile && ox < OW && oy < OH;
            const int z0 = zc * p.zlen, z1 = min(OD, z0 + p.zlen);
#pragma unroll 1
            for (int oz = z0; oz < z1; ++oz, ++acc_cnt) {
                const size_t vox = (size_t(oz) * OH + oy) * OW + ox;
                const uint32_t acc = acc_cnt & 1u;
                mbar_wait(tfull_bar(acc), (acc_cnt >> 1) & 1, 0x4500u | acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + acc * uint32_t(N);
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 16) {
                    float v[16];
                    tmem_ld16(t_row + uint32_t(c0), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += sbias[c0 + j];
                    if (rv) {
                        uint4* out = reinterpret_cast<uint4*>(dst + vox * dst_pitch + c0 * 2);
                        uint4 q0v, q1v;
                        q0v.x = pack2<false>(v[0], v[1]); q0v.y = pack2<false>(v[2], v[3]);
                        q0v.z = pack2<false>(v[4], v[5]); q0v.w = pack2<false>(v[6], v[7]);
                        q1v.x = pack2<false>(v[8], v[9]); q1v.y = pack2<false>(v[10], v[11]);
                        q1v.z = pack2<false>(v[12], v[13]); q1v.w = pack2<false>(v[14], v[15]);
                        out[0] = q0v;
                        out[1] = q1v;
                        if (want_stats) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                ssum[c0 + j] += v[j];
                                ssq[c0 + j] = fmaf(v[j], v[j], ssq[c0 + j]);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
            }
        }
        if (want_stats) {
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                float a[16], qq[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { a[j] = ssum[c0 + j]; qq[j] = ssq[c0 + j]; }
                halve_step_s<8, 16>(a, qq, lane);
                halve_step_s<4, 8>(a, qq, lane);
                halve_step_s<2, 4>(a, qq, lane);
                halve_step_s<1, 2>(a, qq, lane);
                a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
                if ((lane & 1) == 0) {
                    const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    float* ws = sstats + warp * 2 * N;
                    ws[col] = a[0];
                    ws[N + col] = qq[0];
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = r; i < 2 * N; i += 128)
                p.stats[size_t(blockIdx.x) * 2 * N + i] = ((sstats[i] + sstats[2 * N + i]) + sstats[4 * N + i]) + sstats[6 * N + i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, TCOLS);
}

template <int KS, int NCH>
int launch_s2_t(const SParams& sp, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_s2_kernel<KS, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_s2_kernel<KS, NCH><<<grid, kSThreads, smem, stream>>>(sp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace

unsigned int read_device_error_s2() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

// planner hook: Conv3d k3 s2 forward, one source of 16|32 padded channels, <= 64 padded output channels, large output volume
bool conv_s2_wants_kc16(int ks, int stride, int transposed, int cin_padded, int n_sources, int cout_padded, long long out_voxels) {
    static const bool disabled = std::getenv("U3D_NO_S2") != nullptr || std::getenv("U3D_NO_HALO") != nullptr;
    // 16 padded input channels only: with 32 the resident weights (27 x 2 x Cout x 32 B) do not fit next to the 5-plane ring
    return !disabled && !transposed && ks == 3 && stride == 2 && n_sources == 1 && cin_padded == 16 &&
           (cout_padded == 16 || cout_padded == 32 || cout_padded == 64) && out_voxels >= 16384;
}

bool conv_s2_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    static const bool disabled = std::getenv("U3D_NO_S2") != nullptr || std::getenv("U3D_NO_HALO") != nullptr;
    if (disabled || probs.size() != 1) return false;
    const ConvProblem& P = probs[0];
    if (cfg.kc != 16 || cfg.epi != EPI_STORE16 || cfg.a_bf16 || cfg.b_bf16) return false;
    if (P.ntaps != 27 || P.istride != 2 || P.ostep != 1 || P.ntiles != 1 || P.c1p != 0 || P.nch1 != 0 || P.coff0 || P.shuffle_cp || P.banded) return false;
    if (P.c0p != 16 || P.nch0 != 1) return false;
    if (P.ntile != 16 && P.ntile != 32 && P.ntile != 64) return false;
    if (1LL * P.od * P.oh * P.ow < 16384 || P.dst_cp % 8 || P.dst_coff % 8) return false;
    for (int t = 0; t < 27; ++t)   // forward tap order (kz,ky,kx), offsets k-1
        if (P.taps[t].dz != t / 9 - 1 || P.taps[t].dy != (t / 3) % 3 - 1 || P.taps[t].dx != t % 3 - 1) return false;
    return true;
}

int conv_s2_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    SParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.P = probs[0];
    const ConvProblem& P = sp.P;
    sp.ncg = P.c0p / 8;
    sp.n = P.ntile;
    // output tile OTX x OTY with OTY*(OTX+1) <= 128 rows: fewest tiles
    long long best = -1;
    for (int otx = 8; otx <= 63; ++otx) {
        const int hq = otx + 1;
        const int oty_max = 128 / hq;
        if (oty_max < 1) break;
        const int tiles_y0 = (P.oh + oty_max - 1) / oty_max;
        const int oty = (P.oh + tiles_y0 - 1) / tiles_y0;
        const long long tiles = 1LL * ((P.ow + otx - 1) / otx) * tiles_y0;
        if (best < 0 || tiles < best) { best = tiles; sp.OTX = otx; sp.OTY = oty; }
    }
    sp.HQ = sp.OTX + 1;
    sp.HX = 2 * sp.OTX + 1;
    sp.HY = 2 * sp.OTY + 1;
    sp.ROWS = (128 + sp.HQ + 2) | 1;
    sp.tiles_x = (P.ow + sp.OTX - 1) / sp.OTX;
    sp.tiles_y = (P.oh + sp.OTY - 1) / sp.OTY;
    const int sms = device_sm_count();
    const int cols = sp.tiles_x * sp.tiles_y;
    int best_zc = 1;
    double best_eff = -1;
    for (int zc = 1; zc <= std::max(1, P.od / 2); ++zc) {
        const int zl = (P.od + zc - 1) / zc;
        const int zc_eff = (P.od + zl - 1) / zl;
        const long long items = 1LL * cols * zc_eff;
        const long long waves = (items + sms - 1) / sms;
        const double eff = double(items) / double(waves * sms) * double(2 * zl) / double(2 * zl + 1);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_zc = zc_eff; }
    }
    sp.zlen = (P.od + best_zc - 1) / best_zc;
    sp.zchunks = (P.od + sp.zlen - 1) / sp.zlen;
    sp.total_items = cols * sp.zchunks;
    sp.slot_bytes = uint32_t(sp.ncg * 4 * sp.ROWS * 16);
    const int KS = P.c0p / 16;
    sp.w_bytes = uint32_t(27 * KS * sp.n * 32);
    sp.nslots = int(std::min<size_t>(kSMaxSlots, (size_t(222) * 1024 - sp.w_bytes - 9 * sp.n * 4) / sp.slot_bytes));
    if (sp.nslots < 4) { set_error("conv_s2_launch: plane ring does not fit in shared memory"); return 1; }
    sp.off_w = uint32_t(sp.nslots) * sp.slot_bytes;
    sp.off_stats = sp.off_w + sp.w_bytes;
    sp.off_bars = uint32_t((sp.off_stats + 9 * sp.n * 4 + 15) & ~15u);
    const size_t smem = sp.off_bars + 8 * (2 * kSMaxSlots + 5) + 16;
    if (smem > 227 * 1024) { set_error("conv_s2_launch: tile does not fit in shared memory"); return 1; }
    sp.stats = cfg.stats_partials;
    const int grid = std::max(1, std::min(sp.total_items, sms));
    if (cfg.stats_grid_out) *cfg.stats_grid_out = grid;
    const int nch = sp.n / 16;
    if (KS == 1 && nch == 1) return launch_s2_t<1, 1>(sp, grid, smem, stream);
    if (KS == 1 && nch == 2) return launch_s2_t<1, 2>(sp, grid, smem, stream);
    if (KS == 1) return launch_s2_t<1, 4>(sp, grid, smem, stream);
    if (nch == 1) return launch_s2_t<2, 1>(sp, grid, smem, stream);
    if (nch == 2) return launch_s2_t<2, 2>(sp, grid, smem, stream);
    return launch_s2_t<2, 4>(sp, grid, smem, stream);
}

}  // namespace u3d
