// Halo-block 3x3x3 stride-1 convolution on tcgen05 / TMEM (sm_100a) for the large-volume, few-channel layers that carry
// most of the U-Net's FLOPs (levels 0-1: Cin <= 32, Cout <= 32).
//
// Idea: instead of gathering a fresh A tile from L2 for each of the 27 taps (conv_igemm.cu), a CTA loads the input
// voxels of an output tile PLUS a one-voxel halo exactly once into shared memory, in the channel-interleaved layout
//     smem[cg][p][8 ch]      p = (hz*HY + hy)*HX + hx  (flattened halo position), cg = channel group of 8
// which is the canonical SWIZZLE_NONE K-major UMMA operand layout for ANY run of 128 consecutive positions:
// 8 consecutive p x 16 B = one core matrix (SBO = 128 B), the next channel group is LBO = NP*16 B away.
// A tap (dz,dy,dx) is then just a different start address: the same block shifted by ((dz*HY+dy)*HX+dx)*16 bytes.
// M tiles are runs of 128 consecutive flattened positions between the first and last interior voxel; rows that fall
// on halo columns are computed and discarded (27 % of MMA rows at the default tile — the tensor pipe is not the bound,
// shared-memory operand bandwidth is), so there is NO per-tap global traffic, barrier or address arithmetic at all.
// The data gradient of a k3 s1 conv is the same kernel with the flipped/transposed weight pack (plan.cpp).
//
// CTA = 416 threads: warps 0-3 epilogue (TMEM -> bias/stats -> fp16 NDHWC), warps 4-11 producers (cp.async 16 B with
// zero fill, coalesced along x and channels), warp 12 issues tcgen05.mma.  Weights (27 taps) stay resident in smem.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kHIssuers = 2;     // MMA-issuing warps: issuer w owns accumulator w (even / odd M tiles)
constexpr int kHThreads = 32 * (12 + kHIssuers);
constexpr int kProducers = 256;
constexpr int kMaxHaloProb = 2;

struct HParams {
    ConvProblem probs[kMaxHaloProb];
    int nprob;
    int tiles_x, tiles_y, tiles_z, tiles_per_prob, total_tiles;
    int TX, TY, TZ, HX, HY, HZ;
    int NP;            // halo positions
    int NP_alloc;      // positions per channel-group plane incl. slack for the last M tile
    int ncg;           // channel groups of 8 (both sources)
    int ksteps;        // ncg / 2
    int mtiles;        // M tiles per CTA tile
    int p_first;       // flattened position of the first interior voxel
    int nbuf;          // A-block buffers (1 or 2)
    int n;             // N (padded Cout), identical for all problems
    int tmem_cols;
    uint32_t off_w, off_stats, off_bars;
    uint32_t a_buf_bytes, w_bytes;
    float* stats;      // [grid][2][n] or nullptr
    int epi;           // EPI_STORE16 | EPI_ACCUM16
};

template <int HALF, int BIT>
__device__ __forceinline__ void halve_step_h(float (&a)[16], float (&q)[16], int lane) {
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

template <int KSTEPS>
__global__ void __launch_bounds__(kHThreads, 1) conv_halo_kernel(const __grid_constant__ HParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase + p.off_w;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sbase + p.off_bars;
    // barriers: full[2], empty[2], tmem_full[2], tmem_empty[2], wfull
    auto full_bar = [&](int b) { return bars + 8u * b; };
    auto empty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto tfull_bar = [&](int a) { return bars + 8u * (4 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (6 + a); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * 9);

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(full_bar(b), kProducers);
            mbar_init(empty_bar(b), kHIssuers);
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 128);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 8 * p.n; i += kHThreads) sstats[i] = 0.f;
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        uint32_t cnt = 0;
        int loaded_prob = -1;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++cnt) {
            const int pi = tile / p.tiles_per_prob;
            const ConvProblem& P = p.probs[pi];
            int rem = tile - pi * p.tiles_per_prob;
            const int tx = rem % p.tiles_x; rem /= p.tiles_x;
            const int ty = rem % p.tiles_y;
            const int tz = rem / p.tiles_y;
            const int x0 = tx * p.TX - 1, y0 = ty * p.TY - 1, z0 = tz * p.TZ - 1;  // global coords of halo position (0,0,0)
            const int buf = cnt % p.nbuf;
            const uint32_t ph = (cnt / p.nbuf) & 1;
            mbar_wait(empty_bar(buf), ph ^ 1, 0x900u | buf);
            if (pi != loaded_prob) {
                // (re)load the resident weight pack; the empty wait above guarantees the MMAs that read the previous
                // problem's weights have completed when nbuf == 1; with nbuf == 2 problems never alternate mid-flight
                // because a launch with two problems forces nbuf = 1 (see launcher)
                const uint8_t* wsrc = static_cast<const uint8_t*>(P.wpack);
                for (uint32_t o = t * 16u; o < p.w_bytes; o += kProducers * 16u) cp_async16(sW + o, wsrc + o, 16u);
                loaded_prob = pi;
            }
            const uint32_t blk = sbase + buf * p.a_buf_bytes;
            const int D = P.in_d, H = P.in_h, W = P.in_w;
            const int ncg0 = P.c0p / 8, ncg = p.ncg;
            const uint8_t* const s0 = static_cast<const uint8_t*>(P.src0);
            const uint8_t* const s1 = static_cast<const uint8_t*>(P.src1);
            const uint32_t pitch0 = uint32_t(P.c0p) * 2u, pitch1 = uint32_t(P.c1p) * 2u;
            const int total = p.NP * ncg;
            // consecutive lanes: consecutive 16-byte chunks of one voxel, then the next voxel along x (coalesced)
#pragma unroll 4
            for (int idx = t; idx < total; idx += kProducers) {
                const int cg = idx % ncg;
                const int pos = idx / ncg;
                const int hx = pos % p.HX;
                const int q = pos / p.HX;
                const int hy = q % p.HY, hz = q / p.HY;
                const int gx = x0 + hx, gy = y0 + hy, gz = z0 + hz;
                const bool ok = (unsigned)gx < (unsigned)W && (unsigned)gy < (unsigned)H && (unsigned)gz < (unsigned)D;
                const size_t vox = (size_t(gz) * H + gy) * W + gx;
                const uint8_t* src;
                if (cg < ncg0) src = ok ? s0 + vox * pitch0 + cg * 16 : s0;
                else src = ok ? s1 + vox * pitch1 + (cg - ncg0) * 16 : s1;
                cp_async16(blk + (uint32_t(cg) * p.NP_alloc + pos) * 16u, src, ok ? 16u : 0u);
            }
            cp_async_mbar_arrive(full_bar(buf));
        }
        cp_async_wait<0>();
    } else if (warp >= 12) {
        // ===================================== MMA issuers ===================================
        const int wi = warp - 12;
        if (lane == 0) {
            // Lean issue loop (measured with tools/mma_bench.cu: ~45 clk per tcgen05.mma when the only per-MMA work is one
            // 64-bit add per descriptor; ~200-380 clk when descriptors are rebuilt and parameters re-read per MMA).
            // Everything is hoisted into registers; descriptors are (base + offset in 16-byte units).
            uint32_t cnt = 0, acc_cnt = 0;
            const uint32_t idesc = umma_idesc(128, p.n, 0, 0, 0, 0);
            const uint32_t lbo_a = uint32_t(p.NP_alloc) * 16u;
            const uint64_t a_kstep = uint64_t((2u * lbo_a) >> 4);
            const uint64_t b_base = umma_smem_desc(sW, uint32_t(p.n) * 16u, 128u);
            const uint64_t b_step = uint64_t((uint32_t(p.n) * 32u) >> 4);
            const int mtiles = p.mtiles, nbuf = p.nbuf, total_tiles = p.total_tiles, nstride = gridDim.x, ncols = p.n;
            const uint32_t a_buf_bytes = p.a_buf_bytes, p_first16 = uint32_t(p.p_first) * 16u;
            long long toff[27];   // tap offsets in 16-byte units (sign-extended: added to the 64-bit descriptor)
#pragma unroll
            for (int k = 0; k < 27; ++k) toff[k] = (p.probs[0].taps[k].dz * p.HY + p.probs[0].taps[k].dy) * p.HX + p.probs[0].taps[k].dx;
            for (int tile = blockIdx.x; tile < total_tiles; tile += nstride, ++cnt) {
                const int buf = cnt % nbuf;
                mbar_wait(full_bar(buf), (cnt / nbuf) & 1, 0xA00u | buf);
                fence_proxy_async();
                tc_fence_after();
                const uint64_t a_tile = umma_smem_desc(sbase + buf * a_buf_bytes + p_first16, lbo_a, 128u);
#pragma unroll 1
                for (int mt = 0; mt < mtiles; ++mt, ++acc_cnt) {
                    const int acc = acc_cnt & 1;
                    if (acc != wi) continue;
                    mbar_wait(tempty_bar(acc), ((acc_cnt >> 1) & 1) ^ 1, 0xB00u | acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + uint32_t(acc * ncols);
                    const uint64_t a_mt = a_tile + uint64_t(mt) * 128u;
                    uint64_t bd = b_base;
#pragma unroll
                    for (int k = 0; k < 27; ++k) {
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            const uint64_t ad = a_mt + uint64_t(toff[k]) + uint64_t(ks) * a_kstep;
                            if (k == 0 && ks == 0) umma_f16_first(d_tmem, ad, bd, idesc);
                            else umma_f16_acc(d_tmem, ad, bd, idesc);
                            bd += b_step;
                        }
                    }
                    umma_commit(tfull_bar(acc));
                }
                umma_commit(empty_bar(buf));
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        // The epilogue paces the kernel (MMA needs ~1200 clk per M tile), so it is kept minimal: per-thread register partial
        // sums for the norm statistics (one cross-lane butterfly per CTA at the very end instead of one per M tile), voxel
        // position advanced incrementally (no div/mod), bias in shared memory.
        const int r = threadIdx.x;
        uint32_t acc_cnt = 0;
        float ssum[32], ssq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ssum[j] = ssq[j] = 0.f;
        float* sbias = sstats + 8 * p.n;   // [n] after the statistics rows
        int bias_prob = -1;
        const int HX = p.HX, HY = p.HY, TX = p.TX, TY = p.TY, TZ = p.TZ, ncols = p.n;
        const int step_x = 128 % HX, step_y = (128 / HX) % HY, step_z = (128 / HX) / HY;
        const bool want_stats = p.stats != nullptr;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int pi = tile / p.tiles_per_prob;
            const ConvProblem& P = p.probs[pi];
            if (pi != bias_prob) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int j = r; j < ncols; j += 128) sbias[j] = (P.bias != nullptr && j < P.n_real) ? __ldg(P.bias + j) : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_prob = pi;
            }
            int rem = tile - pi * p.tiles_per_prob;
            const int tx = rem % p.tiles_x; rem /= p.tiles_x;
            const int ty = rem % p.tiles_y;
            const int tz = rem / p.tiles_y;
            const int x0 = tx * TX - 1, y0 = ty * TY - 1, z0 = tz * TZ - 1;
            const int W = P.in_w, H = P.in_h, D = P.in_d;
            uint8_t* const dst = static_cast<uint8_t*>(P.dst) + P.dst_coff * 2;
            const uint32_t dst_pitch = uint32_t(P.dst_cp) * 2u;
            const bool accum = p.epi == EPI_ACCUM16;
            // halo coordinates of this thread's row in M tile 0, advanced by 128 positions per M tile
            int pos0 = p.p_first + r;
            int hx = pos0 % HX;
            int q0 = pos0 / HX;
            int hy = q0 % HY, hz = q0 / HY;
#pragma unroll 1
            for (int mt = 0; mt < p.mtiles; ++mt, ++acc_cnt) {
                const int gx = x0 + hx, gy = y0 + hy, gz = z0 + hz;
                const bool rv = hx >= 1 && hx <= TX && hy >= 1 && hy <= TY && hz >= 1 && hz <= TZ && gx < W && gy < H && gz < D;
                const size_t vox = rv ? (size_t(gz) * H + gy) * W + gx : 0;
                hx += step_x; hy += step_y; hz += step_z;
                if (hx >= HX) { hx -= HX; ++hy; }
                if (hy >= HY) { hy -= HY; ++hz; }
                const int acc = acc_cnt & 1;
                mbar_wait(tfull_bar(acc), (acc_cnt >> 1) & 1, 0xC00u | acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(acc * ncols);
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    if (c0 < ncols) {
                        float v[16];
                        tmem_ld16(t_row + c0, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += sbias[c0 + j];
                        uint4* out = reinterpret_cast<uint4*>(dst + vox * dst_pitch + c0 * 2);
                        if (accum && rv) {
                            const uint4 o0 = out[0], o1 = out[1];
                            const uint32_t ow_[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 f = unpack2<false>(ow_[j]);
                                v[2 * j] += f.x;
                                v[2 * j + 1] += f.y;
                            }
                        }
                        if (rv) {
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
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
            }
        }
        if (want_stats) {
            // one butterfly per CTA: per-thread partials -> per-warp column sums -> fixed-order sum over the 4 warps
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                if (c0 < ncols) {
                    float a[16], qq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { a[j] = ssum[c0 + j]; qq[j] = ssq[c0 + j]; }
                    halve_step_h<8, 16>(a, qq, lane);
                    halve_step_h<4, 8>(a, qq, lane);
                    halve_step_h<2, 4>(a, qq, lane);
                    halve_step_h<1, 2>(a, qq, lane);
                    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                    qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
                    if ((lane & 1) == 0) {
                        const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        float* ws = sstats + warp * 2 * ncols;
                        ws[col] = a[0];
                        ws[ncols + col] = qq[0];
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = r; i < 2 * p.n; i += 128)
                p.stats[size_t(blockIdx.x) * 2 * p.n + i] = ((sstats[i] + sstats[2 * p.n + i]) + sstats[4 * p.n + i]) + sstats[6 * p.n + i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

unsigned int read_device_error_halo() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

// Eligibility: k3 s1 p1 problems (same extent in and out), K = 16 | 32 | 64 channels in 16-wide chunks, one N tile of <= 32,
// 16-bit NDHWC epilogues.  The caller packs the weights with kc = 16 (chunk order = [tap][k16]).
bool conv_halo_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    static const bool disabled = std::getenv("U3D_NO_HALO") != nullptr;
    if (disabled || probs.empty() || probs.size() > kMaxHaloProb || probs[0].banded) return false;
    if (cfg.kc != 16 || cfg.epi == EPI_PLANAR32) return false;
    for (const auto& P : probs) {
        if (P.ntaps != 27 || P.istride != 1 || P.ostep != 1 || P.ntiles != 1 || P.ntile > 32) return false;
        if (P.od != P.in_d || P.oh != P.in_h || P.ow != P.in_w) return false;
        if ((P.c0p + P.c1p) != 16 && (P.c0p + P.c1p) != 32 && (P.c0p + P.c1p) != 64) return false;
        if ((P.nch0 + P.nch1) * 16 != P.c0p + P.c1p || P.coff0 || P.coff1) return false;
        if (P.ntile != probs[0].ntile || P.c0p + P.c1p != probs[0].c0p + probs[0].c1p) return false;
        if (P.in_d != probs[0].in_d || P.in_h != probs[0].in_h || P.in_w != probs[0].in_w) return false;
        if (1LL * P.in_d * P.in_h * P.in_w < 32768) return false;
        if (std::memcmp(P.taps, probs[0].taps, sizeof(P.taps)) != 0) return false;
    }
    return true;
}

int conv_halo_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    HParams hp;
    std::memset(&hp, 0, sizeof(hp));
    hp.nprob = int(probs.size());
    for (int i = 0; i < hp.nprob; ++i) hp.probs[i] = probs[i];
    const ConvProblem& P0 = probs[0];
    // tile: 32x8x4 outputs for K <= 32 channels; K = 64 needs a smaller block to fit next to the 110 KB of resident weights
    const int ktot = P0.c0p + P0.c1p;
    hp.TX = 32; hp.TY = ktot <= 32 ? 8 : 4; hp.TZ = ktot <= 32 ? 4 : 2;
    hp.HX = hp.TX + 2; hp.HY = hp.TY + 2; hp.HZ = hp.TZ + 2;
    hp.NP = hp.HX * hp.HY * hp.HZ;
    hp.p_first = (hp.HY + 1) * hp.HX + 1;
    const int p_last = ((hp.TZ * hp.HY) + hp.TY) * hp.HX + hp.TX;
    hp.mtiles = (p_last - hp.p_first + 1 + 127) / 128;
    const int max_off = (hp.HY + 1) * hp.HX + 1;
    hp.NP_alloc = std::max(hp.NP, hp.p_first + hp.mtiles * 128 + max_off) + 8;
    hp.NP_alloc = (hp.NP_alloc + 7) / 8 * 8;
    hp.ncg = (P0.c0p + P0.c1p) / 8;
    hp.ksteps = hp.ncg / 2;
    hp.n = P0.ntile;
    hp.tiles_x = (P0.in_w + hp.TX - 1) / hp.TX;
    hp.tiles_y = (P0.in_h + hp.TY - 1) / hp.TY;
    hp.tiles_z = (P0.in_d + hp.TZ - 1) / hp.TZ;
    hp.tiles_per_prob = hp.tiles_x * hp.tiles_y * hp.tiles_z;
    hp.total_tiles = hp.tiles_per_prob * hp.nprob;
    hp.a_buf_bytes = uint32_t(hp.ncg) * hp.NP_alloc * 16u;
    hp.w_bytes = uint32_t(27) * hp.ksteps * hp.n * 32u;
    hp.nbuf = (hp.nprob == 1 && size_t(2) * hp.a_buf_bytes + hp.w_bytes + 9 * hp.n * 4 + 1024 <= 220 * 1024) ? 2 : 1;
    hp.off_w = hp.nbuf * hp.a_buf_bytes;
    hp.off_stats = hp.off_w + hp.w_bytes;
    hp.off_bars = uint32_t((hp.off_stats + 9 * hp.n * 4 + 15) & ~15u);
    const size_t smem = hp.off_bars + 8 * 9 + 16;
    if (smem > 227 * 1024) { set_error("conv_halo_launch: tile does not fit in shared memory"); return 1; }
    int cols = 32;
    while (cols < 2 * hp.n) cols <<= 1;
    hp.tmem_cols = cols;
    hp.epi = cfg.epi;
    hp.stats = (cfg.epi == EPI_STORE16 && probs.size() == 1) ? cfg.stats_partials : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = std::max(1, std::min(hp.total_tiles, device_sm_count()));
    if (cfg.stats_grid_out) *cfg.stats_grid_out = grid;
    if (hp.ksteps == 1) conv_halo_kernel<1><<<grid, kHThreads, smem, stream>>>(hp);
    else if (hp.ksteps == 2) conv_halo_kernel<2><<<grid, kHThreads, smem, stream>>>(hp);
    else conv_halo_kernel<4><<<grid, kHThreads, smem, stream>>>(hp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Single entry point used by the model and the op-level API: halo kernel when the problem set qualifies, gather kernel otherwise.
int conv_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    if (conv_band_eligible(probs, cfg)) return conv_band_launch(probs, cfg, stream);
    if (!probs.empty() && probs[0].banded) { set_error("conv_launch: banded weight pack but the problem is not eligible for conv_band"); return 1; }
    if (conv_s2_eligible(probs, cfg)) return conv_s2_launch(probs, cfg, stream);
    if (conv_halo_eligible(probs, cfg)) return conv_halo_launch(probs, cfg, stream);
    if (conv_tma_eligible(probs, cfg)) return conv_tma_launch(probs, cfg, stream);
    return conv_igemm_launch(probs, cfg, nullptr, stream);
}

int conv_kernel_kind(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    if (conv_band_eligible(probs, cfg)) return 5;
    if (conv_s2_eligible(probs, cfg) || conv_halo_eligible(probs, cfg)) return 2;
    return conv_tma_eligible(probs, cfg) ? 4 : 0;
}

// Planner hint: a k3 s1 layer with K <= 32 and N <= 32 channels on a big volume is planned with 16-wide K chunks so that
// conv_halo_eligible() accepts it.
bool conv_halo_wants_kc16(int ks, int stride, int transposed, int k_channels_padded, int n_channels_padded, long long voxels) {
    return !transposed && ks == 3 && stride == 1 && (k_channels_padded == 16 || k_channels_padded == 32 || k_channels_padded == 64) &&
           n_channels_padded <= 32 && voxels >= 32768 &&
           std::getenv("U3D_NO_HALO") == nullptr;
}

}  // namespace u3d
