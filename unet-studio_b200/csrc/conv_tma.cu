// TMA-fed implicit-GEMM convolution on tcgen05 / TMEM (sm_100a).
//
// Same problem form as conv_igemm.cu (u3d.h: taps x channel chunks gathered from one or two NDHWC sources), but the A
// operand is no longer gathered by 128 threads issuing 16-byte cp.async copies (measured: ~1500 clk of producer
// instruction chain per 16 KB stage, the bound of every layer of levels >= 2).  Instead an M tile is a BOX of output
// voxels (bx x by x bz <= 128) and one tap of one K chunk is ONE 4-D tensor-map copy (cp.async.bulk.tensor, SASS
// UTMALDG): coordinates (channel, x*istride + dx, y*istride + dy, z*istride + dz), element strides (1, s, s, s) for the
// stride-2 layers, out-of-bounds elements zero-filled by the copy engine (= the conv padding).  The box lands in shared
// memory as dense rows of KC*2 bytes in the hardware swizzle (128B / 64B / 32B for KC = 64 / 32 / 16 channels), which is
// exactly the K-major swizzled UMMA operand layout, so the MMA thread consumes it without any thread ever touching it.
// Weights keep the SWIZZLE_NONE blobs of layout.cu and arrive by 1-D bulk copy on the same mbarrier.
//
// CTA = 192 threads: warps 0-3 epilogue (TMEM lane quarter = warp), warp 4 = copy issuer, warp 5 = MMA issuer.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kTThreads = 192;
constexpr int kTMaxProb = 8;

struct TBox {
    int bx, by, bz;        // box of output voxels per M tile
    int tiles_x, tiles_y;  // tiles along x, y (z follows from mtiles)
    int rows;              // rows the copy engine writes: pitch*by*bz
    int pitch;             // row pitch along x: bx, or bx + 2 in x-halo mode
    int xhalo;             // 1: k3 s1 problem, ONE box with a one-voxel x halo per (dz,dy); the three dx taps are start-row shifts of it
    int min_dx;            // smallest dx of a tap triple (x origin of the halo box)
};

struct TParams {
    ConvProblem probs[kTMaxProb];
    TBox box[kTMaxProb];
    int nprob;
    int total_items;
    int stages;
    int ntile_max, ntot_max, tmem_cols;
    uint32_t a_stage_bytes, b_stage_bytes, off_b, off_stats, off_bars;
    float* stats;
    int ksplit;            // > 1: deterministic split-K -- work item (mt, nt, ks) reduces a contiguous range of the K units and writes its fp32
    float* split_scratch;  // accumulator to split_scratch[ks][lattice voxel][ntot]; conv_splitk_finalize_kernel sums the slices in order
};

struct alignas(64) TMaps {
    CUtensorMap m[2 * kTMaxProb];   // [problem][source]
};

struct TItem {
    int pi, mt, nt, ks;
};

__device__ __forceinline__ TItem decode_titem(const TParams& p, int item) {
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < p.nprob; ++i)
        if (item >= p.probs[i].item_base) pi = i;
    int local = item - p.probs[pi].item_base;
    const int ks = local % p.ksplit;
    local /= p.ksplit;
    const int ntiles = p.probs[pi].ntiles;
    return TItem{pi, local / ntiles, local % ntiles, ks};
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// K-major swizzled operand: rows of SWZ bytes, 8-row groups SBO = 8*SWZ bytes apart (tile base 1024-byte aligned)
template <int SWZ>
__device__ __forceinline__ uint64_t umma_desc_swz(uint32_t addr) {
    constexpr uint64_t layout = SWZ == 128 ? 2 : SWZ == 64 ? 4 : 6;
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t(1) << 16;                          // LBO: unused for swizzled K-major
    d |= uint64_t(((8u * SWZ) >> 4) & 0x3FFF) << 32;  // SBO
    d |= uint64_t(1) << 46;
    d |= layout << 61;
    return d;
}

template <int HALF, int BIT>
__device__ __forceinline__ void halve_step_t(float (&a)[16], float (&q)[16], int lane) {
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

template <int EPI, int KC>
__global__ void __launch_bounds__(kTThreads, 1) conv_tma_kernel(const __grid_constant__ TParams p, const __grid_constant__ TMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int SWZ = KC * 2;
    constexpr int SPG = 64 / KC;                       // K chunks (copies) per pipeline stage
    constexpr uint32_t A_SUB = 128u * KC * 2u;         // one chunk of A: 128 rows
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t sA = smem_u32(smem);
    const uint32_t sB = sA + p.off_b;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sA + p.off_bars;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * S + 4));

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 9 * p.ntot_max; i += kTThreads) sstats[i] = 0.f;   // [4 warps][2][ntot] statistics + [ntot] bias
    if (warp == 5) {
        tmem_alloc(smem_u32(tmem_ptr_smem), p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 4) {
        // ===================================== copy issuer ===================================
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const TItem w = decode_titem(p, item);
                const ConvProblem& P = p.probs[w.pi];
                const TBox bx = p.box[w.pi];
                const int txi = w.mt % bx.tiles_x;
                const int rest = w.mt / bx.tiles_x;
                const int tyi = rest % bx.tiles_y, tzi = rest / bx.tiles_y;
                const int ix0 = txi * bx.bx * P.istride, iy0 = tyi * bx.by * P.istride, iz0 = tzi * bx.bz * P.istride;
                const int nch0 = P.nch0, nch = P.nch0 + P.nch1, ntaps = P.ntaps;
                const int nsteps = ntaps * nch;
                const uint32_t a_bytes = uint32_t(bx.rows) * KC * 2u;
                const uint32_t b_bytes = uint32_t(P.ntile) * KC * 2u;
                const uint8_t* wbase = static_cast<const uint8_t*>(P.wpack);
                const CUtensorMap* m0 = &maps.m[2 * w.pi];
                const CUtensorMap* m1 = &maps.m[2 * w.pi + 1];
                const int coff0 = P.coff0, coff1 = P.coff1, ntiles = P.ntiles;
                int tap = 0, ch = 0;
                if (KC == 64 && bx.xhalo) {
                    // x-halo mode: per (dz,dy) pair and K chunk one box of (bx+2) x by x bz voxels + the weight slices of its 3 dx taps
                    const int U = (ntaps / 3) * nch;
                    const int u0 = w.ks * U / p.ksplit, u1 = (w.ks + 1) * U / p.ksplit;
#pragma unroll 1
                    for (int u = u0; u < u1; ++u) {
                        const int tp3 = (u / nch) * 3, c = u % nch;
                        const ConvTap tp = P.taps[tp3];
                        {
                            mbar_wait(empty_bar(stage), phase ^ 1, 0x1100u | stage);
                            mbar_arrive_expect_tx(full_bar(stage), a_bytes + 3u * b_bytes);
                            const bool first = c < nch0;
                            const int c0 = first ? coff0 + c * KC : coff1 + (c - nch0) * KC;
                            tma_load_4d(sA + stage * p.a_stage_bytes, first ? m0 : m1, full_bar(stage), c0, ix0 + bx.min_dx, iy0 + tp.dy, iz0 + tp.dz);
                            const uint32_t bdst = sB + stage * p.b_stage_bytes;
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                bulk_g2s(bdst + j * b_bytes, wbase + ((size_t(tp3 + j) * nch + c) * ntiles + w.nt) * b_bytes, b_bytes, full_bar(stage));
                            if (++stage == S) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                const int g0 = p.ksplit > 1 ? w.ks * nsteps / p.ksplit : 0;
                const int g1 = p.ksplit > 1 ? (w.ks + 1) * nsteps / p.ksplit : nsteps;
                tap = g0 / nch; ch = g0 % nch;
#pragma unroll 1
                for (int g = g0; g < g1; g += SPG) {
                    const int cnt = min(SPG, g1 - g);
                    mbar_wait(empty_bar(stage), phase ^ 1, 0x1100u | stage);
                    mbar_arrive_expect_tx(full_bar(stage), (a_bytes + b_bytes) * cnt);
                    const uint32_t adst = sA + stage * p.a_stage_bytes;
                    const uint32_t bdst = sB + stage * p.b_stage_bytes;
#pragma unroll 1
                    for (int j = 0; j < cnt; ++j) {
                        const ConvTap tp = P.taps[tap];
                        const bool first = ch < nch0;
                        const int c0 = first ? coff0 + ch * KC : coff1 + (ch - nch0) * KC;
                        tma_load_4d(adst + j * A_SUB, first ? m0 : m1, full_bar(stage), c0, ix0 + tp.dx, iy0 + tp.dy, iz0 + tp.dz);
                        bulk_g2s(bdst + j * b_bytes, wbase + (size_t(g + j) * ntiles + w.nt) * b_bytes, b_bytes, full_bar(stage));
                        if (++ch == nch) { ch = 0; ++tap; }
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================================== MMA issuer ====================================
        // The whole warp runs the loop (uniform control flow: descriptors stay in uniform registers instead of an R2UR / vote loop per
        // tcgen05.mma); one elected lane issues.
        {
            int stage = 0, phase = 0;
            uint32_t acc_cnt = 0;
            const int total_items = p.total_items, nstride = gridDim.x, ntile_max = p.ntile_max;
            const uint64_t a_desc0 = umma_desc_swz<SWZ>(sA);
            const uint64_t a_stage_u = uint64_t(p.a_stage_bytes >> 4), a_sub_u = uint64_t(A_SUB >> 4);
            const uint64_t b_stage_u = uint64_t(p.b_stage_bytes >> 4);
            for (int item = blockIdx.x; item < total_items; item += nstride, ++acc_cnt) {
                const TItem w = decode_titem(p, item);
                const ConvProblem& P = p.probs[w.pi];
                const int acc = acc_cnt & 1;
                mbar_wait(tempty_bar(acc), ((acc_cnt >> 1) & 1) ^ 1, 0x1200u | acc);
                tc_fence_after();
                const int ntile = P.ntile;
                const uint32_t d_tmem = tmem_base + uint32_t(acc * ntile_max);
                const uint32_t idesc = umma_idesc(128, ntile, 0, 0, 0, 0);
                const int nsteps = P.ntaps * (P.nch0 + P.nch1);
                const uint32_t b_lbo = uint32_t(ntile) * 16u;
                const uint64_t b_sub_u = uint64_t((uint32_t(ntile) * KC * 2u) >> 4);
                const uint64_t b_k_u = uint64_t((2u * b_lbo) >> 4);
                const uint64_t b_desc0 = umma_smem_desc(sB, b_lbo, 128u);
                bool first = true;
                if (KC == 64 && p.box[w.pi].xhalo) {
                    const int nch = P.nch0 + P.nch1, min_dx = p.box[w.pi].min_dx;
                    const int U = (P.ntaps / 3) * nch;
                    const int u0 = w.ks * U / p.ksplit, u1 = (w.ks + 1) * U / p.ksplit;
#pragma unroll 1
                    for (int u = u0; u < u1; ++u) {
                        const int tp3 = (u / nch) * 3;
                        uint32_t sh[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) sh[j] = uint32_t(P.taps[tp3 + j].dx - min_dx);   // start row of this dx tap inside the halo box
                        {
                            mbar_wait(full_bar(stage), phase, 0x1300u | stage);
                            tc_fence_after();
                            const uint64_t ad0 = a_desc0 + uint64_t(stage) * a_stage_u;
                            const uint64_t bd0 = b_desc0 + uint64_t(stage) * b_stage_u;
                            const bool f = first;
                            first = false;
                            if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                // start row sh (128 B each) inside the 1024-byte swizzle atom: descriptor address + 8*sh.  The swizzle is applied
                                // on absolute shared-memory address bits, so base_offset stays 0 (measured: setting it to sh gives wrong results)
                                const uint64_t ad = ad0 + uint64_t(8u * sh[j]);
                                const uint64_t bd = bd0 + uint64_t(j) * b_sub_u;
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k) {
                                    if (f && j == 0 && k == 0) umma_f16_first(d_tmem, ad + uint64_t(2 * k), bd + uint64_t(k) * b_k_u, idesc);
                                    else umma_f16_acc(d_tmem, ad + uint64_t(2 * k), bd + uint64_t(k) * b_k_u, idesc);
                                }
                            }
                            umma_commit(empty_bar(stage));
                            }
                            __syncwarp();
                            if (++stage == S) { stage = 0; phase ^= 1; }
                        }
                    }
                    if (elect_one()) umma_commit(tfull_bar(acc));
                    __syncwarp();
                    continue;
                }
                const int g0 = p.ksplit > 1 ? w.ks * nsteps / p.ksplit : 0;
                const int g1 = p.ksplit > 1 ? (w.ks + 1) * nsteps / p.ksplit : nsteps;
#pragma unroll 1
                for (int g = g0; g < g1; g += SPG) {
                    const int cnt = min(SPG, g1 - g);
                    mbar_wait(full_bar(stage), phase, 0x1300u | stage);
                    tc_fence_after();
                    uint64_t ad = a_desc0 + uint64_t(stage) * a_stage_u;
                    uint64_t bd = b_desc0 + uint64_t(stage) * b_stage_u;
                    const bool f = first;
                    first = false;
                    if (elect_one()) {
#pragma unroll 1
                    for (int j = 0; j < cnt; ++j, ad += a_sub_u, bd += b_sub_u) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
                            // inside a swizzle row the next K = 16 slice is 32 bytes further (descriptor address + 2)
                            if (f && j == 0 && k == 0) umma_f16_first(d_tmem, ad + uint64_t(2 * k), bd + uint64_t(k) * b_k_u, idesc);
                            else umma_f16_acc(d_tmem, ad + uint64_t(2 * k), bd + uint64_t(k) * b_k_u, idesc);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(tfull_bar(acc));
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        // Instruction-lean on purpose (ncu: with per-element predicated bias loads and 64-bit index math per 16-column chunk the
        // epilogue, not the copy engine or the tensor pipe, paced the cheap layers at ~13k clk per item): bias staged in shared
        // memory per problem, per-item invariants in registers, parity/column position advanced incrementally.
        const int r = threadIdx.x;  // 0..127 == TMEM lane == row of the box
        uint32_t acc_cnt = 0;
        float* sbias = sstats + 8 * p.ntot_max;   // [ntot_max], indexed by the (stacked) column
        int bias_prob = -1;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++acc_cnt) {
            const TItem w = decode_titem(p, item);
            const ConvProblem& P = p.probs[w.pi];
            const int cp = P.shuffle_cp, ntile = P.ntile;
            if (w.pi != bias_prob) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int ntot = ntile * P.ntiles;
                const int nreal = cp ? P.shuffle_nreal : P.n_real;
                for (int j = r; j < ntot; j += 128) {
                    const int jl = cp ? j % cp : j;
                    sbias[j] = (P.bias != nullptr && jl < nreal) ? __ldg(P.bias + jl) : 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_prob = w.pi;
            }
            const TBox bx = p.box[w.pi];
            const int txi = w.mt % bx.tiles_x;
            const int rest = w.mt / bx.tiles_x;
            const int tyi = rest % bx.tiles_y, tzi = rest / bx.tiles_y;
            const int lx = r % bx.pitch, lrest = r / bx.pitch;
            const int ox = txi * bx.bx + lx, oy = tyi * bx.by + lrest % bx.by, oz = tzi * bx.bz + lrest / bx.by;
            const bool rv = r < bx.rows && lx < bx.bx && ox < P.ow && oy < P.oh && oz < P.od;
            const int OH = P.OH, OW = P.OW, OD = P.OD;
            const long long M = 1LL * P.od * P.oh * P.ow;
            const long long m = (1LL * oz * P.oh + oy) * P.ow + ox;
            size_t vox = 0;
            if (rv && EPI != EPI_PLANAR32)
                vox = (size_t(oz * P.ostep + P.ooff_z) * OH + (oy * P.ostep + P.ooff_y)) * OW + (ox * P.ostep + P.ooff_x);
            uint8_t* const dst_base = static_cast<uint8_t*>(P.dst) + P.dst_coff * 2;
            const uint32_t dst_pitch = uint32_t(P.dst_cp) * 2u;
            const int n_real = P.n_real;
            int ncol = w.nt * ntile;                 // (stacked) column of the current chunk
            int par = 0, nloc = ncol;                // parity block and column inside it
            if (cp) { par = ncol / cp; nloc = ncol - par * cp; }
            const int acc = acc_cnt & 1;
            mbar_wait(tfull_bar(acc), (acc_cnt >> 1) & 1, 0x1400u | acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(acc * p.ntile_max);
#pragma unroll 1
            for (int c0 = 0; c0 < ntile; c0 += 16, ncol += 16) {
                float v[16];
                tmem_ld16(t_row + c0, v);
                if (p.ksplit > 1) {
                    // split-K: raw fp32 partial sums of this K range; bias / statistics / 16-bit store happen in the finalize kernel
                    if (rv) {
                        float4* o4 = reinterpret_cast<float4*>(p.split_scratch + (size_t(w.ks) * size_t(M) + size_t(m)) * size_t(ntile * P.ntiles) + ncol);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    continue;
                }
                {
                    const float4* sb4 = reinterpret_cast<const float4*>(sbias + ncol);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bq = sb4[j];
                        v[4 * j] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
                    }
                }
                bool rvc = rv;
                size_t voxc = vox;
                if (cp) {
                    // parity-stacked columns: block (pz,py,px) of lattice voxel o is destination voxel 2*o + (pz,py,px)
                    const int zz = 2 * oz + (par >> 2), yy = 2 * oy + ((par >> 1) & 1), xx = 2 * ox + (par & 1);
                    rvc = rv && zz < OD && yy < OH && xx < OW;
                    voxc = (size_t(zz) * OH + yy) * OW + xx;
                }
                if constexpr (EPI == EPI_PLANAR32) {
                    if (rv) {
                        float* out = static_cast<float*>(P.dst);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ncol + j < n_real) out[size_t(ncol + j) * M + m] = v[j];
                    }
                } else {
                    uint4* out = reinterpret_cast<uint4*>(dst_base + voxc * dst_pitch + nloc * 2);
                    if constexpr (EPI == EPI_ACCUM16) {
                        if (rvc) {
                            const uint4 o0 = out[0], o1 = out[1];
                            const uint32_t ow_[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 f = unpack2<false>(ow_[j]);
                                v[2 * j] += f.x;
                                v[2 * j + 1] += f.y;
                            }
                        }
                    }
                    if (rvc) {
                        uint4 q0, q1;
                        q0.x = pack2<false>(v[0], v[1]); q0.y = pack2<false>(v[2], v[3]);
                        q0.z = pack2<false>(v[4], v[5]); q0.w = pack2<false>(v[6], v[7]);
                        q1.x = pack2<false>(v[8], v[9]); q1.y = pack2<false>(v[10], v[11]);
                        q1.z = pack2<false>(v[12], v[13]); q1.w = pack2<false>(v[14], v[15]);
                        out[0] = q0;
                        out[1] = q1;
                    }
                    if (EPI == EPI_STORE16 && p.stats != nullptr) {
                        float a[16], q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            a[j] = rv ? v[j] : 0.f;
                            q[j] = a[j] * a[j];
                        }
                        halve_step_t<8, 16>(a, q, lane);
                        halve_step_t<4, 8>(a, q, lane);
                        halve_step_t<2, 4>(a, q, lane);
                        halve_step_t<1, 2>(a, q, lane);
                        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                        q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
                        if ((lane & 1) == 0) {
                            const int col = ncol + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                            float* ws = sstats + warp * 2 * p.ntot_max;   // one row per warp, one lane per column: fixed order
                            ws[col] += a[0];
                            ws[p.ntot_max + col] += q[0];
                        }
                    }
                }
                nloc += 16;
                if (cp && nloc == cp) { nloc = 0; ++par; }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
        if (p.stats != nullptr && p.ksplit == 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = r; i < 2 * p.ntot_max; i += 128)
                p.stats[size_t(blockIdx.x) * 2 * p.ntot_max + i] =
                    ((sstats[i] + sstats[2 * p.ntot_max + i]) + sstats[4 * p.ntot_max + i]) + sstats[6 * p.ntot_max + i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Split-K finalize: out[m][n] = sum_ks scratch[ks][m][n] (fixed order => deterministic) + bias, optional read-add, 16-bit NDHWC store,
// per-block partial rows of (sum, sum of squares) for the norm that follows.  Thread = (voxel, 8-channel chunk).
__global__ void conv_splitk_finalize_kernel(const float* __restrict__ scratch, int ksplit, long long M, int ntot, const float* __restrict__ bias,
                                            int n_real, uint8_t* __restrict__ dst, int dst_cp, int dst_coff, int accumulate,
                                            float* __restrict__ stats) {
    extern __shared__ float sm[];
    const int nch = ntot / 8;
    const int k = blockDim.x / nch;
    const int t = threadIdx.x;
    const int ch = t % nch, vsub = t / nch;
    float s0[8], s1[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s0[j] = s1[j] = 0.f;
        b[j] = (bias != nullptr && ch * 8 + j < n_real) ? bias[ch * 8 + j] : 0.f;
    }
    if (vsub < k) {
        for (long long m = (long long)blockIdx.x * k + vsub; m < M; m += (long long)gridDim.x * k) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            for (int ks = 0; ks < ksplit; ++ks) {
                const float4* q = reinterpret_cast<const float4*>(scratch + (size_t(ks) * size_t(M) + size_t(m)) * size_t(ntot) + ch * 8);
                const float4 a0 = q[0], a1 = q[1];
                v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += b[j];
            uint4* out = reinterpret_cast<uint4*>(dst + (size_t(m) * dst_cp + dst_coff + ch * 8) * 2);
            if (accumulate) {
                const uint4 o = *out;
                const uint32_t ow_[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2<false>(ow_[j]);
                    v[2 * j] += f.x;
                    v[2 * j + 1] += f.y;
                }
            }
            uint4 qv;
            qv.x = pack2<false>(v[0], v[1]); qv.y = pack2<false>(v[2], v[3]);
            qv.z = pack2<false>(v[4], v[5]); qv.w = pack2<false>(v[6], v[7]);
            *out = qv;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s0[j] += v[j];
                s1[j] = fmaf(v[j], v[j], s1[j]);
            }
        }
    }
    if (stats == nullptr) return;
    float* row = sm + size_t(vsub) * ntot * 2;
    if (vsub < k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            row[ch * 8 + j] = s0[j];
            row[ntot + ch * 8 + j] = s1[j];
        }
    }
    __syncthreads();
    for (int i = t; i < 2 * ntot; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < k; ++r) acc += sm[size_t(r) * ntot * 2 + i];
        stats[size_t(blockIdx.x) * 2 * ntot + i] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// fewest boxes of <= 128 voxels that cover the lattice; ties -> longest run along x
void choose_box(int ow, int oh, int od, int istride, int halo, TBox& b) {
    // halo = 2: the box carries a one-voxel x halo on both sides and the M = 128 operand may start up to 2 rows into it
    const int cap = halo ? 126 : 128;
    long long best = -1;
    b.bx = 0;
    for (int bx = 1; bx <= std::min(ow, cap - halo); ++bx) {
        if (bx * istride > 256) break;
        for (int by = 1; by <= std::min(oh, cap / (bx + halo)); ++by) {
            if (by * istride > 256) break;
            const int bz = std::min({od, cap / ((bx + halo) * by), 256 / istride});
            const long long tiles = 1LL * ((ow + bx - 1) / bx) * ((oh + by - 1) / by) * ((od + bz - 1) / bz);
            if (best < 0 || tiles < best || (tiles == best && bx > b.bx)) {
                best = tiles;
                b.bx = bx; b.by = by; b.bz = bz;
            }
        }
    }
    b.tiles_x = (ow + b.bx - 1) / b.bx;
    b.tiles_y = (oh + b.by - 1) / b.by;
    b.pitch = b.bx + halo;
    b.rows = b.pitch * b.by * b.bz;
    b.xhalo = halo ? 1 : 0;
}

// k3 s1 problem whose 27 taps come as 9 triples sharing (dz,dy) with dx covering {-1,0,1}: eligible for the x-halo box
bool xhalo_ok(const ConvProblem& P, int kc, int* min_dx) {
    static const bool disabled = std::getenv("U3D_NO_XHALO") != nullptr;
    if (disabled || kc != 64 || P.ntaps != 27 || P.istride != 1 || P.shuffle_cp) return false;
    for (int t = 0; t < 27; t += 3) {
        int seen = 0;
        for (int j = 0; j < 3; ++j) {
            const ConvTap& a = P.taps[t + j];
            if (a.dz != P.taps[t].dz || a.dy != P.taps[t].dy || a.dx < -1 || a.dx > 1) return false;
            seen |= 1 << (a.dx + 1);
        }
        if (seen != 7) return false;
    }
    *min_dx = -1;
    return true;
}

template <int EPI, int KC>
int launch_tma_t(const TParams& tp, const TMaps& maps, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_tma_kernel<EPI, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_tma_kernel<EPI, KC><<<grid, kTThreads, smem, stream>>>(tp, maps);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int EPI>
int launch_tma_kc(const TParams& tp, const TMaps& maps, int kc, int grid, size_t smem, cudaStream_t stream) {
    if (kc == 64) return launch_tma_t<EPI, 64>(tp, maps, grid, smem, stream);
    if (kc == 32) return launch_tma_t<EPI, 32>(tp, maps, grid, smem, stream);
    return launch_tma_t<EPI, 16>(tp, maps, grid, smem, stream);
}

}  // namespace

unsigned int read_device_error_tma() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

bool conv_tma_available() {
    static const bool disabled = std::getenv("U3D_NO_TMA") != nullptr;
    return !disabled && encode_fn() != nullptr;
}

bool conv_tma_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    static const bool disabled = std::getenv("U3D_NO_TMA") != nullptr;
    static const int min_kc = std::getenv("U3D_TMA_MIN_KC") ? std::atoi(std::getenv("U3D_TMA_MIN_KC")) : 16;
    if (disabled || probs.empty() || probs.size() > kTMaxProb || encode_fn() == nullptr) return false;
    if (cfg.kc != 16 && cfg.kc != 32 && cfg.kc != 64) return false;
    if (cfg.kc < min_kc || cfg.a_bf16 || cfg.b_bf16) return false;
    for (const auto& P : probs) {
        if (P.istride != 1 && P.istride != 2) return false;
        if (P.shuffle_cp && cfg.epi == EPI_PLANAR32) return false;
        if (P.ntile % 16 || P.ntile < 16 || P.ntile > 256 || P.ntaps < 1 || P.ntaps > 27 || P.nch0 + P.nch1 < 1) return false;
        // tensor-map constraints: 16-byte aligned base and strides
        if ((reinterpret_cast<uintptr_t>(P.src0) & 15) || (P.c0p * 2) % 16) return false;
        if (P.nch1 && ((reinterpret_cast<uintptr_t>(P.src1) & 15) || (P.c1p * 2) % 16)) return false;
    }
    return true;
}

int conv_tma_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    TParams tp;
    std::memset(&tp, 0, sizeof(tp));
    TMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    tp.nprob = int(probs.size());
    const int kc = cfg.kc;
    const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    int items = 0, ntile_max = 16, ntot_max = 16;
    bool any_xhalo = false;
    for (int i = 0; i < tp.nprob; ++i) {
        ConvProblem& P = tp.probs[i];
        P = probs[i];
        TBox& b = tp.box[i];
        int min_dx = 0;
        const bool xh = P.ntile <= 128 && xhalo_ok(P, kc, &min_dx);   // 3 weight slices per stage: wider N tiles leave < 3 stages
        choose_box(P.ow, P.oh, P.od, P.istride, xh ? 2 : 0, b);
        b.min_dx = min_dx;
        any_xhalo = any_xhalo || xh;
        P.mtiles = b.tiles_x * b.tiles_y * ((P.od + b.bz - 1) / b.bz);
        P.item_base = items;
        items += P.mtiles * P.ntiles;   // (x ksplit below: split-K only with a single problem)
        ntile_max = std::max(ntile_max, P.ntile);
        ntot_max = std::max(ntot_max, P.ntile * P.ntiles);
        for (int s = 0; s < 2; ++s) {
            const void* base = s ? P.src1 : P.src0;
            const int cp = s ? P.c1p : P.c0p;
            if (s && P.nch1 == 0) continue;
            const cuuint64_t gdim[4] = {cuuint64_t(cp), cuuint64_t(P.in_w), cuuint64_t(P.in_h), cuuint64_t(P.in_d)};
            const cuuint64_t gstr[3] = {cuuint64_t(cp) * 2, cuuint64_t(cp) * 2 * P.in_w, cuuint64_t(cp) * 2 * P.in_w * P.in_h};
            const cuuint32_t box[4] = {cuuint32_t(kc), cuuint32_t(b.pitch * P.istride), cuuint32_t(b.by * P.istride), cuuint32_t(b.bz * P.istride)};
            const cuuint32_t estr[4] = {1, cuuint32_t(P.istride), cuuint32_t(P.istride), cuuint32_t(P.istride)};
            const CUresult r = encode_fn()(&maps.m[2 * i + s], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("conv_tma_launch: cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
                return 1;
            }
        }
    }
    // deterministic split-K for the deep levels: a handful of boxes x N tiles cannot fill 148 SMs, and each item walks hundreds of K
    // units serially.  Split the K units over up to 16 items that write separate fp32 slices; a finalize kernel sums them in order.
    tp.ksplit = 1;
    tp.split_scratch = nullptr;
    {
        const ConvProblem& P0 = tp.probs[0];
        static const bool no_split = std::getenv("U3D_NO_SPLITK") != nullptr;
        const int nch = P0.nch0 + P0.nch1;
        const int units = tp.box[0].xhalo ? (P0.ntaps / 3) * nch : P0.ntaps * nch;
        const long long M = 1LL * P0.od * P0.oh * P0.ow;
        if (!no_split && tp.nprob == 1 && kc == 64 && !P0.shuffle_cp && P0.ostep == 1 && cfg.epi != EPI_PLANAR32 && items * 2 <= device_sm_count() &&
            units >= 8 && cfg.splitk_scratch != nullptr && P0.od == P0.OD && P0.oh == P0.OH && P0.ow == P0.OW) {
            int ks = std::min({units / 4, device_sm_count() / items, 16});
            const size_t per_slice = size_t(M) * size_t(ntot_max) * 4;
            while (ks > 1 && per_slice * ks > cfg.splitk_scratch_bytes) --ks;
            if (ks > 1) {
                tp.ksplit = ks;
                tp.split_scratch = cfg.splitk_scratch;
                items *= ks;
            }
        }
    }
    tp.total_items = items;
    tp.ntile_max = ntile_max;
    tp.ntot_max = ntot_max;
    int cols = 32;
    while (cols < 2 * ntile_max) cols <<= 1;
    tp.tmem_cols = cols;
    const int spg = 64 / kc;
    tp.a_stage_bytes = uint32_t(128 * kc * 2 * spg);                 // 16 KB
    tp.b_stage_bytes = uint32_t(ntile_max * kc * 2 * spg);
    if (any_xhalo) {   // 130 rows (the operand may start 2 rows in), swizzle atoms stay 1024-byte aligned; 3 weight slices per stage
        tp.a_stage_bytes = 17 * 1024;
        tp.b_stage_bytes *= 3;
    }
    const size_t fixed = size_t(9) * ntot_max * 4 + 8 * (2 * 16 + 4) + 16 + 2048;
    int stages = int((220 * 1024 - fixed) / (tp.a_stage_bytes + tp.b_stage_bytes));
    stages = std::min(stages, 12);
    if (stages < 2) { set_error("conv_tma_launch: tile too large for the smem ring"); return 1; }
    tp.stages = stages;
    tp.off_b = uint32_t(stages) * tp.a_stage_bytes;
    tp.off_stats = tp.off_b + uint32_t(stages) * tp.b_stage_bytes;
    tp.off_bars = uint32_t((tp.off_stats + 9 * ntot_max * 4 + 15) & ~15u);
    const size_t smem = tp.off_bars + 8 * (2 * stages + 4) + 16;
    tp.stats = (cfg.epi == EPI_STORE16 && probs.size() == 1) ? cfg.stats_partials : nullptr;
    const int grid = std::max(1, std::min(items, device_sm_count()));
    if (cfg.stats_grid_out) *cfg.stats_grid_out = grid;
    if (tp.ksplit > 1) {
        if (launch_tma_kc<EPI_STORE16>(tp, maps, kc, grid, smem, stream)) return 1;
        const ConvProblem& P0 = tp.probs[0];
        const long long M = 1LL * P0.od * P0.oh * P0.ow;
        const int nch8 = ntot_max / 8;
        const int k = nch8 >= 256 ? 1 : 256 / nch8;
        const int block = nch8 * k;
        const int fgrid = int(std::max<long long>(1, std::min<long long>((M + k - 1) / k, device_sm_count())));
        float* st = tp.stats;
        conv_splitk_finalize_kernel<<<fgrid, block, st ? size_t(k) * ntot_max * 2 * sizeof(float) : 0, stream>>>(
            tp.split_scratch, tp.ksplit, M, ntot_max, P0.bias, P0.n_real, static_cast<uint8_t*>(P0.dst), P0.dst_cp, P0.dst_coff,
            cfg.epi == EPI_ACCUM16 ? 1 : 0, st);
        U3D_CUDA_CHECK(cudaGetLastError());
        if (cfg.stats_grid_out) *cfg.stats_grid_out = fgrid;
        return 0;
    }
    if (cfg.epi == EPI_PLANAR32) return launch_tma_kc<EPI_PLANAR32>(tp, maps, kc, grid, smem, stream);
    if (cfg.epi == EPI_STORE16) return launch_tma_kc<EPI_STORE16>(tp, maps, kc, grid, smem, stream);
    return launch_tma_kc<EPI_ACCUM16>(tp, maps, kc, grid, smem, stream);
}

}  // namespace u3d
