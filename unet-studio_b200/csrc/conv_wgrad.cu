// Weight-gradient GEMM on tcgen05 / TMEM (sm_100a): replaces autograd's conv3d / conv_transpose3d
// weight gradients (reference: total_loss.backward(), /root/reference/train.cpp:706).
//
//   dW[n][m][tap] += sum_{v in lattice} T[v*tstride + tap][m] * U[v][n]            (see u3d.h)
//
// The reduction dimension K is the voxel lattice, so both operands are MN-major in their natural
// NDHWC order: a voxel's channel vector is one 16-byte-chunked row.  Each smem stage holds 64 voxels:
//   A (M side): 16 chunks x 64 voxels x 16 B  — M = 128 rows = (tap group) x (channels), i.e. several
//               taps of the same K-block are stacked along M so that 16-channel layers still fill
//               the 128-row MMA;
//   B (N side): ntile/8 chunks x 64 voxels x 16 B.
// Canonical SWIZZLE_NONE MN-major layout: 8 voxels x 16 B = one core matrix (LBO = 128 B between
// K groups, SBO = 1024 B between channel chunks).  Work item = (problem, M tile, N tile, K split);
// results are reduced across K splits with fp32 atomics straight into the reference-layout
// gradient tensor (summation order across CTAs is not fixed; see DESIGN.md "determinism").
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kThreads = 288;
constexpr int kMaxProb = 8;
constexpr int kKB = 64;  // voxels per stage
// partial-tile mode (scratch + sum kernel) for every problem with a real N: the 2-channel output heads keep their (few) direct atomics
__host__ __device__ inline bool use_partial(const WgradProblem& P) { return P.u_creal >= 8; }

struct WParams {
    WgradProblem probs[kMaxProb];
    int nprob;
    int total_items;
    int stages;
    int t_fmt, u_fmt;
    int ntile_max;
    int tmem_cols;
    uint32_t off_b, off_bars;
    float* scratch;      // != nullptr: every work item stores its 128 x ntile accumulator tile to slot `item` ([n][128 rows]) and
                         // conv_wgrad_sum_kernel reduces the K splits (instead of 128 x ntile fp32 atomics per item)
    int nsum_blocks;
};

struct WItem {
    int pi, mt, nt, ks;
};

__device__ __forceinline__ WItem decode_witem(const WParams& p, int item) {
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < p.nprob; ++i)
        if (item >= p.probs[i].item_base) pi = i;
    const WgradProblem& P = p.probs[pi];
    int local = item - P.item_base;
    WItem w;
    w.pi = pi;
    w.ks = local % P.ksplit;
    local /= P.ksplit;
    w.nt = local % P.ntiles;
    w.mt = local / P.ntiles;
    return w;
}

// channels per M tile and channel tiles per tap
__device__ __forceinline__ int cpt_of(const WgradProblem& P) { return P.t_c <= 128 ? P.t_c : 128; }

__device__ __forceinline__ void kblock_range(const WgradProblem& P, int ks, int& kb0, int& kb1) {
    const int K = P.ld * P.lh * P.lw;
    const int kblocks = (K + kKB - 1) / kKB;
    const int per = (kblocks + P.ksplit - 1) / P.ksplit;
    kb0 = ks * per;
    kb1 = min(kblocks, kb0 + per);
    if (kb0 > kb1) kb0 = kb1;
}

template <int LAG>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ WParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t a_stage_bytes = 16u * kKB * 16u;                      // 16 KB
    const uint32_t b_stage_bytes = uint32_t(p.ntile_max / 8) * kKB * 16u;
    const uint32_t sA = smem_u32(smem);
    const uint32_t sB = sA + p.off_b;
    const uint32_t bars = sA + p.off_bars;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * S + 4));

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(smem_u32(tmem_ptr_smem), p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 4 && warp < 8) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        // adjacent lanes take the two 16-byte halves of one voxel row so every 32-byte L2 sector a warp touches is fully used
        const int h = t & 1, v = t >> 1;
        int stage = 0, phase = 0, lag_stage = 0;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const WItem w = decode_witem(p, item);
            const WgradProblem& P = p.probs[w.pi];
            const int K = P.ld * P.lh * P.lw;
            const int cpt = cpt_of(P);
            const int ctiles = (P.t_c + cpt - 1) / cpt;
            const int tapgrp = w.mt / ctiles, ctile = w.mt % ctiles;
            const int tap0 = tapgrp * P.tg;
            const int ntap_here = min(P.tg, P.ntaps - tap0);
            const int nchunks_n = P.ntile / 8;
            int kb0, kb1;
            kblock_range(P, w.ks, kb0, kb1);
            // per-item constants of this thread's 8 A chunks (tap, channel offset, linear voxel delta): the stage loop below only
            // decodes the voxel once and adds — the producers, not the tensor pipe, were the bound of this kernel
            const int t_d = P.t_d, t_h = P.t_h, t_w = P.t_w, ts = P.tstride, lw = P.lw, lh = P.lh;
            const uint32_t t_pitch = uint32_t(P.t_cp) * 2u, u_pitch = uint32_t(P.u_cp) * 2u;
            const uint8_t* const tbase = static_cast<const uint8_t*>(P.T) + P.t_coff * 2;
            const uint8_t* const ubase = static_cast<const uint8_t*>(P.U) + (P.u_coff + w.nt * P.ntile) * 2;
            int cdz[8], cdy[8], cdx[8], ccoff[8];
            long long cdelta[8];
            uint32_t cok = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row0 = (h + 2 * i) * 8;
                const int tl = row0 / cpt;
                const int c = row0 - tl * cpt + ctile * 128;
                const bool ok = tl < ntap_here && c < P.t_c;
                const ConvTap tp = P.taps[ok ? tap0 + tl : 0];
                cdz[i] = tp.dz; cdy[i] = tp.dy; cdx[i] = tp.dx;
                cdelta[i] = ((long long)tp.dz * t_h + tp.dy) * t_w + tp.dx;
                ccoff[i] = c * 2;
                cok |= (ok ? 1u : 0u) << i;
            }
#pragma unroll 1
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                mbar_wait(empty_bar(stage), phase ^ 1, 0x500u | stage);
                const int vg = kb * kKB + v;
                const bool vv = vg < K;
                int lx = 0, ly = 0, lz = 0;
                if (vv) {
                    lx = vg % lw;
                    const int q = vg / lw;
                    ly = q % lh;
                    lz = q / lh;
                }
                const int bz = lz * ts, by = ly * ts, bx = lx * ts;
                const long long vbase = ((long long)bz * t_h + by) * t_w + bx;
                // ---- A: tapped tensor, 8 chunks per thread
                const uint32_t a_dst = sA + stage * a_stage_bytes + v * 16u + h * (kKB * 16u);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool ok = vv && ((cok >> i) & 1u) && (unsigned)(bz + cdz[i]) < (unsigned)t_d &&
                                    (unsigned)(by + cdy[i]) < (unsigned)t_h && (unsigned)(bx + cdx[i]) < (unsigned)t_w;
                    const uint8_t* src = ok ? tbase + (vbase + cdelta[i]) * t_pitch + ccoff[i] : tbase;
                    cp_async16_ca(a_dst + i * (2u * kKB * 16u), src, ok ? 16u : 0u);
                }
                // ---- B: untapped tensor
                const uint32_t b_dst = sB + stage * b_stage_bytes + v * 16u;
                const uint8_t* usrc = vv ? ubase + size_t(vg) * u_pitch : ubase;
#pragma unroll 1
                for (int nc = h; nc < nchunks_n; nc += 2)
                    cp_async16(b_dst + nc * (kKB * 16u), vv ? usrc + nc * 16 : ubase, vv ? 16u : 0u);
                cp_async_mbar_arrive(full_bar(stage));   // fires when this thread's copies have landed (see conv_igemm.cu)
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        cp_async_wait<0>();
    } else if (warp == 8) {
        // ===================================== MMA issuer ====================================
        {
            // lean issue loop (see conv_halo.cu / tools/mma_bench.cu).  The whole warp runs it (uniform control flow keeps the descriptors in
            // uniform registers); one elected lane issues.
            int stage = 0, phase = 0;
            uint32_t acc_cnt = 0;
            const int total_items = p.total_items, nstride = gridDim.x, ntile_max = p.ntile_max;
            const uint64_t a_desc0 = umma_smem_desc(sA, 128u, kKB * 16u);
            const uint64_t b_desc0 = umma_smem_desc(sB, 128u, kKB * 16u);
            const uint64_t a_stage_u = uint64_t(a_stage_bytes >> 4), b_stage_u = uint64_t(b_stage_bytes >> 4);
            for (int item = blockIdx.x; item < total_items; item += nstride) {
                const WItem w = decode_witem(p, item);
                const WgradProblem& P = p.probs[w.pi];
                int kb0, kb1;
                kblock_range(P, w.ks, kb0, kb1);
                if (kb0 >= kb1) continue;  // empty split: no accumulator is produced (epilogue skips it too)
                const int acc = acc_cnt & 1;
                mbar_wait(tempty_bar(acc), ((acc_cnt >> 1) & 1) ^ 1, 0x600u | acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * ntile_max);
                const uint32_t idesc = umma_idesc(128, P.ntile, p.t_fmt, p.u_fmt, 1, 1);
                bool first = true;
#pragma unroll 1
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x700u | stage);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t ad = a_desc0 + uint64_t(stage) * a_stage_u;
                    const uint64_t bd = b_desc0 + uint64_t(stage) * b_stage_u;
                    const bool f = first;
                    first = false;
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < kKB / 16; ++j) {
                            if (f && j == 0) umma_f16_first(d_tmem, ad + uint64_t(j) * 16u, bd + uint64_t(j) * 16u, idesc);
                            else umma_f16_acc(d_tmem, ad + uint64_t(j) * 16u, bd + uint64_t(j) * 16u, idesc);
                        }
                        umma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(tfull_bar(acc));
                __syncwarp();
                ++acc_cnt;
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        const int r = threadIdx.x;
        uint32_t acc_cnt = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const WItem w = decode_witem(p, item);
            const WgradProblem& P = p.probs[w.pi];
            int kb0, kb1;
            kblock_range(P, w.ks, kb0, kb1);
            if (kb0 >= kb1) continue;
            const int cpt = cpt_of(P);
            const int ctiles = (P.t_c + cpt - 1) / cpt;
            const int tapgrp = w.mt / ctiles, ctile = w.mt % ctiles;
            const int tap0 = tapgrp * P.tg;
            const int ntap_here = min(P.tg, P.ntaps - tap0);
            const int tl = r / cpt;
            const int c = r - tl * cpt + ctile * 128;
            const bool rv = tl < ntap_here && c < P.t_creal;
            const int acc = acc_cnt & 1;
            mbar_wait(tfull_bar(acc), (acc_cnt >> 1) & 1, 0x800u | acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(acc * p.ntile_max);
            float* dwrow = nullptr;
            if (rv) dwrow = P.dw + size_t(P.w_moff + c) * P.w_ktaps + P.tap_ref[tap0 + tl];
            const size_t nstride = size_t(P.w_mtot) * P.w_ktaps;
            float* const slot = (p.scratch && use_partial(P)) ? p.scratch + size_t(item) * (128u * size_t(p.ntile_max)) + r : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < P.ntile; c0 += 16) {
                float v[16];
                tmem_ld16(t_row + c0, v);
                if (slot != nullptr) {   // column-major tile: the lanes (consecutive rows) write consecutive floats
#pragma unroll
                    for (int j = 0; j < 16; ++j) slot[size_t(c0 + j) * 128] = v[j];
                } else if (rv) {
                    const int n0 = w.nt * P.ntile + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < P.u_creal) atomicAdd(dwrow + size_t(P.w_noff + n0 + j) * nstride, v[j]);
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
            ++acc_cnt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Partial-tile mode, second kernel.  For one output channel n the gradient's [ci][reference taps] block is ONE contiguous run of the
// reference-layout tensor, but the accumulator tiles hold it scattered over the tap-group tiles (rows = (tap in group, ci)), and a
// tile row is written 4 bytes at a time 108 bytes apart when it goes straight to the gradient.  Here block (problem, n, channel tile,
// slice of 16 K splits) reads the rows of every tap-group tile for its n (thread = tile row: coalesced 512-byte rows of the
// column-major scratch tiles), sums its K splits in a fixed order, transposes through shared memory and adds the run to the gradient
// with contiguous accesses (plain read-modify-write for one slice = deterministic; fp32 atomics when slices meet).
__global__ void __launch_bounds__(128) conv_wgrad_sum_kernel(const __grid_constant__ WParams p) {
    __shared__ float srun[128 * 27];
    int b = blockIdx.x, pi = 0;
    for (; pi < p.nprob; ++pi) {
        const WgradProblem& Q = p.probs[pi];
        const int cpt_q = cpt_of(Q);
        const int nb = use_partial(Q) ? Q.u_creal * ((Q.t_c + cpt_q - 1) / cpt_q) : 0;
        if (b < nb) break;
        b -= nb;
    }
    if (pi >= p.nprob) return;
    const WgradProblem& P = p.probs[pi];
    const int cpt = cpt_of(P);
    const int ctiles = (P.t_c + cpt - 1) / cpt;
    const int ctile = b % ctiles, n_abs = b / ctiles;
    const int nt = n_abs / P.ntile, n = n_abs % P.ntile;
    const int nsl = (P.ksplit + 15) / 16;                  // slices of THIS problem (the grid is sized for the largest)
    const int k0 = blockIdx.y * 16, k1 = min(P.ksplit, k0 + 16);
    const int r = threadIdx.x;
    const int tl = r / cpt, cl = r - tl * cpt;            // tap in group, channel in tile
    const int c = cl + ctile * 128;
    const int creal = min(cpt, P.t_creal - ctile * 128);  // real channels of this tile
    const int ktaps = P.w_ktaps;
    const size_t slot = 128u * size_t(p.ntile_max);
    if (k0 < k1) {
        const int ngroups = (P.ntaps + P.tg - 1) / P.tg;
        for (int tapgrp = 0; tapgrp < ngroups; ++tapgrp) {
            const int t = tapgrp * P.tg + tl;
            const int mt = tapgrp * ctiles + ctile;
            const float* const src = p.scratch + size_t(P.item_base + (mt * P.ntiles + nt) * P.ksplit) * slot + size_t(n) * 128 + r;
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = k0 + k < k1 ? __ldcs(src + size_t(k0 + k) * slot) : 0.f;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += v[k];
            if (t < P.ntaps && tl < P.tg && c < P.t_creal) srun[cl * ktaps + P.tap_ref[t]] = acc;
        }
    }
    __syncthreads();
    if (k0 >= k1 || creal <= 0) return;
    float* const dst = P.dw + (size_t(P.w_noff + n_abs) * P.w_mtot + P.w_moff + ctile * 128) * ktaps;
    const int run = creal * ktaps;
    // (taps the problem does not cover -- none for the layers of this network -- are left untouched)
    if (P.ntaps == ktaps) {
        for (int i = r; i < run; i += 128) {
            if (nsl > 1) atomicAdd(dst + i, srun[i]);
            else dst[i] += srun[i];
        }
    } else {
        for (int t = 0; t < P.ntaps; ++t)
            for (int cc = r; cc < creal; cc += 128) {
                const int i = cc * ktaps + P.tap_ref[t];
                if (nsl > 1) atomicAdd(dst + i, srun[i]);
                else dst[i] += srun[i];
            }
    }
}

}  // namespace

unsigned int read_device_error_wgrad() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

int conv_wgrad_launch(const std::vector<WgradProblem>& probs, const WgradLaunch& cfg, WgradProblem*, cudaStream_t stream) {
    if (probs.empty()) return 0;
    if (probs.size() > kMaxProb) {
        set_error("conv_wgrad_launch: too many problems");
        return 1;
    }
    WParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.nprob = int(probs.size());
    kp.t_fmt = cfg.t_bf16;
    kp.u_fmt = cfg.u_bf16;
    const int sms = device_sm_count();
    int base_items = 0;
    int ntile_max = 16;
    for (size_t i = 0; i < probs.size(); ++i) {
        WgradProblem& P = kp.probs[i];
        P = probs[i];
        if (P.t_c % 16 || P.u_c % 16 || P.t_c < 16 || P.u_c < 16 || P.ntaps < 1 || P.ntaps > 27) {
            set_error("conv_wgrad_launch: bad problem shape");
            return 1;
        }
        const int cpt = P.t_c <= 128 ? P.t_c : 128;
        P.tg = P.t_c <= 128 ? 128 / P.t_c : 1;
        const int ctiles = (P.t_c + cpt - 1) / cpt;
        P.mtiles = ((P.ntaps + P.tg - 1) / P.tg) * ctiles;
        int nt = std::min(P.u_c, 256);
        while (P.u_c % nt) nt -= 16;
        P.ntile = nt;
        P.ntiles = P.u_c / nt;
        ntile_max = std::max(ntile_max, nt);
        base_items += P.mtiles * P.ntiles;
    }
    int items = 0;
    for (int i = 0; i < kp.nprob; ++i) {
        WgradProblem& P = kp.probs[i];
        const long long K = 1LL * P.ld * P.lh * P.lw;
        const int kblocks = int((K + kKB - 1) / kKB);
        int ks = std::max(1, (2 * sms + base_items - 1) / base_items);
        ks = std::min(ks, std::max(1, kblocks / 4));
        // make sure no split is empty
        while (ks > 1 && (ks - 1) * ((kblocks + ks - 1) / ks) >= kblocks) --ks;
        P.ksplit = ks;
        P.item_base = items;
        items += P.mtiles * P.ntiles * ks;
    }
    kp.total_items = items;
    kp.ntile_max = ntile_max;
    int cols = 32;
    while (cols < 2 * ntile_max) cols <<= 1;
    kp.tmem_cols = cols;
    const size_t a_stage = size_t(16) * kKB * 16, b_stage = size_t(ntile_max / 8) * kKB * 16;
    int stages = int((200 * 1024) / (a_stage + b_stage));
    stages = std::min(stages, 12);
    kp.stages = stages;
    kp.off_b = uint32_t(stages * a_stage);
    kp.off_bars = uint32_t(kp.off_b + stages * b_stage);
    const size_t smem = kp.off_bars + 8 * (2 * stages + 4) + 16;
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = std::max(1, std::min(items, sms));
    static const bool no_partial = std::getenv("U3D_WGRAD_ATOMICS") != nullptr;
    int max_ksplit = 1;
    if (!no_partial && cfg.partial_scratch != nullptr && size_t(items) * 128 * ntile_max * 4 <= cfg.partial_scratch_bytes) {
        kp.scratch = cfg.partial_scratch;
        for (int i = 0; i < kp.nprob; ++i) {
            const WgradProblem& P = kp.probs[i];
            const int cpt = P.t_c <= 128 ? P.t_c : 128;
            if (use_partial(P)) {
                kp.nsum_blocks += P.u_creal * ((P.t_c + cpt - 1) / cpt);
                max_ksplit = std::max(max_ksplit, P.ksplit);
            }
        }
    }
    if (stages >= 8) conv_wgrad_kernel<6><<<grid, kThreads, smem, stream>>>(kp);
    else conv_wgrad_kernel<2><<<grid, kThreads, smem, stream>>>(kp);
    U3D_CUDA_CHECK(cudaGetLastError());
    if (kp.scratch != nullptr && kp.nsum_blocks > 0) {
        conv_wgrad_sum_kernel<<<dim3(unsigned(kp.nsum_blocks), unsigned((max_ksplit + 15) / 16)), 128, 0, stream>>>(kp);
        U3D_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

}  // namespace u3d
