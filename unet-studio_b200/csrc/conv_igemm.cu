// Implicit-GEMM gather convolution on tcgen05 / TMEM (sm_100a).
//
// Replaces the libtorch->cuDNN conv3d / conv_transpose3d calls of the reference
// (/root/reference/unet.cpp:46-72) and their autograd data-gradients.  See u3d.h for the problem form.
//
// CTA = 288 threads, persistent over work items (problem, 128-voxel M tile, N tile):
//   warps 0-3  epilogue : tcgen05.ld of the fp32 accumulator row of "their" voxel (thread r <-> TMEM lane r),
//                         + bias, per-channel sum / sum-of-squares partials for the following norm,
//                         16-bit NDHWC store (or fp32 planar logits)
//   warps 4-7  producers: thread r gathers the K-chunk of voxel r for the current tap with 16-byte
//                         cp.async (zero fill outside the volume = the conv padding) into the canonical
//                         SWIZZLE_NONE K-major layout; one thread also issues the weight tile as a 1-D
//                         bulk copy (UBLKCP) that completes on the same mbarrier
//   warp  8    MMA      : one thread issues tcgen05.mma (M=128, N=ntile, K=16) from shared memory into
//                         a double-buffered TMEM accumulator and tcgen05.commit's the stage back
// Pipelines: smem ring full/empty mbarriers (producers <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"
#include "u3d.h"

namespace u3d {

namespace {

constexpr int kThreads = 288;
constexpr int kMaxProb = 8;

struct KParams {
    ConvProblem probs[kMaxProb];
    int nprob;
    int total_items;
    int kc;
    int stages;
    int spg;
    int a_fmt, b_fmt;
    int ntile_max;   // TMEM columns per accumulator buffer
    int ntot_max;    // stats row length
    int tmem_cols;
    uint32_t off_b, off_stats, off_bars;
    uint32_t off_w;     // resident weight pack (whole layer) when resident_w != 0
    int resident_w;
    uint32_t w_bytes;
    float* stats;    // [grid][2][ntot_max] or nullptr
};

struct Item {
    int pi, mt, nt;
};

__device__ __forceinline__ Item decode_item(const KParams& p, int item) {
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < p.nprob; ++i)
        if (item >= p.probs[i].item_base) pi = i;
    const int local = item - p.probs[pi].item_base;
    const int ntiles = p.probs[pi].ntiles;
    return Item{pi, local / ntiles, local % ntiles};
}

// One recursive-halving step of the 16-column transpose-reduce used for the norm statistics: lanes whose
// `BIT` is set keep the upper HALF columns, the others the lower HALF, and each adds what its partner sent.
template <int HALF, int BIT>
__device__ __forceinline__ void halve_step(float (&a)[16], float (&q)[16], int lane) {
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

// LAG = cp.async groups each producer thread keeps in flight (memory-level parallelism of the gather)
template <int EPI, int LAG, int KC>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    constexpr int kc = KC;
    const int spg = p.spg;                                   // K steps per smem stage
    const uint32_t a_sub_bytes = 128u * kc * 2u;
    const uint32_t a_stage_bytes = a_sub_bytes * spg;
    const uint32_t b_stage_bytes = uint32_t(p.ntile_max) * kc * 2u * spg;
    const uint32_t sA = smem_u32(smem);
    const uint32_t sB = sA + p.off_b;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sA + p.off_bars;
    // barrier slots (8 B each): full[S], empty[S], tmem_full[2], tmem_empty[2]; then tmem base pointer
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * S + 4));

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), p.resident_w ? 128 : 129);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 8 * p.ntot_max; i += kThreads) sstats[i] = 0.f;  // [4 epilogue warps][2][ntot]
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
        const int r = threadIdx.x - 128;
        int stage = 0, phase = 0, lag_stage = 0;
        uint32_t it = 0;
        if (p.resident_w && blockIdx.x < p.total_items) {
            // whole-layer weight pack stays in shared memory for the life of the CTA; it rides in the first cp.async
            // group, so the first full-barrier arrival also publishes it
            const uint8_t* wsrc = static_cast<const uint8_t*>(p.probs[0].wpack);
            for (uint32_t o = r * 16u; o < p.w_bytes; o += 128u * 16u) cp_async16(sA + p.off_w + o, wsrc + o, 16u);
        }
        // lane mapping: CPR adjacent lanes fetch the CPR 16-byte chunks of ONE voxel row, so every 32-byte L2 sector a warp
        // instruction touches is fully used (a row-per-thread mapping reads half of each sector: 2x L2 traffic, measured)
        constexpr int CPR = KC / 8;          // 16-byte chunks per row
        constexpr int RSTEP = 128 / CPR;     // rows covered per pass; each thread serves CPR rows
        const int q = r % CPR, rsub = r / CPR;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const Item w = decode_item(p, item);
            const ConvProblem& P = p.probs[w.pi];
            const int M = P.od * P.oh * P.ow;
            // everything the inner loop needs lives in registers: the asm memory clobbers of cp.async would otherwise make the
            // compiler re-read each field from parameter space (dependent LDC chains) for every tap
            const int nch0 = P.nch0, nch = P.nch0 + P.nch1, ntaps = P.ntaps;
            const int nsteps = ntaps * nch;
            const int ntiles = P.ntiles;
            const bool resident = p.resident_w != 0;
            const uint32_t bbytes = uint32_t(P.ntile) * kc * 2u;
            const uint8_t* wbase = static_cast<const uint8_t*>(P.wpack);
            const int in_d = P.in_d, in_h = P.in_h, in_w = P.in_w;
            const uint32_t pitch0 = uint32_t(P.c0p) * 2u, pitch1 = uint32_t(P.c1p) * 2u;
            const uint8_t* const s0 = static_cast<const uint8_t*>(P.src0) + P.coff0 * 2 + q * 16;
            const uint8_t* const s1 = static_cast<const uint8_t*>(P.src1) + P.coff1 * 2 + q * 16;
            uint32_t vmask[CPR];          // bit t: tap t of this row reads inside the volume
            long long vbase[CPR];
#pragma unroll
            for (int j = 0; j < CPR; ++j) {
                const int m = w.mt * 128 + rsub + j * RSTEP;
                const bool rv = m < M;
                int ox = 0, oy = 0, oz = 0;
                if (rv) {
                    ox = m % P.ow;
                    const int t = m / P.ow;
                    oy = t % P.oh;
                    oz = t / P.oh;
                }
                const int bz = oz * P.istride, by = oy * P.istride, bx = ox * P.istride;
                uint32_t vm = 0;
                if (rv) {
#pragma unroll 1
                    for (int t = 0; t < ntaps; ++t) {
                        const ConvTap tp = P.taps[t];
                        const bool ok = (unsigned)(bz + tp.dz) < (unsigned)in_d && (unsigned)(by + tp.dy) < (unsigned)in_h &&
                                        (unsigned)(bx + tp.dx) < (unsigned)in_w;
                        vm |= (ok ? 1u : 0u) << t;
                    }
                }
                vmask[j] = vm;
                vbase[j] = (long long)(bz * in_h + by) * in_w + bx;
            }
            int tap = 0, ch = 0;
            const int ngroups = (nsteps + spg - 1) / spg;
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g, ++it) {
                // one stage = up to `spg` consecutive K steps (taps x channel chunks): the barrier wait, the proxy fence and the
                // arrive are paid once per stage instead of once per 4 KB
                mbar_wait(empty_bar(stage), phase ^ 1, 0x100u | stage);
                const int cnt = min(spg, nsteps - g * spg);
                if (r == 0 && !resident) {
                    mbar_arrive_expect_tx(full_bar(stage), bbytes * cnt);
                    const uint32_t bdst = sB + stage * b_stage_bytes;
                    if (ntiles == 1)
                        bulk_g2s(bdst, wbase + size_t(g * spg) * bbytes, bbytes * cnt, full_bar(stage));
                    else
                        for (int j = 0; j < cnt; ++j)
                            bulk_g2s(bdst + j * bbytes, wbase + (size_t(g * spg + j) * ntiles + w.nt) * bbytes, bbytes, full_bar(stage));
                }
                uint32_t dst = sA + stage * a_stage_bytes + q * 2048u + rsub * 16u;
#pragma unroll 1
                for (int j = 0; j < cnt; ++j, dst += a_sub_bytes) {
                    const long long delta = P.tap_delta[tap];
                    const bool first = ch < nch0;
                    const uint32_t pitch = first ? pitch0 : pitch1;
                    const uint8_t* const sb = (first ? s0 : s1);
                    const int coff = (first ? ch : ch - nch0) * kc * 2;
#pragma unroll
                    for (int i = 0; i < CPR; ++i) {
                        const bool valid = (vmask[i] >> tap) & 1u;
                        const uint8_t* src = valid ? sb + (vbase[i] + delta) * pitch + coff : sb;
                        cp_async16(dst + i * (RSTEP * 16u), src, valid ? 16u : 0u);
                    }
                    if (++ch == nch) { ch = 0; ++tap; }
                }
                // completion is signalled by the copy engine itself: no wait and no proxy fence in the producer (a producer-side
                // fence.proxy.async waits for the in-flight copies and serialises every stage on a full memory latency)
                cp_async_mbar_arrive(full_bar(stage));
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        cp_async_wait<0>();
    } else if (warp == 8) {
        // ===================================== MMA issuer ====================================
        if (lane == 0) {
            // lean issue loop: every parameter in registers, descriptors advanced by 64-bit adds in 16-byte units
            // (tools/mma_bench.cu: ~45 clk per MMA this way vs ~200+ when descriptors are rebuilt per MMA)
            int stage = 0, phase = 0;
            uint32_t acc_cnt = 0;
            const int total_items = p.total_items, nstride = gridDim.x, ntile_max = p.ntile_max;
            const bool resident = p.resident_w != 0;
            const uint64_t a_desc0 = umma_smem_desc(sA, 2048u, 128u);
            const uint64_t a_stage_u = uint64_t(a_stage_bytes >> 4), a_sub_u = uint64_t(a_sub_bytes >> 4);
            const uint64_t b_stage_u = uint64_t(b_stage_bytes >> 4);
            for (int item = blockIdx.x; item < total_items; item += nstride, ++acc_cnt) {
                const Item w = decode_item(p, item);
                const ConvProblem& P = p.probs[w.pi];
                const int acc = acc_cnt & 1;
                mbar_wait(tempty_bar(acc), ((acc_cnt >> 1) & 1) ^ 1, 0x200u | acc);
                tc_fence_after();
                const int ntile = P.ntile;
                const uint32_t d_tmem = tmem_base + uint32_t(acc * ntile_max);
                const uint32_t idesc = umma_idesc(128, ntile, p.a_fmt, p.b_fmt, 0, 0);
                const int nsteps = P.ntaps * (P.nch0 + P.nch1);
                const uint32_t b_lbo = uint32_t(ntile) * 16u;
                const uint64_t b_sub_u = uint64_t((uint32_t(ntile) * kc * 2u) >> 4);   // one K step of weights
                const uint64_t b_k_u = uint64_t((2u * b_lbo) >> 4);                     // one K=16 slice inside it
                const uint64_t b_desc0 = umma_smem_desc(resident ? sA + p.off_w : sB, b_lbo, 128u);
                const int ngroups = (nsteps + spg - 1) / spg;
                bool first = true;
#pragma unroll 1
                for (int g = 0; g < ngroups; ++g) {
                    mbar_wait(full_bar(stage), phase, 0x300u | stage);
                    fence_proxy_async();   // producers' cp.async (generic proxy) writes -> visible to the MMA's async-proxy reads
                    tc_fence_after();
                    const int cnt = min(spg, nsteps - g * spg);
                    uint64_t ad = a_desc0 + uint64_t(stage) * a_stage_u;
                    uint64_t bd = resident ? b_desc0 + uint64_t(g * spg) * b_sub_u : b_desc0 + uint64_t(stage) * b_stage_u;
#pragma unroll 1
                    for (int j = 0; j < cnt; ++j, ad += a_sub_u, bd += b_sub_u) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
                            if (first) { umma_f16_first(d_tmem, ad + uint64_t(k) * 256u, bd + uint64_t(k) * b_k_u, idesc); first = false; }
                            else umma_f16_acc(d_tmem, ad + uint64_t(k) * 256u, bd + uint64_t(k) * b_k_u, idesc);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        const int r = threadIdx.x;  // 0..127 == TMEM lane
        uint32_t acc_cnt = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++acc_cnt) {
            const Item w = decode_item(p, item);
            const ConvProblem& P = p.probs[w.pi];
            const int M = P.od * P.oh * P.ow;
            const int m = w.mt * 128 + r;
            const bool rv = m < M;
            size_t vox = 0;
            if (rv && EPI != EPI_PLANAR32) {
                const int ox = m % P.ow;
                const int t = m / P.ow;
                const int oy = t % P.oh;
                const int oz = t / P.oh;
                vox = (size_t(oz * P.ostep + P.ooff_z) * P.OH + (oy * P.ostep + P.ooff_y)) * P.OW + (ox * P.ostep + P.ooff_x);
            }
            const int acc = acc_cnt & 1;
            mbar_wait(tfull_bar(acc), (acc_cnt >> 1) & 1, 0x400u | acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(acc * p.ntile_max);
#pragma unroll 1
            for (int c0 = 0; c0 < P.ntile; c0 += 16) {
                float v[16];
                tmem_ld16(t_row + c0, v);
                const int n0 = w.nt * P.ntile + c0;
                if (P.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < P.n_real) v[j] += __ldg(P.bias + n0 + j);
                }
                if constexpr (EPI == EPI_PLANAR32) {
                    if (rv) {
                        float* out = static_cast<float*>(P.dst);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < P.n_real) out[size_t(n0 + j) * M + m] = v[j];
                    }
                } else {
                    uint4* out = reinterpret_cast<uint4*>(static_cast<uint8_t*>(P.dst) +
                                                          (vox * P.dst_cp + P.dst_coff + n0) * 2);
                    if constexpr (EPI == EPI_ACCUM16) {
                        if (rv) {
                            uint4 o0 = out[0], o1 = out[1];
                            const uint32_t ow_[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 f = unpack2<false>(ow_[j]);
                                v[2 * j] += f.x;
                                v[2 * j + 1] += f.y;
                            }
                        }
                    }
                    if (rv) {
                        uint4 q0, q1;
                        q0.x = pack2<false>(v[0], v[1]);
                        q0.y = pack2<false>(v[2], v[3]);
                        q0.z = pack2<false>(v[4], v[5]);
                        q0.w = pack2<false>(v[6], v[7]);
                        q1.x = pack2<false>(v[8], v[9]);
                        q1.y = pack2<false>(v[10], v[11]);
                        q1.z = pack2<false>(v[12], v[13]);
                        q1.w = pack2<false>(v[14], v[15]);
                        out[0] = q0;
                        out[1] = q1;
                    }
                    if (EPI == EPI_STORE16 && p.stats != nullptr) {
                        // column sums over the 32 rows of this warp: recursive-halving butterfly, 16 shuffles per
                        // statistic, then one shared-memory atomic per column per warp
                        float a[16], q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            a[j] = rv ? v[j] : 0.f;
                            q[j] = a[j] * a[j];
                        }
                        halve_step<8, 16>(a, q, lane);
                        halve_step<4, 8>(a, q, lane);
                        halve_step<2, 4>(a, q, lane);
                        halve_step<1, 2>(a, q, lane);
                        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                        q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
                        if ((lane & 1) == 0) {
                            const int col = n0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                                            ((lane >> 1) & 1);
                            // one row per warp and one lane per column: plain adds, fixed order => deterministic
                            float* ws = sstats + warp * 2 * p.ntot_max;
                            ws[col] += a[0];
                            ws[p.ntot_max + col] += q[0];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
        if (p.stats != nullptr) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = r; i < 2 * p.ntot_max; i += 128)
                p.stats[size_t(blockIdx.x) * 2 * p.ntot_max + i] =
                    ((sstats[i] + sstats[2 * p.ntot_max + i]) + sstats[4 * p.ntot_max + i]) + sstats[6 * p.ntot_max + i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, p.tmem_cols);
}

int g_sm_count = 0;

template <int EPI, int LAG, int KC>
int launch_t(const KParams& kp, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<EPI, LAG, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_igemm_kernel<EPI, LAG, KC><<<grid, kThreads, smem, stream>>>(kp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int EPI, int KC>
int launch_kc(const KParams& kp, int grid, size_t smem, cudaStream_t stream) {
    if (kp.stages >= 8) return launch_t<EPI, 6, KC>(kp, grid, smem, stream);
    return launch_t<EPI, 2, KC>(kp, grid, smem, stream);
}

template <int EPI>
int launch_lag(const KParams& kp, int grid, size_t smem, cudaStream_t stream) {
    if (kp.kc == 16) return launch_kc<EPI, 16>(kp, grid, smem, stream);
    if (kp.kc == 32) return launch_kc<EPI, 32>(kp, grid, smem, stream);
    return launch_kc<EPI, 64>(kp, grid, smem, stream);
}

}  // namespace

int device_sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

unsigned int read_device_error() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    if (v) return v;
    v = read_device_error_wgrad();
    if (v) return v;
    v = read_device_error_tma();
    if (v) return v;
    v = read_device_error_band();
    if (v) return v;
    v = read_device_error_wband();
    if (v) return v;
    v = read_device_error_wquad();
    if (v) return v;
    return read_device_error_s2();
}

int conv_igemm_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, ConvProblem*, cudaStream_t stream) {
    if (probs.empty()) return 0;
    if (probs.size() > kMaxProb) {
        set_error("conv_igemm_launch: too many problems");
        return 1;
    }
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.nprob = int(probs.size());
    kp.kc = cfg.kc;
    kp.a_fmt = cfg.a_bf16;
    kp.b_fmt = cfg.b_bf16;
    int items = 0, ntile_max = 16, ntot_max = 16;
    for (size_t i = 0; i < probs.size(); ++i) {
        ConvProblem& P = kp.probs[i];
        P = probs[i];
        if (P.ntile % 16 || P.ntile < 16 || P.ntile > 256 || (cfg.kc != 16 && cfg.kc != 32 && cfg.kc != 64) || P.ntaps < 1 ||
            P.ntaps > 27 || P.nch0 + P.nch1 < 1) {
            set_error("conv_igemm_launch: bad problem shape");
            return 1;
        }
        if (P.shuffle_cp) {
            set_error("conv_igemm_launch: parity-stacked problems need the TMA kernel");
            return 1;
        }
        const long long M = 1LL * P.od * P.oh * P.ow;
        P.mtiles = int((M + 127) / 128);
        for (int t = 0; t < P.ntaps; ++t) P.tap_delta[t] = (P.taps[t].dz * P.in_h + P.taps[t].dy) * P.in_w + P.taps[t].dx;
        P.item_base = items;
        items += P.mtiles * P.ntiles;
        ntile_max = std::max(ntile_max, P.ntile);
        ntot_max = std::max(ntot_max, P.ntile * P.ntiles);
    }
    kp.total_items = items;
    kp.ntile_max = ntile_max;
    kp.ntot_max = ntot_max;
    int cols = 32;
    while (cols < 2 * ntile_max) cols <<= 1;
    kp.tmem_cols = cols;
    kp.spg = std::max(1, 64 / cfg.kc);
    size_t a_stage = size_t(128) * cfg.kc * 2 * kp.spg, b_stage = size_t(ntile_max) * cfg.kc * 2 * kp.spg;
    // single-problem, single-N-tile layers whose whole weight pack is small keep it resident in shared memory
    // (no per-K-step bulk copy: a 512-byte UBLKCP per step costs far more than the MMA it feeds)
    size_t w_bytes = 0;
    if (probs.size() == 1 && kp.probs[0].ntiles == 1) {
        const ConvProblem& P0 = kp.probs[0];
        w_bytes = size_t(P0.ntaps) * (P0.nch0 + P0.nch1) * P0.ntile * cfg.kc * 2;
        if (w_bytes > 96 * 1024) w_bytes = 0;
    }
    kp.resident_w = w_bytes ? 1 : 0;
    kp.w_bytes = uint32_t(w_bytes);
    if (w_bytes) b_stage = 0;
    const size_t fixed = size_t(8) * ntot_max * 4 + 8 * (2 * 32 + 4) + 16 + 256 + w_bytes;
    int stages = int((214 * 1024 - fixed) / (a_stage + b_stage));
    stages = std::min(stages, 32);
    if (stages < 3) {
        set_error("conv_igemm_launch: tile too large for the smem ring");
        return 1;
    }
    kp.stages = stages;
    kp.off_b = uint32_t(stages * a_stage);
    kp.off_w = uint32_t(kp.off_b + stages * b_stage);
    kp.off_stats = uint32_t(kp.off_w + w_bytes);
    kp.off_bars = uint32_t((kp.off_stats + 8 * ntot_max * 4 + 15) & ~15u);
    const size_t smem = kp.off_bars + 8 * (2 * stages + 4) + 16;
    kp.stats = (cfg.epi == EPI_STORE16 && probs.size() == 1) ? cfg.stats_partials : nullptr;
    const int grid = std::max(1, std::min(items, device_sm_count()));
    if (cfg.stats_grid_out) *cfg.stats_grid_out = grid;
    if (cfg.epi == EPI_PLANAR32) return launch_lag<EPI_PLANAR32>(kp, grid, smem, stream);
    if (cfg.epi == EPI_STORE16) return launch_lag<EPI_STORE16>(kp, grid, smem, stream);
    return launch_lag<EPI_ACCUM16>(kp, grid, smem, stream);
}

}  // namespace u3d
