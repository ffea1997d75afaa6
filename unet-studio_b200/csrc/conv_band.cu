// x-banded halo convolution (3x3x3, stride 1) on tcgen05 / TMEM for the full-resolution 16/32-channel layers that carry most
// of the U-Net's FLOPs (levels 0-1), forward and data gradient.
//
// Why: a tcgen05.mma with M = 128, K = 16 costs ~45 clk for every N <= 64 (tools/mma_bench.cu) because the 4 KB A operand
// read from shared memory, not the tensor pipe, is the limit.  With "voxels on M, Cout on N" a 16-channel layer uses N = 16:
// a quarter of what each A read could feed.  Here one M row is a GROUP of G consecutive output voxels along x and N = G*Cout
// (64): D[group][xo*Cout + co].  A tap (dz,dy) needs the G+2 input voxels xi = 0..G+1 of the group; input xi contributes to
// output xo through kernel column kx = xi - xo, so per (dz,dy,xi) ONE MMA multiplies the xi-th input voxel of every group
// (K = Cin) with the weight blocks [W(kx=2) | W(kx=1) | W(kx=0)] placed at the matching output columns.  Only the useful
// blocks are issued (N = Cout, 2Cout, 3Cout sub-ranges of the accumulator), so there is no zero-padding work on the tensor
// pipe and the weights are stored once ([9 (dz,dy)][K chunk][kx reversed][Cout] + one zero block for the first MMA).
// MMAs per G outputs per (dz,dy): G+2 instead of 3G  ->  G = 4: 54 instead of 108 per 128 rows of 4 voxels = 4x fewer A reads
// per voxel than conv_halo.cu.
//
// Shared-memory layout of one input z-plane of a tile ("slot"):   xs[cg][r][row p][8 ch]
//   cg = channel group of 8, r = hx mod G ("phase"), p = hy*HQ + hx div G.  For a fixed (cg, r) the rows are 16 B apart, so any
//   run of 128 consecutive p is a canonical SWIZZLE_NONE K-major operand (SBO = 128 B, LBO = distance between cg planes) and
//   every (dy, xi) is just a different start address:  p0 + dy*HQ + xi div G  in phase plane xi mod G.
// The CTA marches along z through a 4-slot ring of planes (3 in use, 1 loading): every input plane is fetched once per
// (x,y) tile column, there is no halo re-read along z at all.
//
// CTA = 448 threads: warps 0-3 epilogue, warps 4-11 producers (16-byte cp.async, zero fill = padding), warps 12-13 MMA issuers.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "plan.h"
#include "u3d.h"

namespace u3d {
namespace {

constexpr int kBThreads = 32 * 14;
constexpr int kBProducers = 256;
constexpr int kMaxSlots = 8;   // ring of input z-planes: 3 in use + (nslots - 3) in flight

struct BParams {
    ConvProblem P;
    int G, CO, KS;         // outputs per group, padded Cout, K chunks of 16
    int TX, TY, HX, HY, HQ, ROWS;
    int tiles_x, tiles_y, zchunks, zlen, total_items;
    int ncg;               // channel groups of 8 (both sources)
    int NB;                // columns of one weight block: 4*CO
    int nslots;
    uint32_t slot_bytes, w_bytes, off_w, off_stats, off_bars;
    float* stats;
    int epi;
};

template <int HALF, int BIT>
__device__ __forceinline__ void halve_step_b(float (&a)[16], float (&q)[16], int lane) {
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

// ACC = read-add-store epilogue (skip connections sum two data gradients): its own instantiation, because it prefetches the old
// row before waiting for the accumulator and the extra registers must not burden the store-only path
template <int G, int CO, int KS, bool ACC>
__global__ void __launch_bounds__(kBThreads, 1) conv_band_kernel(const __grid_constant__ BParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int N = G * CO;            // accumulator columns (64)
    constexpr int XI = G + 2;
    constexpr int NB = 4 * CO;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase + p.off_w;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sbase + p.off_bars;
    const uint32_t kSlots = uint32_t(p.nslots);
    // ring arithmetic without run-time division (a 32-bit divide is ~100 clk of dependent latency in front of every barrier wait and
    // descriptor): exact for c * kSlots < 2^32
    const uint32_t slot_magic = 0xFFFFFFFFu / kSlots + 1u;
    auto qdiv = [&](uint32_t c) { return __umulhi(c, slot_magic); };
    auto qmod = [&](uint32_t c) { return c - __umulhi(c, slot_magic) * kSlots; };
    auto full_bar = [&](uint32_t s) { return bars + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bars + 8u * (kMaxSlots + s); };
    // 4 TMEM accumulators: issuer wi alternates between accumulators wi and wi+2, so it can issue its next plane while the epilogue
    // still drains its previous one
    auto tfull_bar = [&](uint32_t a) { return bars + 8u * (2 * kMaxSlots + a); };
    auto tempty_bar = [&](uint32_t a) { return bars + 8u * (2 * kMaxSlots + 4 + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * kMaxSlots + 8);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * kMaxSlots + 9));

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kSlots; ++s) {
            mbar_init(full_bar(s), kBProducers);
            mbar_init(empty_bar(s), 2);
        }
        for (uint32_t a = 0; a < 4; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        mbar_init(wfull_bar, kBProducers);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 8 * CO + CO; i += kBThreads) sstats[i] = 0.f;
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 4 * N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const ConvProblem& P = p.P;
    const int D = P.in_d, H = P.in_h, W = P.in_w;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        {   // resident weights
            const uint8_t* wsrc = static_cast<const uint8_t*>(P.wpack);
            for (uint32_t o = t * 16u; o < p.w_bytes; o += kBProducers * 16u) cp_async16(sW + o, wsrc + o, 16u);
            cp_async_mbar_arrive(wfull_bar);
        }
        const int ncg0 = P.c0p / 8, ncg = p.ncg;
        const uint8_t* const s0 = static_cast<const uint8_t*>(P.src0);
        const uint8_t* const s1 = static_cast<const uint8_t*>(P.src1);
        const uint32_t pitch0 = uint32_t(P.c0p) * 2u, pitch1 = uint32_t(P.c1p) * 2u;
        const int HX = p.HX, HY = p.HY, HQ = p.HQ, ROWS = p.ROWS;
        const int per_plane = HY * HX * ncg;
        const uint32_t inv_hx = (1u << 20) / uint32_t(HX) + 1u;     // pos / HX == (pos * inv_hx) >> 20 for pos < HX*HY (no integer division in the loop)
        const int cg_shift = ncg == 2 ? 1 : 2;
        uint32_t cnt = 0;   // planes loaded by this CTA so far (ring position)
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int x0 = tx * p.TX - 1, y0 = ty * p.TY - 1;
            const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
            for (int gz = z0 - 1; gz <= z1; ++gz, ++cnt) {
                const uint32_t slot = qmod(cnt);
                mbar_wait(empty_bar(slot), ((qdiv(cnt)) & 1) ^ 1, 0x2100u | slot);
                const uint32_t blk = sbase + slot * p.slot_bytes;
                const bool zok = (unsigned)gz < (unsigned)D;
                // plane origin (gz, y0, x0); only offsets of in-range voxels are ever added to it
                const long long vox0 = ((long long)(zok ? gz : 0) * H + y0) * W + x0;
                const uint8_t* const p0 = s0 + vox0 * (long long)pitch0;
                const uint8_t* const p1 = s1 + vox0 * (long long)pitch1;
                // consecutive lanes: the 16-byte chunks of one voxel, then the next voxel along x (coalesced)
#pragma unroll 2
                for (int idx = t; idx < per_plane; idx += kBProducers) {
                    const int cg = idx & (ncg - 1);
                    const uint32_t pos = uint32_t(idx) >> cg_shift;
                    const int hy = int((pos * inv_hx) >> 20);
                    const int hx = int(pos) - hy * HX;
                    const bool ok = zok && (unsigned)(x0 + hx) < (unsigned)W && (unsigned)(y0 + hy) < (unsigned)H;
                    const int off = hy * W + hx;
                    const uint8_t* src;
                    if (cg < ncg0) src = ok ? p0 + (long long)off * pitch0 + cg * 16 : s0;
                    else src = ok ? p1 + (long long)off * pitch1 + (cg - ncg0) * 16 : s1;
                    const int r = hx & (G - 1), hq = hx / G;
                    cp_async16_ca(blk + uint32_t((cg * G + r) * ROWS + hy * HQ + hq) * 16u, src, ok ? 16u : 0u);
                }
                cp_async_mbar_arrive(full_bar(slot));
            }
        }
        cp_async_wait<0>();
    } else if (warp >= 12) {
        // ===================================== MMA issuers ===================================
        // Two issuing threads: issuer wi owns TMEM accumulator wi and every output plane whose running index has parity wi.  A single
        // thread sustains one MMA per ~87 clk in this loop (ncu: it never waits, it is issue-bound), the hardware accepts one per ~40.
        // Plane release protocol (empty barrier count 2 = one arrival per issuer): an issuer arrives on plane q after ITS last
        // output that reads q (outputs q-2, q-1, q of its parity); planes it never reads are released as soon as they are resident.
        // The WHOLE warp runs the loop: under `if (lane == 0)` the compiler cannot prove the descriptors uniform and feeds every
        // tcgen05.mma through an R2UR / vote loop; with uniform control flow they stay in uniform registers and one elected lane issues.
        const uint32_t wi = uint32_t(warp - 12);
        {
            const int HQ = p.HQ, ROWS = p.ROWS;
            const uint32_t lbo_a = uint32_t(G * ROWS) * 16u;             // next channel group of 8
            const uint64_t a_ks_u = uint64_t((2u * lbo_a) >> 4);         // next K chunk of 16 channels
            const uint32_t lbo_b = uint32_t(NB) * 16u;
            const uint64_t b_ks_u = uint64_t((2u * lbo_b) >> 4);
            const uint64_t b_desc0 = umma_smem_desc(sW, lbo_b, 128u);
            // per-xi constants: A offset (phase plane + group shift), B column offset, D column offset, N
            uint32_t a_xi[XI], b_xi[XI], d_xi[XI], i_xi[XI];
#pragma unroll
            for (int xi = 0; xi < XI; ++xi) {
                const int xo_lo = xi - 2 < 0 ? 0 : xi - 2, xo_hi = xi < G - 1 ? xi : G - 1;   // outputs fed by this input column
                const int nblk = xo_hi - xo_lo + 1;
                const int kx_hi = xi - xo_lo;                                                  // kernel column of the first block
                a_xi[xi] = uint32_t((xi % G) * ROWS + xi / G);                                 // 16-byte units
                b_xi[xi] = uint32_t((2 - kx_hi) * CO);                                         // blocks are stored kx = 2,1,0,zero
                d_xi[xi] = uint32_t(xo_lo * CO);
                i_xi[xi] = umma_idesc(128, nblk * CO, 0, 0, 0, 0);
            }
            const uint32_t idesc_full = umma_idesc(128, N, 0, 0, 0, 0);
            constexpr int XI0 = G - 2 >= 0 ? (G == 2 ? 1 : 2) : 0;   // the first MMA of a tile must write all N columns:
                                                                      // G = 4: xi = 2 with the zero block, G = 2: xi = 1
            mbar_wait(wfull_bar, 0, 0x2200u);
            fence_proxy_async();
            uint32_t cnt = 0, acc_cnt = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int zc = item % p.zchunks;
                const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
                const int nz = z1 - z0;
                // planes cnt, cnt+1 (relative z0-1, z0) must be resident before output plane 0; plane j+2 before output j
                mbar_wait(full_bar(qmod(cnt)), (qdiv(cnt)) & 1, 0x2300u);
                mbar_wait(full_bar(qmod(cnt + 1)), (qdiv(cnt + 1)) & 1, 0x2301u);
                if ((acc_cnt & 1u) != wi && lane == 0) {   // output 0 belongs to the other issuer: this one never reads plane 0 (nor 1 if nz == 1)
                    mbar_arrive(empty_bar(qmod(cnt)));
                    if (nz == 1) mbar_arrive(empty_bar(qmod(cnt + 1)));
                }
                __syncwarp();
                const uint32_t last_owner = (acc_cnt + uint32_t(nz - 1)) & 1u;
#pragma unroll 1
                for (int j = 0; j < nz; ++j, ++acc_cnt) {
                    if ((acc_cnt & 1u) != wi) continue;
                    const uint32_t c1 = cnt + j + 1, c2 = cnt + j + 2;
                    mbar_wait(full_bar(qmod(c1)), (qdiv(c1)) & 1, 0x2303u);
                    mbar_wait(full_bar(qmod(c2)), (qdiv(c2)) & 1, 0x2302u);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint32_t acc = acc_cnt & 3u;
                    mbar_wait(tempty_bar(acc), ((acc_cnt >> 2) & 1) ^ 1, 0x2400u | acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * uint32_t(N);
                    // row p0 = HQ (hy = 1, hq = 0) of each of the three planes
                    uint64_t a_pl[3];
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz)
                        a_pl[dz] = umma_smem_desc(sbase + (qmod(cnt + j + dz)) * p.slot_bytes + uint32_t(HQ) * 16u, lbo_a, 128u);
                    if (elect_one()) {
                    // first MMA: full width, overwrite
                    {
                        const uint64_t ad = a_pl[0] - uint64_t(HQ) + a_xi[XI0];
                        const uint64_t bd = b_desc0 + (G == 4 ? 0u : uint32_t(CO));
                        umma_f16_first(d_tmem, ad, bd, G == 4 ? idesc_full : i_xi[XI0]);
                    }
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint64_t a_row = a_pl[dz] + uint64_t((long long)(dy - 1) * HQ);
                            const uint64_t b_tap = b_desc0 + uint64_t((dz * 3 + dy) * KS) * b_ks_u;
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                                for (int xi = 0; xi < XI; ++xi) {
                                    if (dz == 0 && dy == 0 && ks == 0 && xi == XI0) continue;   // issued above
                                    umma_f16_acc(d_tmem + d_xi[xi], a_row + uint64_t(ks) * a_ks_u + a_xi[xi],
                                                 b_tap + uint64_t(ks) * b_ks_u + b_xi[xi], i_xi[xi]);
                                }
                            }
                        }
                    }
                    umma_commit(tfull_bar(acc));
                    umma_commit(empty_bar(qmod(cnt + j)));
                    umma_commit(empty_bar(qmod(c1)));
                    if (j >= nz - 2) umma_commit(empty_bar(qmod(c2)));   // this issuer has no later output reading plane j+2
                    }
                    __syncwarp();
                }
                if (last_owner != wi) {              // plane nz+1 is read by output nz-1 only
                    const uint32_t c = cnt + uint32_t(nz + 1);
                    mbar_wait(full_bar(qmod(c)), (qdiv(c)) & 1, 0x2304u);
                    if (lane == 0) mbar_arrive(empty_bar(qmod(c)));
                    __syncwarp();
                }
                cnt += uint32_t(nz + 2);
            }
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ======================================
        const int r = threadIdx.x;
        const int HQ = p.HQ;
        uint32_t acc_cnt = 0;
        float ssum[CO], ssq[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) ssum[j] = ssq[j] = 0.f;
        float* sbias = sstats + 8 * CO;
        for (int j = r; j < CO; j += 128) sbias[j] = (P.bias != nullptr && j < P.n_real) ? __ldg(P.bias + j) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool want_stats = p.stats != nullptr;
        constexpr bool accum = ACC;
        const int hy = 1 + r / HQ, hq = r % HQ;
        const bool row_in_tile = r < p.TY * HQ && hq < p.TX / G;
        uint8_t* const dst = static_cast<uint8_t*>(P.dst) + P.dst_coff * 2;
        const uint32_t dst_pitch = uint32_t(P.dst_cp) * 2u;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int gx0 = tx * p.TX + hq * G, gy = ty * p.TY + hy - 1;
            const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
            const bool rv_xy = row_in_tile && gy < H && gx0 < W;
#pragma unroll 1
            for (int gz = z0; gz < z1; ++gz, ++acc_cnt) {
                const size_t vox0 = (size_t(gz) * H + gy) * W + gx0;
                const uint32_t acc = acc_cnt & 3u;
                // read-add-store: the old row (G voxels x CO channels = 128 bytes) is fetched BEFORE the wait for the accumulator, so its
                // latency hides behind the MMAs instead of stalling the TMEM drain (141 -> 256 us per layer without this)
                uint4 oldv[ACC ? G * CO / 8 : 1];
                if constexpr (ACC) {
#pragma unroll
                    for (int q = 0; q < G * CO / 8; ++q) {
                        const int xo = q / (CO / 8);
                        const bool rvq = rv_xy && gx0 + xo < W;
                        oldv[q] = rvq ? *(reinterpret_cast<const uint4*>(dst + (vox0 + xo) * dst_pitch) + (q % (CO / 8))) : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                mbar_wait(tfull_bar(acc), (acc_cnt >> 2) & 1, 0x2500u | acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16) + acc * uint32_t(N);
#pragma unroll
                for (int xo = 0; xo < G; ++xo) {
                    const bool rv = rv_xy && gx0 + xo < W;
#pragma unroll
                    for (int c0 = 0; c0 < CO; c0 += 16) {
                        float v[16];
                        tmem_ld16(t_row + uint32_t(xo * CO + c0), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += sbias[c0 + j];
                        uint4* out = reinterpret_cast<uint4*>(dst + (vox0 + xo) * dst_pitch + c0 * 2);
                        if (accum && rv) {
                            const uint4 o0 = oldv[ACC ? xo * (CO / 8) + c0 / 8 : 0], o1 = oldv[ACC ? xo * (CO / 8) + c0 / 8 + 1 : 0];
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
#pragma unroll
            for (int c0 = 0; c0 < CO; c0 += 16) {
                float a[16], qq[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { a[j] = ssum[c0 + j]; qq[j] = ssq[c0 + j]; }
                halve_step_b<8, 16>(a, qq, lane);
                halve_step_b<4, 8>(a, qq, lane);
                halve_step_b<2, 4>(a, qq, lane);
                halve_step_b<1, 2>(a, qq, lane);
                a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
                if ((lane & 1) == 0) {
                    const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    float* ws = sstats + warp * 2 * CO;
                    ws[col] = a[0];
                    ws[CO + col] = qq[0];
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = r; i < 2 * CO; i += 128)
                p.stats[size_t(blockIdx.x) * 2 * CO + i] = ((sstats[i] + sstats[2 * CO + i]) + sstats[4 * CO + i]) + sstats[6 * CO + i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 4 * N);
}

// ---------------------------------------------------------------------------------------------------------------------------
// z-stacked variant for 16 -> 16 channels (G = 4, one K chunk).  Measured (tools/mma_ts_bench.cu): a tcgen05.mma costs >= 45 clk per
// SM whatever N <= 90 is, so the 54 narrow MMAs per plane tile of conv_band_kernel are dispatch-bound.  Here ONE MMA per
// (input plane, dy, x column) writes the accumulators of all three output planes that input plane feeds: N = 3 x 64 columns
// (output planes i-2, i-1, i <-> kz = 2, 1, 0) against a [dy][x column][2 k-groups][192][8] weight block expanded in shared memory
// from the banded blob — 18 instructions of 96 clk per plane tile instead of 54 of 45+.  An input plane is consumed by one batch
// of MMAs and released at once.  Eight accumulator slots of 64 columns (output q lives in slot q & 7; a batch that would run past
// slot 7 is split in two); an instruction has one accumulate predicate for all its columns, so the epilogue hands every slot back
// ZEROED (tcgen05.st) and every MMA accumulates.
// Status (round 1): bit-compatible with conv_band_kernel on the 78 operator cases, but SLOWER in situ (206 vs 156 us per layer at
// 160x192x160), so it is opt-in (U3D_ZBAND=1).  Timing experiments: no MMAs 129 us, no MMAs and no plane loads 94 us (epilogue alone),
// MMAs + epilogue without plane loads 191 us: the wide back-to-back MMAs (tensor pipe ~100 % busy, 96 clk each alone) and the
// epilogue's TMEM loads do not overlap — their times add — and a second epilogue warp group (one plane each) changes nothing.
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kZAcc = 8;
constexpr int kZThreads = 32 * 18;   // warps 0-3 + 14-17 epilogue, 4-11 producers, 12 MMA issuer (13 idle)
constexpr uint32_t kZWBytes = 3 * 6 * 2 * 192 * 16;   // 110592

__global__ void __launch_bounds__(kZThreads, 1) conv_zband_kernel(const __grid_constant__ BParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int G = 4, CO = 16, N = 64, XI = 6;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase + p.off_w;
    float* sstats = reinterpret_cast<float*>(smem + p.off_stats);
    const uint32_t bars = sbase + p.off_bars;
    const uint32_t kSlots = uint32_t(p.nslots);
    // ring arithmetic without run-time division (a 32-bit divide is ~100 clk of dependent latency in front of every barrier wait and
    // descriptor): exact for c * kSlots < 2^32
    const uint32_t slot_magic = 0xFFFFFFFFu / kSlots + 1u;
    auto qdiv = [&](uint32_t c) { return __umulhi(c, slot_magic); };
    auto qmod = [&](uint32_t c) { return c - __umulhi(c, slot_magic) * kSlots; };
    auto full_bar = [&](uint32_t s) { return bars + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bars + 8u * (kMaxSlots + s); };
    auto tfull_bar = [&](uint32_t a) { return bars + 8u * (2 * kMaxSlots + a); };
    auto tempty_bar = [&](uint32_t a) { return bars + 8u * (2 * kMaxSlots + kZAcc + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * kMaxSlots + 2 * kZAcc);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bars + 8u * (2 * kMaxSlots + 2 * kZAcc + 1));

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kSlots; ++s) {
            mbar_init(full_bar(s), kBProducers);
            mbar_init(empty_bar(s), 2);
        }
        for (uint32_t a = 0; a < kZAcc; ++a) {
            mbar_init(tfull_bar(a), 2);
            mbar_init(tempty_bar(a), 128);
        }
        mbar_init(wfull_bar, kBProducers);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 16 * CO; i += kZThreads) sstats[i] = 0.f;
    for (int j = threadIdx.x; j < CO; j += kZThreads) sstats[16 * CO + j] = (p.P.bias != nullptr && j < p.P.n_real) ? __ldg(p.P.bias + j) : 0.f;
    if (warp == 12) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp < 4) {   // all eight accumulator slots start zeroed
        const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 4
        for (int c = 0; c < 512; c += 16) tmem_st16_zero(t_row + uint32_t(c));
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const ConvProblem& P = p.P;
    const int D = P.in_d, H = P.in_h, W = P.in_w;

    if (warp >= 4 && warp < 12) {
        // ===================================== producers =====================================
        const int t = threadIdx.x - 128;
        {   // resident weights: expand the banded blob [9 (dz,dy)][2][kx = 2,1,0,zero][16][8] into [dy][xi][2][pl*64 + xo*16 + co][8]
            const uint4* wsrc = static_cast<const uint4*>(P.wpack);
            for (int n = t; n < 3 * XI * 2 * 192; n += kBProducers) {
                const int row = n % 192, kc = (n / 192) & 1, dyxi = n / 384;
                const int dy = dyxi / XI, xi = dyxi % XI;
                const int pl = row >> 6, xo = (row >> 4) & 3, co = row & 15;
                const int kx = xi - xo;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (kx >= 0 && kx <= 2) v = __ldg(wsrc + ((((2 - pl) * 3 + dy) * 2 + kc) * 64 + (2 - kx) * 16 + co));
                *reinterpret_cast<uint4*>(smem + p.off_w + uint32_t(n) * 16u) = v;
            }
            fence_proxy_async();
            mbar_arrive(wfull_bar);
        }
        const int ncg = p.ncg;
        const uint8_t* const s0 = static_cast<const uint8_t*>(P.src0);
        const uint32_t pitch0 = uint32_t(P.c0p) * 2u;
        const int HX = p.HX, HY = p.HY, HQ = p.HQ, ROWS = p.ROWS;
        const int per_plane = HY * HX * ncg;
        const uint32_t inv_hx = (1u << 20) / uint32_t(HX) + 1u;
        uint32_t cnt = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int x0 = tx * p.TX - 1, y0 = ty * p.TY - 1;
            const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
            for (int gz = z0 - 1; gz <= z1; ++gz, ++cnt) {
                const uint32_t slot = qmod(cnt);
                mbar_wait(empty_bar(slot), ((qdiv(cnt)) & 1) ^ 1, 0x2100u | slot);
                const uint32_t blk = sbase + slot * p.slot_bytes;
                const bool zok = (unsigned)gz < (unsigned)D;
                const long long vox0 = ((long long)(zok ? gz : 0) * H + y0) * W + x0;
                const uint8_t* const p0 = s0 + vox0 * (long long)pitch0;
#pragma unroll 2
                for (int idx = t; idx < per_plane; idx += kBProducers) {
                    const int cg = idx & 1;
                    const uint32_t pos = uint32_t(idx) >> 1;
                    const int hy = int((pos * inv_hx) >> 20);
                    const int hx = int(pos) - hy * HX;
                    const bool ok = zok && (unsigned)(x0 + hx) < (unsigned)W && (unsigned)(y0 + hy) < (unsigned)H;
                    const uint8_t* src = ok ? p0 + (long long)(hy * W + hx) * pitch0 + cg * 16 : s0;
                    const int r = hx & (G - 1), hq = hx / G;
                    cp_async16_ca(blk + uint32_t((cg * G + r) * ROWS + hy * HQ + hq) * 16u, src, ok ? 16u : 0u);
                }
                cp_async_mbar_arrive(full_bar(slot));
            }
        }
        cp_async_wait<0>();
    } else if (warp == 12 || warp == 13) {
        // ===================================== MMA issuers ===================================
        // two issuing threads share every batch (even / odd (dy, x column) pairs): one thread's descriptor arithmetic + issue costs
        // ~190 clk per MMA in this loop, twice the 96 clk the instruction takes.  Both accumulate into the same columns (addition
        // commutes; the tensor pipe executes the instructions one after the other) and both commit to every barrier (count 2).
        // The WHOLE warp runs the loop (uniform control flow keeps the descriptor arithmetic in uniform registers; under `if (lane == 0)`
        // every operand went through an R2UR chain, ~300 clk per MMA); one elected lane issues the tcgen05 instructions.
        const int wi = warp - 12;
        {
            const int HQ = p.HQ, ROWS = p.ROWS;
            const uint32_t lbo_a = uint32_t(G * ROWS) * 16u;
            const uint64_t b_desc0 = umma_smem_desc(sW, 192u * 16u, 128u);
            uint32_t a_xi[XI];
#pragma unroll
            for (int xi = 0; xi < XI; ++xi) a_xi[xi] = uint32_t((xi % G) * ROWS + xi / G);
            const uint32_t idesc_n[4] = {0u, umma_idesc(128, 64, 0, 0, 0, 0), umma_idesc(128, 128, 0, 0, 0, 0), umma_idesc(128, 192, 0, 0, 0, 0)};
            mbar_wait(wfull_bar, 0, 0x2200u);
            fence_proxy_async();
            uint32_t cnt = 0, q_base = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int zc = item % p.zchunks;
                const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
                const int nz = z1 - z0;
#pragma unroll 1
                for (int i = 0; i < nz + 2; ++i) {
                    const uint32_t c = cnt + uint32_t(i);
                    mbar_wait(full_bar(qmod(c)), (qdiv(c)) & 1, 0x2300u);
                    if (i < nz) {   // output plane i receives its first contribution from this input plane: its slot must be drained + zeroed
                        const uint32_t q = q_base + uint32_t(i);
                        mbar_wait(tempty_bar(q & 7u), ((q >> 3) & 1u) ^ 1u, 0x2400u | (q & 7u));
                    }
                    fence_proxy_async();
                    tc_fence_after();
                    const int jlo = i - 2 < 0 ? 0 : i - 2, jhi = i < nz - 1 ? i : nz - 1;
                    const int pl_lo = jlo - (i - 2), count = jhi - jlo + 1;
                    const uint32_t sl0 = (q_base + uint32_t(jlo)) & 7u;
                    const int n1 = count < int(8u - sl0) ? count : int(8u - sl0), n2 = count - n1;
                    const uint32_t d1 = tmem_base + sl0 * uint32_t(N), d2 = tmem_base;
                    const uint32_t id1 = idesc_n[n1], id2 = idesc_n[n2];
                    const uint64_t a_pl = umma_smem_desc(sbase + (qmod(c)) * p.slot_bytes + uint32_t(HQ) * 16u, lbo_a, 128u);
                    const uint64_t b_pl = b_desc0 + uint64_t(pl_lo * 64);        // 64 rows of 16 B per plane block (16-byte units)
                    const uint64_t b_pl2 = b_pl + uint64_t(n1 * 64);
                    if (elect_one()) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint64_t a_row = a_pl + uint64_t((long long)(dy - 1) * HQ);
#pragma unroll
                        for (int xi = 0; xi < XI; ++xi) {
                            if (((dy * XI + xi) & 1) != wi) continue;
                            const uint64_t boff = uint64_t((dy * XI + xi) * (2 * 192));
                            umma_f16_acc(d1, a_row + a_xi[xi], b_pl + boff, id1);
                            if (n2) umma_f16_acc(d2, a_row + a_xi[xi], b_pl2 + boff, id2);
                        }
                    }
                    umma_commit(empty_bar(qmod(c)));                                   // the plane is consumed by this batch alone
                    if (i >= 2) umma_commit(tfull_bar((q_base + uint32_t(i - 2)) & 7u));   // output i-2 has all 27 taps
                    }
                    __syncwarp();
                }
                cnt += uint32_t(nz + 2);
                q_base += uint32_t(nz);
            }
        }
        __syncwarp();
    } else if (warp < 4 || warp >= 14) {
        // ===================================== epilogue ======================================
        // two groups of four warps (0-3 and 14-17; a warp reaches the TMEM lanes of quarter warp % 4): group eg drains the output planes
        // of parity eg, so the per-plane latency chain (barrier wait, 4 x TMEM load, stores) of one plane overlaps the next plane's
        const uint32_t eg = warp < 4 ? 0u : 1u;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int HQ = p.HQ;
        uint32_t acc_cnt = 0;
        float ssum[CO], ssq[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) ssum[j] = ssq[j] = 0.f;
        const float* sbias = sstats + 16 * CO;
        const bool want_stats = p.stats != nullptr;
        const bool accum = p.epi == EPI_ACCUM16;
        const int hy = 1 + r / HQ, hq = r % HQ;
        const bool row_in_tile = r < p.TY * HQ && hq < p.TX / G;
        uint8_t* const dst = static_cast<uint8_t*>(P.dst) + P.dst_coff * 2;
        const uint32_t dst_pitch = uint32_t(P.dst_cp) * 2u;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int rem = item;
            const int zc = rem % p.zchunks; rem /= p.zchunks;
            const int tx = rem % p.tiles_x;
            const int ty = rem / p.tiles_x;
            const int gx0 = tx * p.TX + hq * G, gy = ty * p.TY + hy - 1;
            const int z0 = zc * p.zlen, z1 = min(D, z0 + p.zlen);
            const bool rv_xy = row_in_tile && gy < H && gx0 < W;
#pragma unroll 1
            for (int gz = z0; gz < z1; ++gz, ++acc_cnt) {
                if ((acc_cnt & 1u) != eg) continue;
                const size_t vox0 = (size_t(gz) * H + gy) * W + gx0;
                const uint32_t acc = acc_cnt & 7u;
                mbar_wait(tfull_bar(acc), (acc_cnt >> 3) & 1, 0x2500u | acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * uint32_t(N);
#pragma unroll
                for (int xo = 0; xo < G; ++xo) {
                    const bool rv = rv_xy && gx0 + xo < W;
                    float v[16];
                    tmem_ld16(t_row + uint32_t(xo * CO), v);
                    tmem_st16_zero(t_row + uint32_t(xo * CO));
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += sbias[j];
                    uint4* out = reinterpret_cast<uint4*>(dst + (vox0 + xo) * dst_pitch);
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
                                ssum[j] += v[j];
                                ssq[j] = fmaf(v[j], v[j], ssq[j]);
                            }
                        }
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
            }
        }
        if (want_stats) {
            float a[16], qq[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { a[j] = ssum[j]; qq[j] = ssq[j]; }
            halve_step_b<8, 16>(a, qq, lane);
            halve_step_b<4, 8>(a, qq, lane);
            halve_step_b<2, 4>(a, qq, lane);
            halve_step_b<1, 2>(a, qq, lane);
            a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
            qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
            if ((lane & 1) == 0) {
                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                float* ws = sstats + (eg * 4 + quarter) * 2 * CO;
                ws[col] = a[0];
                ws[CO + col] = qq[0];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (eg == 0)
                for (int i = r; i < 2 * CO; i += 128) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) t += sstats[w * 2 * CO + i];   // fixed order: deterministic
                    p.stats[size_t(blockIdx.x) * 2 * CO + i] = t;
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace

unsigned int read_device_error_band() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(v));
    return v;
}

size_t pack_bytes_band(const PackDesc& d) { return size_t(9) * (d.nch[0] + d.nch[1]) * 2 * (4 * d.band_co) * 8 * 2; }


// planner hook: k3 s1 layer with 16|32 padded input channels (both sources together) and 16|32 padded output channels
bool conv_band_wants(int k_channels_padded, int n_channels_padded, long long voxels) {
    static const bool disabled = std::getenv("U3D_NO_BAND") != nullptr;
    return !disabled && (k_channels_padded == 16 || k_channels_padded == 32) && (n_channels_padded == 16 || n_channels_padded == 32) &&
           voxels >= 32768;
}

bool conv_band_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    if (probs.empty() || !probs[0].banded) return false;
    if (probs.size() == 2) {   // two-pass split of a 32 + 32 channel concat layer
        if (probs[0].band_pass != 1 || probs[1].band_pass != 2 || cfg.epi != EPI_STORE16) return false;
        if (probs[0].c0p != 32 || probs[0].c1p != 32) return false;
    } else if (probs.size() != 1 || probs[0].band_pass != 0)
        return false;
    ConvProblem P = probs[0];
    if (P.band_pass) { P.c1p = 0; P.nch1 = 0; }
    if (cfg.kc != 16 || cfg.epi == EPI_PLANAR32 || cfg.a_bf16 || cfg.b_bf16) return false;
    if (P.ntaps != 27 || P.istride != 1 || P.ostep != 1 || P.ntiles != 1) return false;
    if (P.od != P.in_d || P.oh != P.in_h || P.ow != P.in_w || P.coff0 || P.coff1) return false;
    const int k = P.c0p + P.c1p;
    if ((k != 16 && k != 32) || (P.ntile != 16 && P.ntile != 32) || (P.nch0 + P.nch1) * 16 != k) return false;
    if (P.dst_cp % 8 || P.dst_coff % 8) return false;
    return true;
}

template <int G, int CO, int KS, bool ACC>
static int launch_band_ta(const BParams& bp, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_band_kernel<G, CO, KS, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_band_kernel<G, CO, KS, ACC><<<grid, kBThreads, smem, stream>>>(bp);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}
template <int G, int CO, int KS>
static int launch_band_t(const BParams& bp, int grid, size_t smem, cudaStream_t stream) {
    if (bp.epi == EPI_ACCUM16) return launch_band_ta<G, CO, KS, true>(bp, grid, smem, stream);
    return launch_band_ta<G, CO, KS, false>(bp, grid, smem, stream);
}

static int conv_band_launch_one(const ConvProblem& P, const ConvLaunch& cfg, bool stats, cudaStream_t stream);

int conv_band_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    if (probs.size() == 2) {
        ConvProblem A = probs[0], B = probs[1];
        A.c1p = 0; A.nch1 = 0; A.src1 = nullptr;
        B.src0 = B.src1; B.c0p = B.c1p; B.nch0 = B.nch1; B.c1p = 0; B.nch1 = 0; B.src1 = nullptr; B.bias = nullptr;
        ConvLaunch c1 = cfg, c2 = cfg;
        c1.epi = EPI_STORE16; c1.stats_partials = nullptr; c1.stats_grid_out = nullptr;
        c2.epi = EPI_ACCUM16;
        if (conv_band_launch_one(A, c1, false, stream)) return 1;
        return conv_band_launch_one(B, c2, cfg.stats_partials != nullptr, stream);
    }
    return conv_band_launch_one(probs[0], cfg, cfg.epi == EPI_STORE16 && cfg.stats_partials != nullptr, stream);
}

static int conv_band_launch_one(const ConvProblem& Pin, const ConvLaunch& cfg, bool stats, cudaStream_t stream) {
    BParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.P = Pin;
    const ConvProblem& P = bp.P;
    bp.CO = P.ntile;
    bp.G = bp.CO == 16 ? 4 : 2;
    bp.KS = (P.c0p + P.c1p) / 16;
    bp.ncg = (P.c0p + P.c1p) / 8;
    bp.NB = 4 * bp.CO;
    // tile: TX x TY outputs per z-plane with TY*HQ <= 128 rows; pick the (TX, TY) with the fewest plane tiles
    long long best = -1;
    for (int TX = 4 * bp.G; TX <= 64; TX += 4 * bp.G) {
        const int HQ = (TX + 2 + bp.G - 1) / bp.G;
        const int ty_max = 128 / HQ;
        if (ty_max < 1) continue;
        const int tiles_y0 = (P.in_h + ty_max - 1) / ty_max;
        const int TY = (P.in_h + tiles_y0 - 1) / tiles_y0;
        const long long tiles = 1LL * ((P.in_w + TX - 1) / TX) * tiles_y0;
        if (best < 0 || tiles < best) { best = tiles; bp.TX = TX; bp.TY = TY; }
    }
    bp.HX = bp.TX + 2; bp.HY = bp.TY + 2;
    bp.HQ = (bp.HX + bp.G - 1) / bp.G;
    bp.ROWS = (2 * bp.HQ + 129) | 1;   // odd: the 16-byte chunks of the ncg*G phase planes fall into different banks (conflict-free cp.async writes)
    bp.tiles_x = (P.in_w + bp.TX - 1) / bp.TX;
    bp.tiles_y = (P.in_h + bp.TY - 1) / bp.TY;
    // z chunks: enough items for >= ~4 waves of the SMs, at least 8 planes per chunk (2 halo planes are re-read per chunk)
    const int sms = device_sm_count();
    const int cols = bp.tiles_x * bp.tiles_y;
    int zchunks = std::max(1, std::min(P.in_d / 8, (4 * sms + cols - 1) / cols));
    {   // prefer a chunk count whose item total fills whole waves
        double best_eff = -1;
        int best_zc = zchunks;
        for (int zc = std::max(1, zchunks / 2); zc <= std::max(1, std::min(P.in_d / 4, zchunks * 2)); ++zc) {
            const int zl = (P.in_d + zc - 1) / zc;
            const int zc_eff = (P.in_d + zl - 1) / zl;
            const long long items = 1LL * cols * zc_eff;
            const long long waves = (items + sms - 1) / sms;
            const double eff = double(items) / double(waves * sms) * double(zl) / double(zl + 2);
            if (eff > best_eff) { best_eff = eff; best_zc = zc_eff; }
        }
        zchunks = best_zc;
    }
    bp.zlen = (P.in_d + zchunks - 1) / zchunks;
    bp.zchunks = (P.in_d + bp.zlen - 1) / bp.zlen;
    bp.total_items = cols * bp.zchunks;
    bp.slot_bytes = uint32_t(bp.ncg * bp.G * bp.ROWS * 16);
    // experimental, opt-in (U3D_ZBAND=1): parity green, but 206 us vs 156 us per 16 -> 16 layer at 160x192x160 (see the kernel's header)
    static const bool use_zband = std::getenv("U3D_ZBAND") != nullptr;
    const bool zband = use_zband && bp.CO == 16 && bp.KS == 1 && P.c1p == 0;
    bp.w_bytes = zband ? kZWBytes : uint32_t(9 * bp.KS * 2 * bp.NB * 16);
    const uint32_t stats_bytes = uint32_t((zband ? 17 : 9) * bp.CO * 4);
    bp.nslots = int(std::min<size_t>(kMaxSlots, (size_t(222) * 1024 - bp.w_bytes - stats_bytes) / bp.slot_bytes));
    if (bp.nslots < (zband ? 3 : 4)) { set_error("conv_band_launch: plane ring does not fit in shared memory"); return 1; }
    bp.off_w = bp.nslots * bp.slot_bytes;
    bp.off_stats = bp.off_w + bp.w_bytes;
    bp.off_bars = uint32_t((bp.off_stats + stats_bytes + 15) & ~15u);
    const size_t smem = bp.off_bars + 8 * (2 * kMaxSlots + 2 * kZAcc + 2) + 16;
    if (smem > 227 * 1024) { set_error("conv_band_launch: tile does not fit in shared memory"); return 1; }
    bp.epi = cfg.epi;
    bp.stats = stats ? cfg.stats_partials : nullptr;
    const int grid = std::max(1, std::min(bp.total_items, sms));
    if (cfg.stats_grid_out) *cfg.stats_grid_out = grid;
    if (zband) {
        static bool attr_set = false;
        if (!attr_set) {
            U3D_CUDA_CHECK(cudaFuncSetAttribute(conv_zband_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        conv_zband_kernel<<<grid, kZThreads, smem, stream>>>(bp);
        U3D_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    if (bp.CO == 16 && bp.KS == 1) return launch_band_t<4, 16, 1>(bp, grid, smem, stream);
    if (bp.CO == 16 && bp.KS == 2) return launch_band_t<4, 16, 2>(bp, grid, smem, stream);
    if (bp.CO == 32 && bp.KS == 1) return launch_band_t<2, 32, 1>(bp, grid, smem, stream);
    return launch_band_t<2, 32, 2>(bp, grid, smem, stream);
}

}  // namespace u3d
