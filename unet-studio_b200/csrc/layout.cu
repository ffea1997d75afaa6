// Layout kernels: reference NCDHW fp32 <-> internal NDHWC 16-bit (channels padded to 16), and the
// weight pack from the reference parameter layout (tensorN order/shape, /root/reference/main.cpp:193-204)
// into the canonical UMMA B-operand blobs consumed by conv_igemm.cu.
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "plan.h"

namespace u3d {
namespace {

// first-layer precision: the input as fp16 hi + lo pairs in the padded channels, [hi (C) | lo (C) | hi (C) | 0...]
__global__ void pack_act_split_kernel(const float* __restrict__ in, uint4* __restrict__ out, int C, int Cp, long long V) {
    const int nch = Cp / 8;
    for (long long vox = blockIdx.x * (long long)blockDim.x + threadIdx.x; vox < V; vox += (long long)gridDim.x * blockDim.x) {
        for (int chunk = 0; chunk < nch; ++chunk) {
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (chunk * 8 < 3 * C) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ch = chunk * 8 + j;
                    float v = 0.f;
                    if (ch < 3 * C) {
                        const int part = ch / C;
                        const float x = in[(long long)(ch - part * C) * V + vox];   // re-reads of the same address hit L1
                        const float hi = __half2float(__float2half_rn(x));
                        v = part == 1 ? x - hi : hi;
                    }
                    f[j] = v;
                }
                q.x = pack2<false>(f[0], f[1]);
                q.y = pack2<false>(f[2], f[3]);
                q.z = pack2<false>(f[4], f[5]);
                q.w = pack2<false>(f[6], f[7]);
            }
            out[vox * nch + chunk] = q;
        }
    }
}

// weight value of K index kk of source s for the pack kernels (PackDesc::split_k)
__device__ __forceinline__ bool pack_k_lookup(const PackDesc& d, int s, int kk, int& kidx, int& part) {
    part = 0;
    if (d.split_k > 0 && s == 0) {
        if (kk >= 3 * d.split_k) return false;
        part = kk / d.split_k;
        kidx = d.k_off[0] + kk % d.split_k;
        return true;
    }
    if (kk >= d.k_real[s]) return false;
    kidx = d.k_off[s] + kk;
    return true;
}
__device__ __forceinline__ float pack_k_value(float w, int part) {
    return part == 2 ? w - __half2float(__float2half_rn(w)) : w;
}

template <bool BF16>
__global__ void pack_act_kernel(const float* __restrict__ in, uint4* __restrict__ out, int C, int Cp, long long V) {
    // thread = voxel: reads its C planar fp32 values (coalesced per channel plane), writes the whole padded NDHWC row (Cp * 2 bytes,
    // contiguous across the threads of a warp); no 64-bit div/mod per element, zero chunks cost only the store
    const int nch = Cp / 8;
    for (long long vox = blockIdx.x * (long long)blockDim.x + threadIdx.x; vox < V; vox += (long long)gridDim.x * blockDim.x) {
        for (int chunk = 0; chunk < nch; ++chunk) {
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (chunk * 8 < C) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = chunk * 8 + j;
                    f[j] = c < C ? in[(long long)c * V + vox] : 0.f;
                }
                q.x = pack2<BF16>(f[0], f[1]);
                q.y = pack2<BF16>(f[2], f[3]);
                q.z = pack2<BF16>(f[4], f[5]);
                q.w = pack2<BF16>(f[6], f[7]);
            }
            out[vox * nch + chunk] = q;
        }
    }
}

template <bool BF16>
__global__ void unpack_act_kernel(const uint4* __restrict__ in, float* __restrict__ out, int C, int Cp, long long V) {
    const long long total = V * (Cp / 8);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long vox = i % V;
        const int chunk = int(i / V);
        const uint4 q = in[vox * (Cp / 8) + chunk];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<BF16>(w[j]);
            const int c = chunk * 8 + 2 * j;
            if (c < C) out[(long long)c * V + vox] = f.x;
            if (c + 1 < C) out[(long long)(c + 1) * V + vox] = f.y;
        }
    }
}

// generic blob: element i of [tap][chunk][ntile][kc/8][n][8]; (vblock, nvblocks) = this block's position in the job's virtual grid
__device__ __forceinline__ void pack_generic_body(const PackDesc& d, uint32_t vblock, uint32_t nvblocks) {
    // 32-bit index arithmetic (a pack has < 2^31 elements): the six div/mod per element dominated this kernel in 64 bit
    const uint32_t nch = uint32_t(d.nch[0] + d.nch[1]);
    const uint32_t ntile = uint32_t(d.ntile), ntiles = uint32_t(d.ntiles), kg_n = uint32_t(d.kc / 8);
    const uint32_t total = uint32_t(d.ntaps) * nch * ntiles * ntile * uint32_t(d.kc);
    for (uint32_t i = vblock * blockDim.x + threadIdx.x; i < total; i += nvblocks * blockDim.x) {
        uint32_t r = i;
        const int k8 = int(r & 7u); r >>= 3;
        const int n = int(r % ntile); r /= ntile;
        const int kg = int(r % kg_n); r /= kg_n;
        const int nt = int(r % ntiles); r /= ntiles;
        const int ch = int(r % nch); r /= nch;
        const int tap = int(r);
        const int s = ch < d.nch[0] ? 0 : 1;
        const int kk = (ch - (s ? d.nch[0] : 0)) * d.kc + kg * 8 + k8;
        int nn = nt * d.ntile + n;
        int ref = d.tap_ref[tap];
        if (d.stack_cp) {   // parity-stacked columns: the reference tap depends on the column block
            ref = d.stack_ref[tap][nn / d.stack_cp];
            nn = nn % d.stack_cp;
        }
        float v = 0.f;
        int kidx = 0, part = 0;
        if (ref >= 0 && nn < d.n_real && pack_k_lookup(d, s, kk, kidx, part)) {
            const int nidx = d.n_off + nn;
            const long long a = d.n_is_A ? nidx : kidx, b = d.n_is_A ? kidx : nidx;
            v = pack_k_value(d.w[(a * d.dimB + b) * d.ktaps + ref], part);
        }
        if (d.out_bf16)
            static_cast<__nv_bfloat16*>(d.out)[i] = __float2bfloat16_rn(v);
        else
            static_cast<__half*>(d.out)[i] = __float2half_rn(v);
    }
}

// banded blob of conv_band.cu: [9 (dz,dy)][KS][2 k-groups][NB = 4*CO columns][8] fp16, columns = kernel column kx = 2,1,0 then a zero
// block; the (dz,dy,dx) offsets come from the problem's own tap list (forward: k-1, dgrad: 1-k)
__device__ __forceinline__ void pack_band_body(const PackDesc& d, uint32_t vblock, uint32_t nvblocks) {
    const int KS = d.nch[0] + d.nch[1];
    const int NB = 4 * d.band_co;
    const uint32_t total = uint32_t(9 * KS * 2 * NB * 8);
    for (uint32_t i = vblock * blockDim.x + threadIdx.x; i < total; i += nvblocks * blockDim.x) {
        uint32_t r = i;
        const int k8 = int(r % 8); r /= 8;
        const int col = int(r % NB); r /= NB;
        const int kg = int(r % 2); r /= 2;
        const int ks = int(r % KS); r /= KS;
        const int t9 = int(r);
        const int blk = col / d.band_co, nn = col % d.band_co;
        float v = 0.f;
        if (blk < 3) {
            const int oz = t9 / 3 - 1, oy = t9 % 3 - 1, ox = (2 - blk) - 1;   // input offset of this block relative to the output voxel
            int tap = -1;
            for (int t = 0; t < 27; ++t)
                if (d.band_taps[t].dz == oz && d.band_taps[t].dy == oy && d.band_taps[t].dx == ox) tap = t;
            const int s = ks < d.nch[0] ? 0 : 1;
            const int kk = (ks - (s ? d.nch[0] : 0)) * 16 + kg * 8 + k8;
            int kidx = 0, part = 0;
            if (tap >= 0 && nn < d.n_real && pack_k_lookup(d, s, kk, kidx, part)) {
                const int nidx = d.n_off + nn;
                const long long a = d.n_is_A ? nidx : kidx, b = d.n_is_A ? kidx : nidx;
                v = pack_k_value(d.w[(a * d.dimB + b) * d.ktaps + d.tap_ref[tap]], part);
            }
        }
        static_cast<__half*>(d.out)[i] = __float2half_rn(v);
    }
}

// Tiled form of the generic blob for the plain (not parity-stacked, not split) case.  pack_generic_body reads one fp32 per element at
// a stride of ktaps floats (108 bytes for 3x3x3): every 4-byte read is its own 32-byte sector and the re-pack of the whole net spent
// ~200 us on that gather.  Here a block stages a CONTIGUOUS piece of the reference tensor in shared memory (coalesced reads) and
// writes 16-byte pieces [8 k] that are contiguous over n (coalesced writes):
//   n indexes dimA (forward conv weights [Cout][Cin][taps]):      tile = 4 n x (kc k x taps), contiguous per n
//   n indexes dimB (data-gradient packs, transposed conv weights): tile = 8 k x (32 n x taps), contiguous per k
constexpr int kPackTileFloats = 4 * 64 * 27;   // 27.6 KB of shared memory
__device__ __forceinline__ bool pack_tiled_ok(const PackDesc& d) {
    return !d.banded && d.stack_cp == 0 && d.split_k == 0 && !d.out_bf16 && d.kc * d.ktaps * 4 <= kPackTileFloats && 8 * 32 * d.ktaps <= kPackTileFloats &&
           d.ntile % 8 == 0;
}
__device__ __forceinline__ void pack_tiled_body(const PackDesc& d, uint32_t vblock, uint32_t nvblocks, float* sm) {
    const int nch = d.nch[0] + d.nch[1];
    const int kg_n = d.kc / 8;
    const int ktaps = d.ktaps;
    __half* const out = static_cast<__half*>(d.out);
    if (d.n_is_A) {
        const int n4s = d.ntile / 4;
        const int tiles = nch * d.ntiles * n4s;
        const int run = d.kc * ktaps;                  // floats per n
        for (int tile = int(vblock); tile < tiles; tile += int(nvblocks)) {
            const int n4 = tile % n4s;
            const int nt = (tile / n4s) % d.ntiles;
            const int ch = tile / (n4s * d.ntiles);
            const int s = ch < d.nch[0] ? 0 : 1;
            const int kk0 = (ch - (s ? d.nch[0] : 0)) * d.kc;
            const int kvalid = min(d.kc, d.k_real[s] - kk0);          // may be <= 0: an all-padding chunk
            __syncthreads();
            for (int e = threadIdx.x; e < 4 * run; e += blockDim.x) {
                const int i = e / run, r = e - i * run;
                const int nn = nt * d.ntile + n4 * 4 + i;
                float v = 0.f;
                if (nn < d.n_real && r < kvalid * ktaps) v = d.w[(size_t(d.n_off + nn) * d.dimB + d.k_off[s] + kk0) * ktaps + r];
                sm[e] = v;
            }
            __syncthreads();
            const int pieces = d.ntaps * kg_n * 4;
            for (int q = threadIdx.x; q < pieces; q += blockDim.x) {
                const int i = q & 3;
                const int kg = (q >> 2) % kg_n;
                const int tap = (q >> 2) / kg_n;
                const int ref = d.tap_ref[tap];
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = ref >= 0 ? sm[i * run + (kg * 8 + j) * ktaps + ref] : 0.f;
                uint4 o;
                o.x = pack2<false>(f[0], f[1]); o.y = pack2<false>(f[2], f[3]); o.z = pack2<false>(f[4], f[5]); o.w = pack2<false>(f[6], f[7]);
                const size_t idx = ((((size_t(tap) * nch + ch) * d.ntiles + nt) * kg_n + kg) * d.ntile + (n4 * 4 + i)) * 8;
                *reinterpret_cast<uint4*>(out + idx) = o;
            }
        }
    } else {
        const int n32s = (d.ntile + 31) / 32;
        const int tiles = nch * d.ntiles * kg_n * n32s;
        for (int tile = int(vblock); tile < tiles; tile += int(nvblocks)) {
            const int n32 = tile % n32s;
            const int kg = (tile / n32s) % kg_n;
            const int nt = (tile / (n32s * kg_n)) % d.ntiles;
            const int ch = tile / (n32s * kg_n * d.ntiles);
            const int s = ch < d.nch[0] ? 0 : 1;
            const int kk0 = (ch - (s ? d.nch[0] : 0)) * d.kc + kg * 8;
            const int nbase = nt * d.ntile + n32 * 32;                 // first n of this tile
            const int ncount = min(32, d.ntile - n32 * 32);
            const int run = 32 * ktaps;                                // floats per k (stride in shared memory)
            const int nvalid = min(ncount, d.n_real - nbase);          // may be <= 0
            __syncthreads();
            for (int e = threadIdx.x; e < 8 * ncount * ktaps; e += blockDim.x) {
                const int j = e / (ncount * ktaps), r = e - j * (ncount * ktaps);
                float v = 0.f;
                if (kk0 + j < d.k_real[s] && r < nvalid * ktaps) v = d.w[(size_t(d.k_off[s] + kk0 + j) * d.dimB + d.n_off + nbase) * ktaps + r];
                sm[j * run + r] = v;
            }
            __syncthreads();
            const int pieces = d.ntaps * ncount;
            for (int q = threadIdx.x; q < pieces; q += blockDim.x) {
                const int n = q % ncount;
                const int tap = q / ncount;
                const int ref = d.tap_ref[tap];
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = ref >= 0 ? sm[j * run + n * ktaps + ref] : 0.f;
                uint4 o;
                o.x = pack2<false>(f[0], f[1]); o.y = pack2<false>(f[2], f[3]); o.z = pack2<false>(f[4], f[5]); o.w = pack2<false>(f[6], f[7]);
                const size_t idx = ((((size_t(tap) * nch + ch) * d.ntiles + nt) * kg_n + kg) * d.ntile + (n32 * 32 + n)) * 8;
                *reinterpret_cast<uint4*>(out + idx) = o;
            }
        }
    }
}

__global__ void pack_weights_kernel(const __grid_constant__ PackDesc d) {
    __shared__ float sm[kPackTileFloats];
    if (d.banded) pack_band_body(d, blockIdx.x, gridDim.x);
    else if (pack_tiled_ok(d) && !d.force_elementwise) pack_tiled_body(d, blockIdx.x, gridDim.x, sm);
    else pack_generic_body(d, blockIdx.x, gridDim.x);
}

// every weight blob of a model in ONE launch (the re-pack after each optimizer step was 69 launches of a few microseconds of work
// each): job j owns the blocks [first_block[j], first_block[j+1])
__global__ void pack_all_kernel(const PackDesc* __restrict__ descs, const int* __restrict__ first_block, int njobs) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {   // last job whose first block is <= blockIdx.x
        const int mid = (lo + hi + 1) >> 1;
        if (first_block[mid] <= int(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    __shared__ float sm[kPackTileFloats];
    const PackDesc& d = descs[lo];
    const uint32_t vb = blockIdx.x - uint32_t(first_block[lo]), nvb = uint32_t(first_block[lo + 1] - first_block[lo]);
    if (d.banded) pack_band_body(d, vb, nvb);
    else if (pack_tiled_ok(d) && !d.force_elementwise) pack_tiled_body(d, vb, nvb, sm);
    else pack_generic_body(d, vb, nvb);
}

__global__ void trace_marker_kernel(int) {}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

bool pack_force_elementwise() {
    static const bool v = std::getenv("U3D_PACK_ELEMENTWISE") != nullptr;
    return v;
}

int pack_weights_launch(const PackDesc& din, cudaStream_t stream) {
    PackDesc d = din;
    d.force_elementwise = pack_force_elementwise() ? 1 : 0;
    const long long total = (long long)pack_bytes(d) / 2;
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(d);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// U3D_TRACE_LAUNCHES: an empty marker kernel in front of every traced launch site separates the layers in an ncu launch list
void trace_marker_launch(cudaStream_t stream) { trace_marker_kernel<<<1, 32, 0, stream>>>(0); }

int pack_job_blocks(const PackDesc& d) {
    const long long total = (long long)pack_bytes(d) / 2;
    const long long b = (total + 256 * 8 - 1) / (256 * 8);   // ~8 elements per thread
    return int(b < 1 ? 1 : (b > 4096 ? 4096 : b));
}

int pack_all_launch(const PackDesc* descs_dev, const int* first_block_dev, int njobs, int total_blocks, cudaStream_t stream) {
    if (njobs <= 0) return 0;
    pack_all_kernel<<<total_blocks, 256, 0, stream>>>(descs_dev, first_block_dev, njobs);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int pack_act_launch(const float* in, void* out, int C, int Cp, long long V, bool bf16, cudaStream_t stream, int split) {
    if (split) {
        if (bf16 || 3 * C > Cp || Cp > 64) { set_error("pack_act: split needs fp16 and 3*C <= Cp <= 64"); return 1; }
        pack_act_split_kernel<<<grid_for(V, 256), 256, 0, stream>>>(in, static_cast<uint4*>(out), C, Cp, V);
        U3D_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    if (bf16)
        pack_act_kernel<true><<<grid_for(V, 256), 256, 0, stream>>>(in, static_cast<uint4*>(out), C, Cp, V);
    else
        pack_act_kernel<false><<<grid_for(V, 256), 256, 0, stream>>>(in, static_cast<uint4*>(out), C, Cp, V);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int unpack_act_launch(const void* in, float* out, int C, int Cp, long long V, bool bf16, cudaStream_t stream) {
    const long long total = V * (Cp / 8);
    if (bf16)
        unpack_act_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), out, C, Cp, V);
    else
        unpack_act_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), out, C, Cp, V);
    U3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace u3d
