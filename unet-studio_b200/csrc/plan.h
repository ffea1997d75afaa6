// Host-side planner: turns one reference layer (Conv3d / ConvTranspose3d, unet.cpp:46-72) into the
// gather-GEMM problems of conv_igemm.cu / conv_wgrad.cu and the weight-pack descriptors of layout.cu.
#pragma once
#include "u3d.h"

namespace u3d {

inline int pad16(int c) { return (c + 15) / 16 * 16; }

// Generic weight pack (layout.cu): reference fp32 tensor [dimA][dimB][ktaps] -> 16-bit blobs
// [tap][chunk][ntile][kc/8][ntile rows][8]  (canonical K-major SWIZZLE_NONE B operand).
struct PackDesc {
    const float* w;
    int dimA, dimB, ktaps;
    int n_is_A;              // 1: output channel n indexes dimA, K indexes dimB; 0: the other way round
    int n_off, n_real, ntile, ntiles;
    int k_off[2], k_real[2], nch[2];
    int kc;
    int ntaps;
    int tap_ref[27];
    void* out;
    int out_bf16;
    int stack_cp;            // > 0: N = 8 parity blocks of stack_cp columns; block `par` of GEMM tap t reads reference tap stack_ref[t][par] (-1 = zero)
    signed char stack_ref[8][8];
    int banded;              // 1: x-banded layout of conv_band.cu ([9 (dz,dy)][K chunk][kx 2,1,0,zero][band_co])
    int band_co;
    ConvTap band_taps[27];   // the problem's tap offsets (which reference tap each (dz,dy,dx) offset reads)
    int force_elementwise;   // debug (U3D_PACK_ELEMENTWISE): one thread per blob element instead of the shared-memory tiled form
    int split_k;             // > 0 (first layer of the network, source 0 only): K channel kk = g*split_k + c carries reference input channel c;
                             // g = 0, 1: fp16(w), g = 2: fp16(w - fp16(w)).  With the input packed as [hi | lo | hi] (pack_act_launch split = 1) the
                             // layer computes w_hi*x_hi + w_hi*x_lo + w_lo*x_hi = w*x to ~22 bits in the spare padded channels
};
size_t pack_bytes_band(const PackDesc& d);
size_t pack_bytes(const PackDesc& d);
int pack_weights_launch(const PackDesc& d, cudaStream_t stream);
// all blobs of a model in one launch: descs_dev[njobs], first_block_dev[njobs + 1] (prefix sums of pack_job_blocks)
void trace_marker_launch(cudaStream_t stream);
bool pack_force_elementwise();
int pack_job_blocks(const PackDesc& d);
int pack_all_launch(const PackDesc* descs_dev, const int* first_block_dev, int njobs, int total_blocks, cudaStream_t stream);

// NCDHW fp32 (reference order, train.cpp:619-621) <-> NDHWC 16-bit with channels zero-padded to Cp
// split != 0 (needs 3*C <= Cp): channels [0,C) = fp16(x), [C,2C) = fp16(x - fp16(x)), [2C,3C) = fp16(x) again (see PackDesc::split_k)
int pack_act_launch(const float* in, void* out, int C, int Cp, long long V, bool bf16, cudaStream_t stream, int split = 0);
int unpack_act_launch(const void* in, float* out, int C, int Cp, long long V, bool bf16, cudaStream_t stream);

struct LayerGeom {
    int transposed;          // 0 = Conv3d (k1 s1 | k3 s1 | k3 s2, pad (k-1)/2), 1 = ConvTranspose3d k2 s2
    int ks, stride;
    int cin[2];              // real input channels per source (cin[1] = 0 unless the input is a folded concat)
    int cout;
    int in_d, in_h, in_w;
    int out_d, out_h, out_w;
};

int choose_kc(int c0p, int c1p);
void choose_ntile(int np, long long m_voxels, int& ntile, int& ntiles);

// Forward: fills `probs` (operand/destination pointers left null) and `packs` (one per problem, w/out null).
void plan_forward(const LayerGeom& g, std::vector<ConvProblem>& probs, std::vector<PackDesc>& packs, int& kc, int force_kc = 0);
// Data gradient wrt source `src` (0/1): dy (cout channels, output extent) -> dx (cin[src] channels, input extent).
void plan_dgrad(const LayerGeom& g, int src, std::vector<ConvProblem>& probs, std::vector<PackDesc>& packs, int& kc, int force_kc = 0);
// Weight gradient wrt the channels of source `src`.
void plan_wgrad(const LayerGeom& g, int src, WgradProblem& prob);

}  // namespace u3d
