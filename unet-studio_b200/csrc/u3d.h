// Internal host-side declarations shared by the .cu/.cpp files of libunet3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace u3d {

void set_error(const std::string& msg);  // thread-local last error (capi.cpp)
const char* last_error();

// ------------------------------------------------------------------------------------------
// Implicit-GEMM gather convolution (conv_igemm.cu).  One "problem" computes, for every voxel o of an
// output lattice (od x oh x ow) and every output channel n:
//     acc[o][n] = sum_{t < ntaps} sum_{k} src[(o*istride + tap_t)][k] * W_t[n][k]
// where src voxels outside [0,in_*) read as zero.  The result goes to the voxel
// (o*ostep + ooff) of the destination tensor.  This one form covers:
//   conv k3 s1/s2 p1, conv k1           istride = stride, taps = {-1,0,1}^3 or {0}
//   dgrad of conv k3 s1                 istride = 1, flipped taps, W transposed
//   dgrad of conv k3 s2                 8 output-parity problems with 1/2/4/8 taps
//   conv_transpose k2 s2 forward        8 output-parity problems, 1 tap each (ostep = 2)
//   dgrad of conv_transpose k2 s2       istride = 2, taps = {0,1}^3
// The channel concat of the decoder (unet.cpp:181) never materialises: K runs over src0's channels
// then src1's.
// ------------------------------------------------------------------------------------------
struct ConvTap {
    int8_t dz, dy, dx, pad;
};

enum EpiMode : int {
    EPI_STORE16 = 0,    // 16-bit NDHWC store (+bias) (+column sum / sum-of-squares partials)
    EPI_ACCUM16 = 1,    // 16-bit NDHWC read-add-store
    EPI_PLANAR32 = 2,   // fp32 planar [n][lattice voxel] store (+bias): logits in reference NCDHW order
};

// Optional on-the-fly normalisation + activation of the input of head_fwd_kernel (loss.cu): the pointer it gets is the RAW output of
// the last conv and it applies y = act(gamma*rstd*(x - mean) + beta) itself, exactly the arithmetic of norm_act_fwd_kernel
// (unet.cpp:74-98) -- that kernel's HBM round trip over the full-resolution tensor disappears.
struct SrcTransform {
    int enabled;
    int C;                 // real channels of the source (padded channels stay 0)
    int has_norm, act;     // as norm_act_fwd_launch
    const float* mean;     // nullptr = 0
    const float* rstd;     // nullptr = 1
    const float* gamma;
    const float* beta;
    void* writeback;       // training: the activated voxels are also stored (same pitch) for the fused head backward
};

struct ConvProblem {
    const void* src0;
    const void* src1;
    int c0p, c1p;          // channel pitch (elements) of each source tensor
    int coff0, coff1;      // first channel used in each source
    int nch0, nch1;        // number of kc-wide K chunks taken from each source
    int in_d, in_h, in_w;  // source spatial extent
    int istride;
    int ntaps;
    ConvTap taps[27];
    int od, oh, ow;        // lattice extent; M = od*oh*ow
    void* dst;
    int dst_cp;            // destination channel pitch (EPI_*16) ; unused for planar
    int dst_coff;          // first destination channel
    int OD, OH, OW;        // destination tensor extent
    int ostep, ooff_z, ooff_y, ooff_x;
    int ntile;             // N per tile (multiple of 16, <= 256)
    int ntiles;            // number of N tiles
    int n_real;            // real output channels (<= ntile*ntiles)
    const void* wpack;     // blobs [tap][chunk][ntile] of ntile x kc 16-bit, canonical K-major interleaved
    const float* bias;     // n_real entries or nullptr
    int mtiles;            // ceil(M/128)              (filled by the launcher)
    int tap_delta[27];     // (dz*in_h + dy)*in_w + dx  (filled by the launcher)
    int item_base;         // first work item           (filled by the launcher)
    int banded;            // planner: weights packed for the x-banded halo kernel (conv_band.cu)
    int band_pass;         // 0 = whole layer; 1 / 2 = a 32+32-channel concat layer split in two launches: source 0 (store), then source 1
                           // (read-add-store + statistics) -- the resident weights of K = 64 do not fit next to the plane ring
    int shuffle_cp;        // > 0: the N columns are 8 output-parity blocks of shuffle_cp channels; block (pz,py,px) of lattice voxel o goes
                           // to destination voxel 2*o + (pz,py,px) (conv_transpose k2 s2 forward, data gradient of conv k3 s2): conv_tma.cu only
    int shuffle_nreal;     // real channels per parity block
};

struct ConvLaunch {
    int kc;                // K chunk (16, 32 or 64 channels)
    int a_bf16, b_bf16;    // operand formats (0 = fp16, 1 = bf16)
    int out_bf16;          // 16-bit output format
    EpiMode epi;
    float* stats_partials; // [grid][2][ntile*ntiles] or nullptr (single-problem launches only)
    int* stats_grid_out;   // host pointer: receives the grid size used (number of partial rows)
    float* splitk_scratch; // optional fp32 scratch for the deterministic split-K of conv_tma (deep levels); nullptr = no split-K
    size_t splitk_scratch_bytes;
};

int conv_igemm_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, ConvProblem* dev_scratch,
                      cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// Weight-gradient GEMM (conv_wgrad.cu).  For every tap t of a problem:
//     dW[nch][mch][t] += sum_{v in lattice} T[(v*tstride + tap_t)][mch] * U[v][nch]
// T = "tapped" tensor (M side, up to 128 rows = tap-group x channels), U = "untapped" tensor (N side).
//   conv:        T = layer input x (fp16), U = dy (bf16), dW is [Cout][Cin][k^3]
//   conv_trans:  T = dy (bf16), U = layer input x (fp16), dW is [Cin][Cout][2^3]
// ------------------------------------------------------------------------------------------
struct WgradProblem {
    const void* T;
    const void* U;
    int t_cp, t_coff, t_c;      // pitch, first channel, padded channel count (multiple of 16) on the M side
    int t_creal;                // real channels on the M side
    int t_d, t_h, t_w;          // tapped tensor extent
    int tstride;
    int ntaps;
    ConvTap taps[27];
    int tap_ref[27];            // index of each tap in the reference weight's trailing k^3 dims
    int ld, lh, lw;             // lattice = untapped tensor extent (K = ld*lh*lw)
    int u_cp, u_coff, u_c;      // pitch, first channel, padded channel count on the N side
    int u_creal;
    float* dw;                  // reference-layout gradient, element (n, m, tap) at (n*w_mtot + w_moff + m)*w_ktaps + tap_ref
    int w_mtot, w_moff, w_ktaps;
    int w_ntot, w_noff;         // N-side channel offset inside the reference tensor's leading dim
    int tg;                     // taps per M tile   (filled by launcher)
    int mtiles;                 // M tiles          (filled by launcher)
    int ntile, ntiles;          // N tiling         (filled by launcher)
    int ksplit;                 // K splits         (filled by launcher)
    int item_base;
};
struct WgradLaunch {
    int t_bf16, u_bf16;
    float* partial_scratch;        // optional: conv_wgrad_band writes per-CTA gradient blocks here and a second kernel sums them into the
    size_t partial_scratch_bytes;  // gradient (an order of magnitude fewer fp32 atomics); nullptr = atomics straight from the accumulators
};
int conv_wgrad_launch(const std::vector<WgradProblem>& probs, const WgradLaunch& cfg, WgradProblem* dev_scratch,
                      cudaStream_t stream);

int conv_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream);  // band, s2, TMA or gather kernel (dispatch.cpp)
int conv_kernel_kind(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg);  // profile family conv_launch will use: 0 igemm, 2 s2, 4 tma, 5 band
bool conv_band_wants(int k_channels_padded, int n_channels_padded, long long voxels);
bool conv_s2_wants_kc16(int ks, int stride, int transposed, int cin_padded, int n_sources, int cout_padded, long long out_voxels);
bool conv_s2_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg);
int conv_s2_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream);
unsigned int read_device_error_s2();
bool conv_band_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg);
int conv_band_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream);
unsigned int read_device_error_band();
bool conv_tma_available();   // the driver exposes cuTensorMapEncodeTiled and the kernel is not disabled
bool conv_tma_eligible(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg);
int conv_tma_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream);
unsigned int read_device_error_tma();
bool conv_band_wants_kc16(int ks, int stride, int transposed, int k_channels_padded, int n_channels_padded, long long voxels);

bool conv_wgrad_quad_eligible(const WgradProblem& P);   // 16 x 16 channels, k3 s1, big volume: 2 x 2 g rows per MMA pair (conv_wgrad_quad.cu)
int conv_wgrad_quad_launch(const WgradProblem& P, cudaStream_t stream, float* partial_scratch = nullptr, size_t partial_scratch_bytes = 0);
// sums per-CTA gradient blocks [co_grp][ci_grp][27] (CTA index = rank * npairs + pair, pair = gi * ngo + go) into P.dw (conv_wgrad_band.cu)
int wgrad_block_sum_launch(const float* scratch, const WgradProblem& P, int npairs, int ngo, int nranks, int co_grp, int ci_grp, cudaStream_t stream);
unsigned int read_device_error_wquad();
bool conv_wgrad_band_eligible(const WgradProblem& P);
int conv_wgrad_band_launch(const WgradProblem& P, cudaStream_t stream, float* partial_scratch = nullptr, size_t partial_scratch_bytes = 0);
size_t conv_wgrad_band_scratch_bytes();   // enough for any problem
unsigned int read_device_error_wband();
// N-stacked band kernel per eligible problem, generic kernel for the rest; *launches = kernels launched
int conv_wgrad_dispatch(const std::vector<WgradProblem>& probs, const WgradLaunch& cfg, cudaStream_t stream, int* launches);

int device_sm_count();
unsigned int read_device_error();  // first non-zero mbarrier-timeout code of any kernel TU (0 = ok)
unsigned int read_device_error_wgrad();

}  // namespace u3d
