// Host launchers of the bandwidth-bound kernels (elementwise.cu, loss.cu, optim.cu).
#pragma once
#include "u3d.h"

namespace u3d {

enum ActKind : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_ELU = 3 };

int reduce_rows_max();  // max partial rows any reduction writes (sizing of the partials scratch)

// per-channel (sum, sum of squares) partial rows of a [V][Cp] fp16 tensor
int channel_stats_launch(const void* x, long long V, int C, int Cp, float* partials, int* rows, cudaStream_t s);
// partial rows [rows][2][ntot] -> mean, rstd = 1/sqrt(var_biased + eps); optional BatchNorm running-stat update
int finalize_stats_launch(const float* partials, int rows, int ntot, int C, double count, float eps, float* mean, float* rstd,
                          float* running_mean, float* running_var, float momentum, cudaStream_t s);
// rstd[c] = 1/sqrt(var[c] + eps)  (BatchNorm3d eval mode with running statistics)
int rstd_from_var_launch(const float* var, float* rstd, int C, float eps, cudaStream_t s);
// y = act(gamma*(x-mean)*rstd + beta); mean/rstd may be null (affine only: BatchNorm eval after prepare_for_inference)
int norm_act_fwd_launch(const void* x, void* y, long long V, int C, int Cp, int has_norm, int act, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, cudaStream_t s);
// dx from dy through act and (instance|batch N=1) norm; dgamma/dbeta += (accumulated across micro-batches)
int norm_act_bwd_launch(const void* x, const void* dy, void* dx, long long V, int C, int Cp, int has_norm, int act,
                        const float* mean, const float* rstd, const float* gamma, const float* beta, float* partials,
                        float* sums, float* dgamma, float* dbeta, cudaStream_t s, unsigned int* counter = nullptr);
// out[c] += sum_v x[v][c]   (bias gradients)
int colsum_accumulate_launch(const void* x, long long V, int C, int Cp, float* partials, float* out, cudaStream_t s);

int maxpool_fwd_launch(const void* x, void* y, int* idx, int Cp, int od, int oh, int ow, cudaStream_t s);
int maxpool_bwd_launch(const void* dy, const int* idx, void* dx, int Cp, int od, int oh, int ow, cudaStream_t s);
int upsample_fwd_launch(const void* x, void* y, int Cp, int id, int ih, int iw, cudaStream_t s);
int upsample_bwd_launch(const void* dy, void* dx, int Cp, int id, int ih, int iw, cudaStream_t s);
int add16_launch(void* dst, const void* src, long long n_chunks, cudaStream_t s);

// ---- loss head (loss.cu): calc_losses of train.cpp:501-552 for one deep-supervision level, forward + gradient ----
struct LossLevel {
    const float* logits;   // fp32 planar [C][d*h*w] (reference NCDHW order)
    const float* label;    // float-stored integer labels at FULL resolution [D0][H0][W0] (train.cpp:615-617)
    void* dlogits;         // fp16 NDHWC [d*h*w][Cp] gradient (x loss_scale), or nullptr (validation)
    int C, Cp;
    int collapse_before;
    int d, h, w;           // this level's extent
    int H0, W0;            // full-resolution pitch of `label`
    int shift;             // level index k: label voxel = (z<<k, y<<k, x<<k)  (nearest interpolate, train.cpp:645-662)
    float w_ce, w_dice, w_mse;  // level weight (1/2^k)/sum x cost flags (train.cpp:686-699)
    float loss_scale;
    double* acc;           // device scratch: 3 + 2*32 doubles (totals, written by the finalize kernel)
    float* part;           // device scratch: [loss_part_rows()][3 + 2*32] per-block partial sums (fixed-order reduction => deterministic losses)
    float* out3;           // device: ce, dice, mse of this level
};
int loss_level_launch(const LossLevel& L, cudaStream_t s);
int loss_part_rows();     // rows of LossLevel::part
int loss_part_cols();

// 1x1 output head (Conv3d k1 of the `output` token, unet.cpp:186-187) on CUDA cores for head inputs of <= 32 channels:
// these levels are bandwidth-bound (2*C*xc FLOP per 32..64 bytes), so the tensor path only adds launches and round trips.
struct HeadFuse {
    const void* x;         // head input activations, fp16 NDHWC [v][xcp]
    int xc, xcp;
    const float* w;        // [C][xc] fp32 master weights (reference [C,xc,1,1,1])
    void* dx;              // gradient wrt x, fp16 [v][xcp] (x loss_scale)
    int dx_accum;          // 0 = store, 1 = read-add-store
    float* dw;             // [C][xc] fp32 gradient, accumulated atomically (x loss_scale)
    float* db;             // [C]
};
bool head_fwd_supported(int C, int xcp);
bool head_bwd_supported(int C, int xcp);
int head_fwd_launch(const void* x, int xc, int xcp, const float* w, const float* b, float* logits, int C, long long nv, cudaStream_t s,
                    const SrcTransform* xf = nullptr);
// loss of one level with the gradient pushed straight through the head (no dlogits tensor); Hd == nullptr = plain loss
int loss_level_launch(const LossLevel& L, const HeadFuse* Hd, cudaStream_t s);

// ---- optimizer (optim.cu): grad/batch, clip_grad_norm_(12), Nesterov SGD (train.cpp:759-766, unet.cpp:246-277) ----
struct SgdChunk {
    long long offset;
    int count;
    float weight_decay;
};
struct SgdStatus {      // device-resident result of one step
    double sumsq;       // squared global L2 norm of grad/batch (pre-clip)
    int nonfinite;      // overflow of the fp16 gradient path -> step skipped
    int pad;
};
int sgd_step_launch(float* params, float* grads, float* momentum, long long n, const SgdChunk* chunks_dev, int nchunks,
                    float inv_scale_batch, float lr, float mu, float max_norm, int first_step, SgdStatus* status_dev,
                    cudaStream_t s);

}  // namespace u3d
