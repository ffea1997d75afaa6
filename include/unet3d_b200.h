/* C-ABI of libunet3d_b200.so — the B200-native drop-in for UNet-Studio's libtorch/cuDNN model path.
 *
 * The reference has no FFI layer: its boundary is the C++ class UNet3dImpl (/root/reference/unet.hpp:13-70)
 * plus the training-step body (train.cpp:628-706,755-766), the inference window loop
 * (evaluate.cpp:223-230) and visual_perception_augmentation (train.hpp:43-48).  Every entry point below
 * cites the reference interface it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions: every function returns 0 on success, non-zero on failure; unet3d_last_error() then
 * returns a thread-local message (constructor errors carry the reference's std::runtime_error text,
 * unet.cpp:53,66,88,117).  Nothing throws or aborts.  Host tensors are fp32, NCDHW with x fastest
 * (train.cpp:619-621).  A handle owns its device memory, belongs to one GPU, is not internally locked;
 * distinct handles may be driven from distinct threads (train.cpp:592-600).  The library never keeps a
 * caller pointer past the call (train.cpp:615-621).
 */
#ifndef UNET3D_B200_H
#define UNET3D_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* unet3d_last_error(void);

/* ---- operator level (one reference layer, host buffers) — used by the per-layer parity tests ---------
 * Conv3d k1s1|k3s1|k3s2 pad (k-1)/2 (unet.cpp:59-72) or ConvTranspose3d k2s2 (unet.cpp:46-57), input given
 * as up to two tensors x0|x1 whose channel concat {x0,x1} (unet.cpp:181) is folded into the GEMM K loop.
 * weight/bias use the reference parameter layout ([Cout][Cin][k][k][k]; conv_trans [Cin][Cout][2][2][2]).
 * stats_sum_sumsq (optional, 2*cout doubles): per-channel sum and sum of squares of y, as produced for the
 * following InstanceNorm3d.  planar_fp32 != 0 selects the logits epilogue (fp32 NCDHW, no 16-bit rounding). */
int u3d_op_conv_forward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                        const float* x0, const float* x1, const float* weight, const float* bias, float* y,
                        double* stats_sum_sumsq, int planar_fp32);
/* autograd of the same layer (train.cpp:706): gx0/gx1 data gradients (NULL to skip), gw weight gradient. */
int u3d_op_conv_backward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                         const float* x0, const float* x1, const float* weight, const float* dy, float* gx0,
                         float* gx1, float* gw, int accumulate_gx0);

#ifdef __cplusplus
}
#endif
#endif
