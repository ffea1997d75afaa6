// Inference pre/post-processing launchers (postproc.cu).
#pragma once
#include <stdint.h>

#include <vector>

#include "u3d.h"

namespace u3d {

constexpr int kPostMaxC = 32;

// window start positions along one axis: 0, stride, 2*stride, ... while the window ends inside the volume, then the window that ends
// at the border; a volume not larger than the window has the single origin 0 (zero padding at the far end)
std::vector<int> window_origins(int vdim, int wdim, int stride);

int crop_window_launch(const float* vol, float* win, int C, int vw, int vh, int vd, int ww, int wh, int wd, int ox, int oy, int oz, cudaStream_t s);
int softmax_accumulate_launch(const float* logits, float* acc, float* cnt, int C, int vw, int vh, int vd, int ww, int wh, int wd, int ox, int oy,
                              int oz, cudaStream_t s);
int mask_argmax_launch(float* acc, const float* cnt, uint8_t* label, float* fg, int C, long long V, float threshold, int write_prob, cudaStream_t s);
int resample_launch(const float* src, float* dst, int C, int sw, int sh, int sd, int dw, int dh, int dd, int nearest, cudaStream_t s);

}  // namespace u3d
