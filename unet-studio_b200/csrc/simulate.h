// simulate_modality plan (host-drawn scalars) and device entry points (simulate.cu).
#pragma once
#include "u3d.h"

namespace u3d {

constexpr int kSimTerms = 20;     // train.cpp:50
constexpr int kSimMaxLut = 512;   // labels 0..511 (kernel-parameter resident LUT)

struct SimPlan {
    int W, H, D, labelled, n_lut;
    float lut[kSimMaxLut];
    uint8_t a[kSimTerms], b[kSimTerms], c[kSimTerms], d[kSimTerms];
    float w[kSimTerms];
    float gamma;
};

int simulate_make_plan(int labelled, unsigned max_label, unsigned seed, int W, int H, int D, SimPlan& plan);
size_t simulate_workspace_bytes(int W, int H, int D);
int simulate_run(const SimPlan& plan, float* t1w_dev, const float* label_dev, void* workspace, cudaStream_t s, long long* launches);

}  // namespace u3d
