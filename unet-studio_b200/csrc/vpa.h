// visual_perception_augmentation plan (host-drawn scalars) and device entry points (vpa.cu).
#pragma once
#include "u3d.h"

namespace u3d {

constexpr int kVpaMaxC = 8;
constexpr int kVpaMaxFoci = 12;

struct VpaFocus {
    int loc[3];
    int ri;
    float radius, mag, coef, pir;
};

struct VpaPlan {
    int W, H, D, C, is_label;
    uint32_t seed;
    int ds, lw, lh, ld;
    int crop, crop_loc[3], crop_r;
    float crop_value;
    int trunc, top, bot;
    int noise;
    int noise_mt;          // library option noise_mt19937: the reference CPU path's sequential std::mt19937 stream (bit-exact) instead of the hash
    float noise_mag;
    int ambient;
    float ambient_add;
    int diffuse;
    float diff_f[3];
    int specular, spec_loc[3];
    float spec_freq, spec_mag, spec_b;
    float M[12];
    int has_persp;
    float persp[3];
    int use_disp, has_lens;
    float lens_k;
    int nfoci;
    VpaFocus foci[kVpaMaxFoci];
    int zero_bg, rubber, perlin, final_norm;
    float rubberM[5][12];
    float rubber_upper[kVpaMaxC][5];
    float zoom, perlin_upper;
    int perm[512];
};

int vpa_make_plan(const char* const* keys, const float* vals, int n_opts, int is_label, int W, int H, int D, int C, uint64_t seed,
                  VpaPlan& plan);
size_t vpa_workspace_bytes(int W, int H, int D, int C);
int vpa_run(const VpaPlan& plan, float* image_dev, float* label_dev, void* workspace, cudaStream_t s, long long* launches);

}  // namespace u3d
