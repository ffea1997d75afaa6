// Kernel selection for one conv / weight-gradient problem set (model.cpp and the op-level API call only these).
//   forward + data gradient:  conv_band (k3 s1, <= 32 channels) -> conv_s2 (stride-2 forward, 16 input channels) -> conv_tma -> conv_igemm
//   weight gradient:          conv_wgrad_quad (k3 s1, 16 x 16 channels) -> conv_wgrad_band (k3 s1) per eligible problem, conv_wgrad
//                             (generic split-K) for the rest
// U3D_NO_BAND / U3D_NO_S2 / U3D_NO_TMA / U3D_NO_WQUAD / U3D_NO_WBAND take a family out (tests/test_conv_ops_gpu.py runs the fallbacks that way).
#include <cstdlib>

#include "u3d.h"

namespace u3d {

int conv_launch(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg, cudaStream_t stream) {
    if (conv_band_eligible(probs, cfg)) return conv_band_launch(probs, cfg, stream);
    if (!probs.empty() && probs[0].banded) { set_error("conv_launch: banded weight pack but the problem is not eligible for conv_band"); return 1; }
    if (conv_s2_eligible(probs, cfg)) return conv_s2_launch(probs, cfg, stream);
    if (conv_tma_eligible(probs, cfg)) return conv_tma_launch(probs, cfg, stream);
    return conv_igemm_launch(probs, cfg, nullptr, stream);
}

// profile family of the kernel conv_launch picks: 0 conv_igemm, 2 conv_s2, 4 conv_tma, 5 conv_band
int conv_kernel_kind(const std::vector<ConvProblem>& probs, const ConvLaunch& cfg) {
    if (conv_band_eligible(probs, cfg)) return 5;
    if (conv_s2_eligible(probs, cfg)) return 2;
    return conv_tma_eligible(probs, cfg) ? 4 : 0;
}

// Planner hint: a k3 s1 layer with K <= 64 and N <= 32 channels on a big volume is planned with 16-wide K chunks, the form the banded
// kernel takes (K = 64 is the 32 + 32 concat layer that conv_band runs as two passes).
bool conv_band_wants_kc16(int ks, int stride, int transposed, int k_channels_padded, int n_channels_padded, long long voxels) {
    static const bool disabled = std::getenv("U3D_NO_BAND") != nullptr;
    return !disabled && !transposed && ks == 3 && stride == 1 &&
           (k_channels_padded == 16 || k_channels_padded == 32 || k_channels_padded == 64) && n_channels_padded <= 32 && voxels >= 32768;
}

int conv_wgrad_dispatch(const std::vector<WgradProblem>& probs, const WgradLaunch& cfg, cudaStream_t stream, int* launches) {
    std::vector<WgradProblem> rest;
    int n = 0;
    static const bool block_atomics = std::getenv("U3D_WBAND_ATOMICS") != nullptr, tile_atomics = std::getenv("U3D_WGRAD_ATOMICS") != nullptr;
    const int per_block_launch = (cfg.partial_scratch != nullptr && !block_atomics) ? 2 : 1;   // main kernel + summing kernel
    for (const auto& P : probs) {
        if (conv_wgrad_quad_eligible(P)) {
            if (conv_wgrad_quad_launch(P, stream, cfg.partial_scratch, cfg.partial_scratch_bytes)) return 1;
            n += per_block_launch;
        } else if (conv_wgrad_band_eligible(P)) {
            if (conv_wgrad_band_launch(P, stream, cfg.partial_scratch, cfg.partial_scratch_bytes)) return 1;
            n += per_block_launch;
        } else
            rest.push_back(P);
    }
    if (!rest.empty()) {
        if (conv_wgrad_launch(rest, cfg, nullptr, stream)) return 1;
        n += (cfg.partial_scratch != nullptr && !tile_atomics) ? 2 : 1;   // (the summing kernel is skipped when no problem has enough K splits)
    }
    if (launches) *launches = n;
    return 0;
}

}  // namespace u3d
