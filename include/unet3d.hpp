// C++ drop-in surface over the C-ABI (unet3d_b200.h): same member names and call shapes as the reference's
// UNet3dImpl (/root/reference/unet.hpp:13-70), so train.cpp / evaluate.cpp-style drivers port 1:1 with
//   torch::Tensor  ->  float* (fp32 NCDHW, x fastest)      model->forward(x)[0]  ->  model->forward(x, out, 1)
// Errors surface as std::runtime_error carrying unet3d_last_error(), like the reference's constructor.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "unet3d_b200.h"

struct UNet3d {
    int in_count = 1, out_count = 1;                       // unet.hpp:16-17
    std::string architecture, preproc, postproc = "softmax+create_mask+argmax", orientation, fov_strategy = "align_top", error_msg;
    std::vector<float> testing_errors, training_errors;    // 3 floats (ce, dice, mse) per step (train.cpp:746-752)
    float voxel_size[3] = {1.f, 1.f, 1.f};                 // unet.hpp:37
    int dim[3] = {192, 224, 192};                          // unet.hpp:38 (W, H, D)

    UNet3d(int32_t in_count_, int32_t out_count_, const std::string& feature_string, int gpu = 0)
        : in_count(in_count_), out_count(out_count_), architecture(feature_string) {
        check(unet3d_create(in_count_, out_count_, feature_string.c_str(), gpu, &h_));
    }
    ~UNet3d() { unet3d_destroy(h_); }
    UNet3d(const UNet3d&) = delete;
    UNet3d& operator=(const UNet3d&) = delete;

    static std::string default_feature(int out_count) {    // train.cpp:1054-1069
        std::string s(4096, '\0');
        if (unet3d_default_feature(out_count, &s[0], s.size())) throw std::runtime_error("default_feature");
        s.resize(std::char_traits<char>::length(s.c_str()));
        return s;
    }
    void set_dim(int w, int h, int d) { dim[0] = w; dim[1] = h; dim[2] = d; check(unet3d_set_dim(h_, w, h, d)); }

    // parameters() in tensorN order (main.cpp:193-204)
    int parameter_count() const { return unet3d_param_count(h_); }
    std::vector<int64_t> parameter_shape(int i) const {
        int64_t d[5]; int nd = 0;
        check(unet3d_param_shape(h_, i, d, &nd));
        return std::vector<int64_t>(d, d + nd);
    }
    void get_parameter(int i, float* host) { check(unet3d_get_param(h_, i, host)); }
    void set_parameter(int i, const float* host) { check(unet3d_set_param(h_, i, host)); }

    void train(bool on = true) { check(unet3d_set_mode(h_, on ? 1 : 2)); }           // unet.hpp:58-62 (train(false) == eval())
    void eval() { check(unet3d_set_mode(h_, 2)); }
    void prepare_for_inference() { check(unet3d_set_mode(h_, 0)); }                   // unet.cpp:7-22
    void create_optimizer(float learning_rate) { check(unet3d_create_optimizer(h_, learning_rate)); }  // unet.cpp:246-277
    void copy_from(const UNet3d& r) {                                                 // unet.cpp:195-222
        check(unet3d_copy_from(h_, r.h_));
        for (int k = 0; k < 3; ++k) { dim[k] = r.dim[k]; voxel_size[k] = r.voxel_size[k]; }
        fov_strategy = r.fov_strategy; postproc = r.postproc; preproc = r.preproc;
    }
    // forward (unet.cpp:168-193): out_levels[k] <- results[k]
    void forward(const float* in, float* const* out_levels, int n_levels, int where = 0) {
        check(unet3d_forward(h_, in, out_levels, n_levels, where));
    }
    // one micro-batch of train.cpp:628-706; returns level-0 {ce, dice, mse}
    void train_microbatch(const float* in, const float* label, float loss[3], int collapse_before = 0, bool cost_ce = true,
                          bool cost_dice = true, bool cost_mse = true, int where = 0) {
        check(unet3d_train_microbatch(h_, in, label, collapse_before, cost_ce, cost_dice, cost_mse, loss, nullptr, where));
    }
    // train.cpp:755-766 (nccl_comm = ncclComm_t for data parallel, or nullptr)
    void step(int batch_size, double lr, void* nccl_comm = nullptr) { check(unet3d_step(h_, batch_size, lr, nccl_comm)); }
    unet3d_t* handle() const { return h_; }

  private:
    unet3d_t* h_ = nullptr;
    static void check(int rc) {
        if (rc != 0) throw std::runtime_error(unet3d_last_error());
    }
};
