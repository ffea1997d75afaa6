// Compiles include/unet3d.hpp and the binding stub of INTEGRATION.md ("The binding stub") with a plain host compiler and runs it
// against libunet3d_b200.so.  tipl::image<3> and training_param are reduced to the members the stub touches.
// Without a GPU the constructor must fail loudly (no CPU fallback): prints NO_GPU and exits 0.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "unet3d.hpp"

namespace tipl {
template <int N>
struct image {
    std::vector<float> v;
    const float* data() const { return v.data(); }
    float* data() { return v.data(); }
};
}  // namespace tipl
struct training_param {   // train.hpp: the fields the step body reads
    float learning_rate = 1e-3f;
    int batch_size = 1, epoch = 100;
    bool cost_ce = true, cost_dice = true, cost_mse = true;
};
using ncclComm_t = void*;

// ---- verbatim from INTEGRATION.md ----
using UNet3dPtr = std::shared_ptr<UNet3d>;

inline void train_step_body(UNet3d& m, const tipl::image<3>& in, const tipl::image<3>& label,   // train.cpp:615-706
                            const training_param& param, int collapse_before, float logged[3]) {
    m.train_microbatch(in.data(), label.data(), logged, collapse_before, param.cost_ce, param.cost_dice, param.cost_mse);
}
inline void update(UNet3d& m, const training_param& param, size_t cur_epoch, ncclComm_t comm) {  // train.cpp:566,755-766
    const double lr = param.learning_rate * std::pow(1.0 - double(cur_epoch) / param.epoch, 0.9);
    m.step(param.batch_size, lr, comm);
}
// ---------------------------------------

int main() {
    const std::string feature = UNet3d::default_feature(2);
    std::printf("FEATURE_LINES %d\n", int(std::count(feature.begin(), feature.end(), '\n')) + 1);
    try {
        UNet3d bad(1, 2, "conv8,ks5,stride1\nconv8\nconv2,ks1");
        std::printf("UNEXPECTED\n");
        return 1;
    } catch (const std::runtime_error& e) {
        std::printf("CTOR_ERROR %s\n", e.what());   // the reference's message (unet.cpp:66) or "no CUDA device ..."
    }
    UNet3dPtr model;
    try {
        model = std::make_shared<UNet3d>(1, 2, feature, 0);
    } catch (const std::runtime_error& e) {
        std::printf("NO_GPU %s\n", e.what());
        return 0;
    }
    const int W = 32, H = 32, D = 32;
    model->set_dim(W, H, D);
    if (unet3d_init_params(model->handle(), 1)) return 2;
    model->train();
    model->create_optimizer(1e-3f);
    tipl::image<3> in, label;
    in.v.resize(size_t(W) * H * D);
    label.v.resize(in.v.size());
    for (size_t i = 0; i < in.v.size(); ++i) {
        const int x = int(i % W), y = int(i / W % H), z = int(i / (size_t(W) * H));
        const float r = std::sqrt(float((x - 16) * (x - 16) + (y - 16) * (y - 16) + (z - 16) * (z - 16))) / 12.f;
        in.v[i] = r < 1.f ? 1.f - 0.5f * r : 0.f;
        label.v[i] = r < 1.f ? 1.f : 0.f;
    }
    training_param param;
    float logged[3] = {0, 0, 0};
    for (size_t epoch = 0; epoch < 3; ++epoch) {
        train_step_body(*model, in, label, param, 0, logged);
        update(*model, param, epoch, nullptr);
    }
    std::printf("LOSSES %.6f %.6f %.6f\n", logged[0], logged[1], logged[2]);
    model->prepare_for_inference();
    std::vector<float> out(size_t(2) * W * H * D);
    float* outs[1] = {out.data()};
    model->forward(in.data(), outs, 1);
    std::printf("LOGIT0 %.6f\n", out[0]);
    return std::isfinite(logged[0]) && std::isfinite(out[0]) ? 0 : 3;
}
