// TEST INFRASTRUCTURE ONLY (oracle).  Not product code; nothing under unet-studio_b200/ may call this.
//
// Driver around the UNMODIFIED reference model: /root/reference/unet.cpp + unet.hpp are compiled in
// place (see oracle/build_ref.sh) against the image's libtorch with the TIPL shim in shim/TIPL.
// calc_losses (train.cpp:501-552) and default_feature (train.cpp:1054-1069) are pure
// libtorch / std::string functions; build_ref.sh extracts their text from /root/reference/train.cpp
// at build time into the git-ignored oracle/_ref/train_extract.inc which is #included below, so the
// loss arithmetic is the reference's own code as well.  What is restated here (train.cpp cannot be
// compiled: Qt + TIPL NIfTI) is only the glue around them:
//   * deep-supervision micro-batch body            train.cpp:628-706
//   * grad /= batch, clip_grad_norm_(12), SGD step  train.cpp:759-766 (+ unet.cpp:246-277, real code)
//   * poly learning-rate                            train.cpp:566
//   * inference window forward                      evaluate.cpp:223-230
//
// Sub-commands
//   dump  : run forward (and optionally one optimizer step over --batch micro-batches), write raw
//           little-endian float32 .bin files + manifest.json
//   time  : time forward or a full step on CPU (or cuda when available), print one JSON line
#include <torch/torch.h>
#include <ATen/Parallel.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

#include "unet.hpp"
#include "train_extract.inc"  // generated: calc_losses + default_feature, verbatim from train.cpp

namespace {

struct Args {
    std::map<std::string, std::vector<std::string>> kv;
    std::string cmd;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string s(const std::string& k, const std::string& d = "") const {
        auto it = kv.find(k);
        return it == kv.end() || it->second.empty() ? d : it->second[0];
    }
    double f(const std::string& k, double d) const { return has(k) ? std::atof(s(k).c_str()) : d; }
    long i(const std::string& k, long d) const { return has(k) ? std::atol(s(k).c_str()) : d; }
};

Args parse(int argc, char** argv) {
    Args a;
    if (argc > 1) a.cmd = argv[1];
    std::string cur;
    for (int i = 2; i < argc; ++i) {
        std::string t = argv[i];
        if (t.rfind("--", 0) == 0) {
            cur = t.substr(2);
            a.kv[cur];
        } else
            a.kv[cur].push_back(t);
    }
    return a;
}

std::vector<float> read_f32(const std::string& path, size_t expect) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::vector<float> v(expect);
    in.read(reinterpret_cast<char*>(v.data()), expect * sizeof(float));
    if (size_t(in.gcount()) != expect * sizeof(float)) throw std::runtime_error("short read " + path);
    return v;
}

void write_tensor(const std::string& path, const torch::Tensor& t) {
    auto c = t.detach().to(torch::kCPU).to(torch::kFloat32).contiguous();
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(c.data_ptr<float>()), c.numel() * sizeof(float));
}

std::string shape_json(const torch::Tensor& t) {
    std::ostringstream o;
    o << "[";
    for (int64_t i = 0; i < t.dim(); ++i) o << (i ? "," : "") << t.size(i);
    o << "]";
    return o.str();
}

std::string read_text(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

// -1 = leave libtorch's defaults (cuDNN convolutions may use TF32), 0 = strict fp32, 1 = allow TF32 everywhere
void set_tf32(int mode) {
    if (mode < 0) return;
    at::globalContext().setAllowTF32CuDNN(mode != 0);
    at::globalContext().setAllowTF32CuBLAS(mode != 0);
}

struct StepFlags {
    bool ce = true, dice = true, mse = true;
    int collapse = 0;
};

// train.cpp:628-706 — one micro-batch: forward, 5-level deep supervision, backward.
// Returns the level-0 [ce,dice,mse] (what the reference logs, train.cpp:675-681) and, if asked,
// every level's losses and logits.
torch::Tensor micro_batch(UNet3d& model, torch::Tensor in, torch::Tensor target, const StepFlags& fl,
                          std::vector<torch::Tensor>* logits_out, std::vector<torch::Tensor>* level_losses,
                          bool do_backward) {
    auto outputs = model->forward(in);
    if (logits_out) *logits_out = outputs;
    torch::Tensor active_target = target;
    torch::Tensor total_loss, logged;
    const size_t out_sz = outputs.size();
    float weight_sum = 0.0f;
    for (size_t k = 0; k < out_sz; ++k) weight_sum += 1.0f / (1 << k);
    const float inv_weight_sum = 1.0f / weight_sum;
    for (size_t k = 0; k < out_sz; ++k) {
        if (k > 0) {
            int64_t d = active_target.size(1) >> 1, h = active_target.size(2) >> 1, w = active_target.size(3) >> 1;
            if (d <= 0 || h <= 0 || w <= 0) throw std::runtime_error("deep supervision target size became zero");
            auto tf = active_target.unsqueeze(1).to(torch::kFloat32);
            auto opt = torch::nn::functional::InterpolateFuncOptions()
                           .size(std::vector<int64_t>{d, h, w})
                           .mode(torch::kNearest);
            active_target = torch::nn::functional::interpolate(tf, opt).squeeze(1).to(torch::kLong);
        }
        if (!outputs[k].defined()) throw std::runtime_error("undefined deep supervision output");
        auto [ce, dice, mse] = calc_losses(outputs[k], active_target, model->out_count, fl.collapse);
        auto e = torch::stack({ce.detach(), dice.detach(), mse.detach()});
        if (k == 0) logged = e;
        if (level_losses) level_losses->push_back(e);
        const float norm_weight = (1.0f / (1 << k)) * inv_weight_sum;
        torch::Tensor level_loss;
        if (fl.ce) level_loss = level_loss.defined() ? level_loss + ce : ce;
        if (fl.dice) level_loss = level_loss.defined() ? level_loss + dice : dice;
        if (fl.mse) level_loss = level_loss.defined() ? level_loss + mse : mse;
        if (!level_loss.defined()) level_loss = ce;
        level_loss = level_loss * norm_weight;
        total_loss = total_loss.defined() ? total_loss + level_loss : level_loss;
    }
    if (do_backward) total_loss.backward();
    return logged;
}

// train.cpp:759-766
void update(UNet3d& model, int batch_size, double lr) {
    for (auto& group : model->optimizer->param_groups())
        static_cast<torch::optim::SGDOptions&>(group.options()).lr(lr);
    for (auto& p : model->parameters())
        if (p.grad().defined()) p.grad().div_(batch_size);
    torch::nn::utils::clip_grad_norm_(model->parameters(), 12.0);
    model->optimizer->step();
    model->optimizer->zero_grad();
}

UNet3d build(const Args& a, int& in_c, int& out_c, std::string& feature) {
    in_c = int(a.i("in_c", 1));
    out_c = int(a.i("out_c", 1));
    std::string f = a.s("feature", "default");
    feature = (f == "default") ? default_feature(out_c) : read_text(f);
    torch::manual_seed(uint64_t(a.i("seed", 0)));
    return UNet3d(in_c, out_c, feature);
}

int cmd_dump(const Args& a) {
    int in_c, out_c;
    std::string feature;
    UNet3d model = build(a, in_c, out_c, feature);
    auto dims = a.kv.at("dim");
    const int64_t W = std::atol(dims[0].c_str()), H = std::atol(dims[1].c_str()), D = std::atol(dims[2].c_str());
    model->dim = {unsigned(W), unsigned(H), unsigned(D)};
    const std::string outdir = a.s("outdir", ".");
    const int batch = int(a.i("batch", 1));
    const bool train = a.i("train", 0) != 0;
    const int steps = int(a.i("steps", 1));
    const int total_steps = int(a.i("total_steps", 1000));
    const double lr0 = a.f("lr", 1e-3);
    const bool eval_mode = a.i("eval", 0) != 0;
    StepFlags fl;
    fl.ce = a.i("ce", 1) != 0;
    fl.dice = a.i("dice", 1) != 0;
    fl.mse = a.i("mse", 1) != 0;
    fl.collapse = int(a.i("collapse", 0));
    // --device cuda: the SAME reference code on libtorch CUDA + cuDNN (what the north star calls "the reference's own libtorch
    // implementation"); dump defaults to strict fp32 (--tf32 0) so that it is an fp32 oracle at full-grid sizes in seconds
    torch::Device dev(a.s("device", "cpu") == "cuda" ? torch::kCUDA : torch::kCPU);
    if (dev.is_cuda() && !torch::cuda::is_available()) throw std::runtime_error("--device cuda: no CUDA device");
    set_tf32(int(a.i("tf32", 0)));
    if (a.has("threads")) at::set_num_threads(int(a.i("threads", 1)));
    const int logits_levels = int(a.i("logits_levels", 99));   // how many levels' logits to write (full-grid dumps are large)
    const bool write_after = a.i("write_after", 1) != 0;

    auto params = model->parameters();
    if (a.has("params_in")) {
        torch::NoGradGuard ng;
        for (size_t i = 0; i < params.size(); ++i) {
            char nm[64];
            std::snprintf(nm, sizeof nm, "/param_%03zu.bin", i);
            auto v = read_f32(a.s("params_in") + nm, params[i].numel());
            params[i].copy_(torch::from_blob(v.data(), params[i].sizes(), torch::kFloat32));
        }
    }
    std::ostringstream man;
    man << "{\n \"in_c\": " << in_c << ", \"out_c\": " << out_c << ", \"dim\": [" << W << "," << H << "," << D
        << "], \"batch\": " << batch << ", \"train\": " << (train ? 1 : 0) << ", \"steps\": " << steps
        << ", \"lr\": " << lr0 << ", \"total_steps\": " << total_steps << ",\n \"torch\": \"" << TORCH_VERSION
        << "\",\n \"params\": [";
    {
        auto named = model->named_parameters();
        size_t i = 0;
        for (auto& p : named) {
            char nm[64];
            std::snprintf(nm, sizeof nm, "/param_%03zu.bin", i);
            write_tensor(outdir + nm, p.value());
            man << (i ? "," : "") << "\n  {\"name\": \"" << p.key() << "\", \"shape\": " << shape_json(p.value()) << "}";
            ++i;
        }
    }
    man << "],\n";

    const size_t vox = size_t(W) * H * D;
    auto in_all = read_f32(a.s("input"), size_t(batch) * in_c * vox);
    std::vector<float> lab_all;
    if (a.has("label")) lab_all = read_f32(a.s("label"), size_t(batch) * vox);

    if (!train) {
        // evaluate.cpp:223-230 — forward only (NoGradGuard), all heads dumped; the caller uses [0].
        if (eval_mode) model->prepare_for_inference(dev);
        else model->to(dev);
        torch::NoGradGuard ng;
        auto in = torch::from_blob(in_all.data(), {1, in_c, D, H, W}, torch::kFloat32).clone().to(dev);
        if (a.i("dump_acts", 0)) {
            // restated forward (unet.cpp:168-193) with per-level taps, for layer-level debugging
            std::vector<torch::Tensor> skips(model->encoding.size() - 1);
            auto x = in;
            for (size_t l = 0; l < model->encoding.size(); ++l) {
                x = model->encoding[l]->forward(x);
                write_tensor(outdir + "/act_enc" + std::to_string(l) + ".bin", x);
                if (l + 1 < model->encoding.size()) skips[l] = x;
            }
            for (int l = int(model->encoding.size()) - 2; l >= 0; --l) {
                x = torch::cat({skips[l], x}, 1);
                x = model->decoding[l]->forward(x);
                write_tensor(outdir + "/act_dec" + std::to_string(l) + ".bin", x);
                if (!model->decoding_tail[l]->is_empty()) x = model->decoding_tail[l]->forward(x);
            }
        }
        auto outs = model->forward(in);
        man << " \"logits\": [";
        for (size_t k = 0; k < outs.size(); ++k) {
            if (int(k) < logits_levels) write_tensor(outdir + "/logits_" + std::to_string(k) + ".bin", outs[k]);
            man << (k ? "," : "") << shape_json(outs[k]);
        }
        man << "]\n}\n";
    } else {
        model->to(dev);
        params = model->parameters();
        model->train();
        model->create_optimizer(float(lr0));
        man << " \"losses\": [";
        for (int s = 0; s < steps; ++s) {
            const double lr = lr0 * std::pow(1.0 - double(s) / total_steps, 0.9);  // train.cpp:566
            torch::Tensor logged_sum;
            for (int b = 0; b < batch; ++b) {
                auto in = torch::from_blob(in_all.data() + size_t(b) * in_c * vox, {1, in_c, D, H, W}, torch::kFloat32).clone().to(dev);
                auto tg = torch::from_blob(lab_all.data() + size_t(b) * vox, {1, D, H, W}, torch::kFloat32).clone().to(torch::kLong).to(dev);
                std::vector<torch::Tensor> logits, lv;
                auto e = micro_batch(model, in, tg, fl, &logits, &lv, true);
                logged_sum = logged_sum.defined() ? logged_sum + e : e;
                if (s == 0 && b == 0) {
                    for (size_t k = 0; k < logits.size() && int(k) < logits_levels; ++k)
                        write_tensor(outdir + "/logits_" + std::to_string(k) + ".bin", logits[k]);
                    write_tensor(outdir + "/level_losses.bin", torch::stack(lv));
                }
                if (s == 0) {   // every micro-batch's level losses of the first step (gradient accumulation cases)
                    char nm[64];
                    std::snprintf(nm, sizeof nm, "/level_losses_mb%02d.bin", b);
                    write_tensor(outdir + nm, torch::stack(lv));
                }
            }
            auto logged = (logged_sum / double(batch)).contiguous();
            man << (s ? "," : "") << "[" << logged[0].item<float>() << "," << logged[1].item<float>() << ","
                << logged[2].item<float>() << "]";
            if (s == 0) {
                for (size_t i = 0; i < params.size(); ++i) {
                    char nm[64];
                    std::snprintf(nm, sizeof nm, "/grad_%03zu.bin", i);
                    write_tensor(outdir + nm, params[i].grad().defined() ? params[i].grad() : torch::zeros_like(params[i]));
                }
            }
            update(model, batch, lr);
        }
        man << "]";
        if (a.i("validate", 0)) {
            // train.cpp:834-840: output_model->eval(); calc_losses(forward(test_in)[0], test_out, out_count) under NoGradGuard
            torch::NoGradGuard ng;
            model->eval();
            auto in = torch::from_blob(in_all.data(), {1, in_c, D, H, W}, torch::kFloat32).clone().to(dev);
            auto tg = torch::from_blob(lab_all.data(), {1, D, H, W}, torch::kFloat32).clone().to(torch::kLong).to(dev);
            auto out0 = model->forward(in)[0];
            auto [ce, dice, mse] = calc_losses(out0, tg, model->out_count);
            man << ",\n \"validate\": [" << ce.item<float>() << "," << dice.item<float>() << "," << mse.item<float>() << "]";
            write_tensor(outdir + "/validate_logits_0.bin", out0);
            write_tensor(outdir + "/validate_losses.bin", torch::stack({ce, dice, mse}));
        }
        man << "\n}\n";
        for (size_t i = 0; write_after && i < params.size(); ++i) {
            char nm[64];
            std::snprintf(nm, sizeof nm, "/param_after_%03zu.bin", i);
            write_tensor(outdir + nm, params[i]);
        }
    }
    std::ofstream(outdir + "/manifest.json") << man.str();
    return 0;
}

int cmd_time(const Args& a) {
    int in_c, out_c;
    std::string feature;
    UNet3d model = build(a, in_c, out_c, feature);
    auto dims = a.kv.at("dim");
    const int64_t W = std::atol(dims[0].c_str()), H = std::atol(dims[1].c_str()), D = std::atol(dims[2].c_str());
    const int threads = int(a.i("threads", long(std::thread::hardware_concurrency())));
    at::set_num_threads(threads);
    const std::string mode = a.s("mode", "fwd");
    const int steps = int(a.i("steps", 3)), warmup = int(a.i("warmup", 1)), batch = int(a.i("batch", 1));
    torch::Device dev(a.s("device", "cpu") == "cuda" ? torch::kCUDA : torch::kCPU);
    if (dev.is_cuda() && !torch::cuda::is_available()) throw std::runtime_error("--device cuda: no CUDA device");
    const int tf32 = int(a.i("tf32", -1));
    set_tf32(tf32);
    StepFlags fl;
    torch::manual_seed(1);
    auto in = torch::rand({1, in_c, D, H, W});
    auto tg = torch::randint(0, std::max(out_c, 1), {1, D, H, W}, torch::kLong);
    std::vector<double> ms;
    double resident_ms = -1.0;
    // --budget_s: stop taking timed steps once this much wall time has been spent on them (the reported "steps" is what ran)
    const double budget_s = a.f("budget_s", 1e30);
    const auto t_begin = std::chrono::steady_clock::now();
    auto over_budget = [&]() {
        return !ms.empty() && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() > budget_s;
    };
    if (mode == "fwd") {
        model->prepare_for_inference(dev);
        torch::NoGradGuard ng;
        for (int s = 0; s < warmup + steps && !over_budget(); ++s) {
            auto t0 = std::chrono::steady_clock::now();
            // evaluate.cpp:226-229: H2D, forward()[0], D2H
            auto r = model->forward(in.to(dev))[0].to(torch::kCPU).contiguous();
            auto t1 = std::chrono::steady_clock::now();
            if (s >= warmup) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
        if (dev.is_cuda()) {   // the same forward with the window resident in device memory (no H2D / D2H)
            auto in_dev = in.to(dev);
            torch::cuda::synchronize();
            auto t0 = std::chrono::steady_clock::now();
            for (int s = 0; s < steps; ++s) auto r = model->forward(in_dev)[0];
            torch::cuda::synchronize();
            resident_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
        }
    } else {
        model->to(dev);
        model->train();
        model->create_optimizer(1e-3f);
        for (int s = 0; s < warmup + steps && !over_budget(); ++s) {
            auto t0 = std::chrono::steady_clock::now();
            for (int b = 0; b < batch; ++b) {
                // train.cpp:615-626: every micro-batch is uploaded from host memory; the logged losses come back (train.cpp:675-681)
                auto logged = micro_batch(model, in.to(dev), tg.to(dev), fl, nullptr, nullptr, true);
                logged.to(torch::kCPU);
            }
            update(model, batch, 1e-3);
            if (dev.is_cuda()) torch::cuda::synchronize();
            auto t1 = std::chrono::steady_clock::now();
            if (s >= warmup) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
    }
    double sum = 0;
    for (double m : ms) sum += m;
    std::sort(ms.begin(), ms.end());
    std::printf("{\"mode\": \"%s\", \"threads\": %d, \"dim\": [%ld,%ld,%ld], \"in_c\": %d, \"out_c\": %d, \"batch\": %d, "
                "\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.3f, \"ms_median\": %.3f, \"ms_resident\": %.3f, \"device\": \"%s\", "
                "\"tf32\": %d, \"cudnn_tf32\": %d, \"torch\": \"%s\"}\n",
                mode.c_str(), threads, long(W), long(H), long(D), in_c, out_c, batch, int(ms.size()), warmup, sum / ms.size(), ms[ms.size() / 2],
                resident_ms, dev.is_cuda() ? "cuda" : "cpu", tf32, int(at::globalContext().allowTF32CuDNN()), TORCH_VERSION);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    try {
        Args a = parse(argc, argv);
        if (a.cmd == "dump") return cmd_dump(a);
        if (a.cmd == "time") return cmd_time(a);
        if (a.cmd == "feature") {  // print default_feature(out_c) (train.cpp:1054-1069)
            std::cout << default_feature(int(a.i("out_c", 1)));
            return 0;
        }
        std::cerr << "usage: unet_ref dump|time|feature ...\n";
        return 2;
    } catch (const std::exception& e) {
        std::cerr << "unet_ref: " << e.what() << std::endl;
        return 1;
    }
}
