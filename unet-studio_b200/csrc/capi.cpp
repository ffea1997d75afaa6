// C-ABI glue (include/unet3d_b200.h): error string + model-level entry points.  Nothing throws across
// the boundary: the reference converts worker exceptions into error_msg + aborted (train.cpp:709-721).
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/unet3d_b200.h"
#include "model.h"
#include "vpa.h"
#include "simulate.h"
#include "modelfile.h"
#include "postproc.h"
#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>
#include <map>

namespace u3d {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }
int nccl_unique_id(void* out128);
int nccl_comm_init(void** comm, int nranks, int rank, const void* id128);
int nccl_comm_destroy(void* comm);
}  // namespace u3d

using u3d::Model;
using u3d::set_error;

struct unet3d {
    Model* m;
};

#define GUARD_BEGIN try {
#define GUARD_END                                   \
    }                                               \
    catch (const std::exception& e) {               \
        set_error(e.what());                        \
        return 1;                                   \
    }                                               \
    catch (...) {                                   \
        set_error("unknown error");                 \
        return 1;                                   \
    }
#define NEED(h)                                     \
    if (!(h) || !(h)->m) {                          \
        set_error("null handle");                   \
        return 1;                                   \
    }

extern "C" {

const char* unet3d_last_error(void) { return u3d::last_error(); }

int unet3d_default_feature(int out_count, char* buf, size_t buflen) {
    const std::string f = u3d::default_feature(out_count);
    if (!buf || buflen < f.size() + 1) { set_error("buffer too small"); return int(f.size() + 1); }
    std::memcpy(buf, f.c_str(), f.size() + 1);
    return 0;
}

int unet3d_describe(int in_count, int out_count, const char* feature_string, char* json, size_t json_len) {
    GUARD_BEGIN
    if (!feature_string) { set_error("null argument"); return 1; }
    Model m(in_count, out_count, feature_string, true);
    std::string o = "{\"levels\": " + std::to_string(m.n_levels()) + ", \"params\": [";
    for (size_t i = 0; i < m.params.size(); ++i) {
        const auto& p = m.params[i];
        o += std::string(i ? "," : "") + "{\"name\": \"" + p.name + "\", \"shape\": [";
        for (size_t k = 0; k < p.shape.size(); ++k) o += std::string(k ? "," : "") + std::to_string(p.shape[k]);
        o += "], \"decay\": " + std::string(p.decay ? "1" : "0") + "}";
    }
    o += "]}";
    if (!json || json_len < o.size() + 1) { set_error("buffer too small"); return int(o.size() + 1); }
    std::memcpy(json, o.c_str(), o.size() + 1);
    return 0;
    GUARD_END
}

int unet3d_create(int in_count, int out_count, const char* feature_string, int gpu, unet3d_t** out) {
    GUARD_BEGIN
    if (!out || !feature_string) { set_error("null argument"); return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libunet3d_b200 is a B200 (sm_100a) library and has no CPU fallback");
        return 1;
    }
    if (gpu < 0 || gpu >= ndev) { set_error("invalid gpu index"); return 1; }
    if (cudaSetDevice(gpu) != cudaSuccess) { set_error("cudaSetDevice failed"); return 1; }
    unet3d* h = new unet3d{nullptr};
    try {
        h->m = new Model(in_count, out_count, feature_string);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    return 0;
    GUARD_END
}

void unet3d_destroy(unet3d_t* h) {
    if (!h) return;
    delete h->m;
    delete h;
}

int unet3d_in_count(const unet3d_t* h) { return h && h->m ? h->m->in_count : -1; }
int unet3d_out_count(const unet3d_t* h) { return h && h->m ? h->m->out_count : -1; }
int unet3d_levels(const unet3d_t* h) { return h && h->m ? h->m->n_levels() : -1; }
const char* unet3d_architecture(const unet3d_t* h) { return h && h->m ? h->m->architecture.c_str() : ""; }

int unet3d_set_dim(unet3d_t* h, int w, int hh, int d) {
    GUARD_BEGIN NEED(h) return h->m->set_dim(w, hh, d);
    GUARD_END
}
int unet3d_get_dim(const unet3d_t* h, int dim3[3]) {
    if (!h || !h->m) return 1;
    for (int k = 0; k < 3; ++k) dim3[k] = h->m->dim[k];
    return 0;
}
int unet3d_set_voxel_size(unet3d_t* h, float x, float y, float z) {
    if (!h || !h->m) return 1;
    h->m->voxel_size[0] = x; h->m->voxel_size[1] = y; h->m->voxel_size[2] = z;
    return 0;
}

int unet3d_param_count(const unet3d_t* h) { return h && h->m ? int(h->m->params.size()) : -1; }
long long unet3d_param_total(const unet3d_t* h) {
    if (!h || !h->m) return -1;
    long long n = 0;
    for (auto& p : h->m->params) n += p.numel;
    return n;
}
int unet3d_param_shape(const unet3d_t* h, int i, int64_t dims[5], int* ndim) {
    if (!h || !h->m || i < 0 || i >= int(h->m->params.size())) { set_error("parameter index out of range"); return 1; }
    const auto& s = h->m->params[i].shape;
    *ndim = int(s.size());
    for (size_t k = 0; k < s.size(); ++k) dims[k] = s[k];
    return 0;
}
const char* unet3d_param_name(const unet3d_t* h, int i) {
    if (!h || !h->m || i < 0 || i >= int(h->m->params.size())) return "";
    return h->m->params[i].name.c_str();
}
int unet3d_param_decay(const unet3d_t* h, int i) {
    if (!h || !h->m || i < 0 || i >= int(h->m->params.size())) return -1;
    return h->m->params[i].decay ? 1 : 0;
}
int unet3d_get_param(unet3d_t* h, int i, float* host) {
    GUARD_BEGIN NEED(h) return h->m->get_flat(h->m->d_params, i, host, 1.f);
    GUARD_END
}
int unet3d_set_param(unet3d_t* h, int i, const float* host) {
    GUARD_BEGIN NEED(h) return h->m->set_param(i, host);
    GUARD_END
}
int unet3d_get_grad(unet3d_t* h, int i, float* host) {
    GUARD_BEGIN NEED(h)
    const float s = h->m->loss_scale > 0.f ? 1.0f / h->m->loss_scale : 1.f;
    return h->m->get_flat(h->m->d_grads, i, host, s);
    GUARD_END
}
int unet3d_get_momentum(unet3d_t* h, int i, float* host) {
    GUARD_BEGIN NEED(h) return h->m->get_flat(h->m->d_mom, i, host, 1.f);
    GUARD_END
}
int unet3d_set_momentum(unet3d_t* h, int i, const float* host) {
    GUARD_BEGIN NEED(h) return h->m->set_momentum(i, host);
    GUARD_END
}
int unet3d_init_params(unet3d_t* h, uint64_t seed) {
    GUARD_BEGIN NEED(h) return h->m->init_params(seed);
    GUARD_END
}

int unet3d_set_mode(unet3d_t* h, int mode) {
    GUARD_BEGIN NEED(h) return h->m->set_mode(mode);
    GUARD_END
}

int unet3d_forward(unet3d_t* h, const float* in, float* const* out_levels, int n_levels, int where) {
    GUARD_BEGIN NEED(h) return h->m->forward(in, out_levels, n_levels, where);
    GUARD_END
}

int unet3d_evaluate_windows(unet3d_t* h, const float* const* in_windows, float* const* out_windows, int n_windows, int where) {
    GUARD_BEGIN NEED(h)
    return h->m->evaluate_windows(in_windows, out_windows, n_windows, where);
    GUARD_END
}

int unet3d_train_microbatch(unet3d_t* h, const float* in, const float* label, int collapse_before, int use_ce, int use_dice,
                            int use_mse, float loss_out[3], float* all_level_losses, int where) {
    GUARD_BEGIN NEED(h)
    return h->m->train_microbatch(in, label, collapse_before, use_ce, use_dice, use_mse, loss_out, all_level_losses, where);
    GUARD_END
}

int unet3d_validate(unet3d_t* h, const float* in, const float* label, int collapse_before, float loss_out[3], int where) {
    GUARD_BEGIN NEED(h) return h->m->validate(in, label, collapse_before, loss_out, where);
    GUARD_END
}

int unet3d_validate_async(unet3d_t* h, const float* in, const float* label, int collapse_before, int where) {
    GUARD_BEGIN NEED(h) return h->m->validate(in, label, collapse_before, nullptr, where);
    GUARD_END
}
int unet3d_validate_result(unet3d_t* h, float loss_out[3]) {
    GUARD_BEGIN NEED(h) return h->m->validate_result(loss_out);
    GUARD_END
}

int unet3d_create_optimizer(unet3d_t* h, float learning_rate) {
    GUARD_BEGIN NEED(h)
    h->m->optimizer_created = true;
    h->m->lr0 = learning_rate;
    return 0;
    GUARD_END
}

int unet3d_step(unet3d_t* h, int batch_size, double lr, void* nccl_comm) {
    GUARD_BEGIN NEED(h) return h->m->step(batch_size, lr, nccl_comm);
    GUARD_END
}

double unet3d_last_grad_norm(const unet3d_t* h) {
    if (!h || !h->m || h->m->resolve_status()) return -1.0;
    return h->m->last_grad_norm;
}
int unet3d_last_step_skipped(const unet3d_t* h) {
    if (!h || !h->m || h->m->resolve_status()) return -1;
    return h->m->last_step_skipped;
}
float unet3d_loss_scale(const unet3d_t* h) {
    if (!h || !h->m || h->m->resolve_status()) return -1.f;
    return h->m->loss_scale;
}
int unet3d_set_loss_scale(unet3d_t* h, float s) {
    if (!h || !h->m) return 1;
    h->m->loss_scale = s;
    h->m->loss_scale_max = s;
    return 0;
}
long long unet3d_launch_count(const unet3d_t* h) { return h && h->m ? h->m->launches : -1; }

int unet3d_copy_from(unet3d_t* dst, const unet3d_t* src) {
    GUARD_BEGIN NEED(dst) NEED(src) return dst->m->copy_from(*src->m);
    GUARD_END
}

int unet3d_timer_start(unet3d_t* h) {
    GUARD_BEGIN NEED(h) return h->m->timer_start();
    GUARD_END
}
int unet3d_timer_stop(unet3d_t* h, float* ms) {
    GUARD_BEGIN NEED(h) return h->m->timer_stop(ms);
    GUARD_END
}

int unet3d_profile(unet3d_t* h, int enable) {
    GUARD_BEGIN NEED(h)
    h->m->prof_on = enable != 0;
    return 0;
    GUARD_END
}
int unet3d_profile_read(unet3d_t* h, double out24[24], int reset) {
    GUARD_BEGIN NEED(h) return h->m->prof_read(out24, reset);
    GUARD_END
}

int unet3d_evaluate_volume(unet3d_t* h, const float* volume, int w, int hgt, int d, int stride_x, int stride_y, int stride_z,
                           float mask_threshold, uint8_t* label_out, float* fg_prob_out, float* label_prob_out, int where, int* n_windows) {
    GUARD_BEGIN NEED(h)
    return h->m->evaluate_volume(volume, w, hgt, d, stride_x, stride_y, stride_z, mask_threshold, label_out, fg_prob_out, label_prob_out,
                                 where, n_windows);
    GUARD_END
}

int unet3d_window_origins(int volume_dim, int window_dim, int stride, int* origins, int max_origins) {
    const std::vector<int> o = u3d::window_origins(volume_dim, window_dim, stride);
    for (size_t i = 0; i < o.size() && int(i) < max_origins; ++i) origins[i] = o[i];
    return int(o.size());
}

int u3d_postproc(const float* logits, int channels, long long voxels, float mask_threshold, uint8_t* label_out, float* fg_prob_out,
                 float* label_prob_out, int gpu) {
    GUARD_BEGIN
    if (!logits || !label_out || channels < 1 || voxels < 1) { set_error("u3d_postproc: invalid argument"); return 1; }
    if (cudaSetDevice(gpu) != cudaSuccess) { set_error("no CUDA device: libunet3d_b200 has no CPU fallback"); return 1; }
    float* buf = nullptr;
    const size_t V = size_t(voxels);
    if (cudaMalloc(reinterpret_cast<void**>(&buf), (size_t(2 * channels + 2) * V) * 4 + V) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
    float* d_log = buf;
    float* d_acc = d_log + size_t(channels) * V;
    float* d_cnt = d_acc + size_t(channels) * V;
    float* d_fg = d_cnt + V;
    uint8_t* d_lab = reinterpret_cast<uint8_t*>(d_fg + V);
    int rc = 0;
    cudaMemcpy(d_log, logits, size_t(channels) * V * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_acc, 0, (size_t(channels) + 1) * V * 4);
    // one "window" covering the whole buffer, viewed as a 1-D volume
    if (voxels >= (1LL << 31)) { set_error("u3d_postproc: too many voxels"); rc = 1; }
    else rc = u3d::softmax_accumulate_launch(d_log, d_acc, d_cnt, channels, int(voxels), 1, 1, int(voxels), 1, 1, 0, 0, 0, nullptr);
    if (!rc) rc = u3d::mask_argmax_launch(d_acc, d_cnt, d_lab, d_fg, channels, voxels, mask_threshold, label_prob_out != nullptr, nullptr);
    if (!rc) {
        cudaMemcpy(label_out, d_lab, V, cudaMemcpyDeviceToHost);
        if (fg_prob_out) cudaMemcpy(fg_prob_out, d_fg, V * 4, cudaMemcpyDeviceToHost);
        if (label_prob_out) cudaMemcpy(label_prob_out, d_acc, size_t(channels) * V * 4, cudaMemcpyDeviceToHost);
        if (cudaDeviceSynchronize() != cudaSuccess) { set_error("u3d_postproc: device error"); rc = 1; }
    }
    cudaFree(buf);
    return rc;
    GUARD_END
}

int u3d_resample(const float* src, int channels, int sw, int sh, int sd, float* dst, int dw, int dh, int dd, int nearest, int gpu) {
    GUARD_BEGIN
    if (!src || !dst || channels < 1 || sw < 1 || sh < 1 || sd < 1 || dw < 1 || dh < 1 || dd < 1) { set_error("u3d_resample: invalid argument"); return 1; }
    if (cudaSetDevice(gpu) != cudaSuccess) { set_error("no CUDA device: libunet3d_b200 has no CPU fallback"); return 1; }
    const size_t SV = size_t(sw) * sh * sd * channels, DV = size_t(dw) * dh * dd * channels;
    float* buf = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&buf), (SV + DV) * 4) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
    cudaMemcpy(buf, src, SV * 4, cudaMemcpyHostToDevice);
    int rc = u3d::resample_launch(buf, buf + SV, channels, sw, sh, sd, dw, dh, dd, nearest, nullptr);
    if (!rc) {
        cudaMemcpy(dst, buf + SV, DV * 4, cudaMemcpyDeviceToHost);
        if (cudaDeviceSynchronize() != cudaSuccess) { set_error("u3d_resample: device error"); rc = 1; }
    }
    cudaFree(buf);
    return rc;
    GUARD_END
}

// ---- model file (.nz), main.cpp:157-233 -------------------------------------------------------------------

struct u3d_nz {
    u3d::NzFile* f;
};

int u3d_nz_create(u3d_nz_t** out) {
    GUARD_BEGIN
    if (!out) { set_error("null argument"); return 1; }
    *out = new u3d_nz{u3d::nz_new()};
    return 0;
    GUARD_END
}
int u3d_nz_load(const char* path, u3d_nz_t** out) {
    GUARD_BEGIN
    if (!out || !path) { set_error("null argument"); return 1; }
    u3d_nz* z = new u3d_nz{u3d::nz_new()};
    if (u3d::nz_load(path, *z->f)) { u3d::nz_delete(z->f); delete z; return 1; }
    *out = z;
    return 0;
    GUARD_END
}
void u3d_nz_free(u3d_nz_t* z) {
    if (!z) return;
    u3d::nz_delete(z->f);
    delete z;
}
int u3d_nz_save(const u3d_nz_t* z, const char* path) {
    GUARD_BEGIN
    if (!z || !path) { set_error("null argument"); return 1; }
    return u3d::nz_save(*z->f, path);
    GUARD_END
}
int u3d_nz_count(const u3d_nz_t* z) { return z ? u3d::nz_count(*z->f) : -1; }
int u3d_nz_info(const u3d_nz_t* z, int i, char* name, size_t name_len, int* type, int* rows, int* cols) {
    GUARD_BEGIN
    if (!z) { set_error("null argument"); return 1; }
    std::string n;
    int t = 0, r = 0, c = 0;
    if (u3d::nz_info(*z->f, i, n, t, r, c)) return 1;
    if (name && name_len) { std::strncpy(name, n.c_str(), name_len - 1); name[name_len - 1] = 0; }
    if (type) *type = t;
    if (rows) *rows = r;
    if (cols) *cols = c;
    return 0;
    GUARD_END
}
int u3d_nz_add(u3d_nz_t* z, const char* name, int type, int rows, int cols, const void* data) {
    GUARD_BEGIN
    if (!z || !name || (!data && rows * cols > 0)) { set_error("null argument"); return 1; }
    return u3d::nz_add(*z->f, name, type, rows, cols, data);
    GUARD_END
}
int u3d_nz_read_f32(const u3d_nz_t* z, const char* name, float* out, size_t n) {
    GUARD_BEGIN
    if (!z || !name) { set_error("null argument"); return 1; }
    std::vector<float> v;
    if (u3d::nz_read_f32(*z->f, name, v)) return 1;
    if (v.size() != n) { set_error("matrix " + std::string(name) + " has " + std::to_string(v.size()) + " elements"); return 1; }
    if (n) std::memcpy(out, v.data(), n * 4);
    return 0;
    GUARD_END
}

int unet3d_load_from_file(const char* file_name, int gpu, unet3d_t** out) {
    GUARD_BEGIN
    if (!file_name || !out) { set_error("null argument"); return 1; }
    std::unique_ptr<u3d::NzFile, void (*)(u3d::NzFile*)> f(u3d::nz_new(), u3d::nz_delete);
    if (u3d::nz_load(file_name, *f)) return 1;
    int in_c = 1, out_c = 1;
    std::string arch;
    if (u3d::nz_model_header(*f, in_c, out_c, arch)) return 1;
    unet3d_t* h = nullptr;
    if (unet3d_create(in_c, out_c, arch.c_str(), gpu, &h)) return 1;
    if (u3d::nz_to_model(*f, *h->m)) { unet3d_destroy(h); return 1; }
    *out = h;
    return 0;
    GUARD_END
}
int unet3d_save_to_file(unet3d_t* h, const char* file_name) {
    GUARD_BEGIN NEED(h)
    if (!file_name) { set_error("null argument"); return 1; }
    std::unique_ptr<u3d::NzFile, void (*)(u3d::NzFile*)> f(u3d::nz_new(), u3d::nz_delete);
    if (u3d::model_to_nz(*h->m, *f)) return 1;
    return u3d::nz_save(*f, file_name);
    GUARD_END
}
int unet3d_save_optimizer(unet3d_t* h, const char* file_name) {
    GUARD_BEGIN NEED(h)
    if (!file_name) { set_error("null argument"); return 1; }
    std::unique_ptr<u3d::NzFile, void (*)(u3d::NzFile*)> f(u3d::nz_new(), u3d::nz_delete);
    if (u3d::model_momentum_to_nz(*h->m, *f)) return 1;
    return u3d::nz_save(*f, file_name);
    GUARD_END
}
int unet3d_load_optimizer(unet3d_t* h, const char* file_name) {
    GUARD_BEGIN NEED(h)
    if (!file_name) { set_error("null argument"); return 1; }
    std::unique_ptr<u3d::NzFile, void (*)(u3d::NzFile*)> f(u3d::nz_new(), u3d::nz_delete);
    if (u3d::nz_load(file_name, *f)) return 1;
    return u3d::nz_to_model_momentum(*f, *h->m);
    GUARD_END
}
int unet3d_export_raw(unet3d_t* h, const char* directory) {
    GUARD_BEGIN NEED(h)
    if (!directory) { set_error("null argument"); return 1; }
    return u3d::model_export_raw(*h->m, directory);
    GUARD_END
}
static std::string* info_field(Model* m, const char* key) {
    const std::string k = key ? key : "";
    if (k == "preproc") return &m->preproc;
    if (k == "postproc") return &m->postproc;
    if (k == "orientation") return &m->orientation;
    if (k == "fov_strategy") return &m->fov_strategy;
    return nullptr;
}
int unet3d_set_info(unet3d_t* h, const char* key, const char* value) {
    GUARD_BEGIN NEED(h)
    std::string* f = info_field(h->m, key);
    if (!f || !value) { set_error("unet3d_set_info: key must be preproc, postproc, orientation or fov_strategy"); return 1; }
    *f = value;
    return 0;
    GUARD_END
}
int unet3d_get_info(unet3d_t* h, const char* key, char* buf, size_t buflen) {
    GUARD_BEGIN NEED(h)
    const std::string* f = info_field(h->m, key);
    if (!f) { set_error("unet3d_get_info: key must be preproc, postproc, orientation or fov_strategy"); return 1; }
    if (!buf || buflen < f->size() + 1) { set_error("buffer too small"); return int(f->size() + 1); }
    std::memcpy(buf, f->c_str(), f->size() + 1);
    return 0;
    GUARD_END
}
int unet3d_set_errors(unet3d_t* h, int testing, const float* ce_dice_mse, int n_steps) {
    GUARD_BEGIN NEED(h)
    if (n_steps < 0 || (n_steps && !ce_dice_mse)) { set_error("null argument"); return 1; }
    auto& v = testing ? h->m->testing_errors : h->m->training_errors;
    v.assign(ce_dice_mse, ce_dice_mse + size_t(3) * n_steps);
    return 0;
    GUARD_END
}
int unet3d_get_errors(unet3d_t* h, int testing, float* ce_dice_mse, int max_steps) {
    if (!h || !h->m) return -1;
    const auto& v = testing ? h->m->testing_errors : h->m->training_errors;
    const int n = int(v.size() / 3);
    if (ce_dice_mse)
        for (int i = 0; i < std::min(n, max_steps) * 3; ++i) ce_dice_mse[i] = v[size_t(i)];
    return n;
}

int unet3d_sync(unet3d_t* h) {
    GUARD_BEGIN NEED(h) return h->m->sync();
    GUARD_END
}

// ---- visual_perception_augmentation (train.hpp:43-48) ------------------------------------------------------
namespace {
struct VpaWs { void* p = nullptr; size_t bytes = 0; cudaStream_t s = nullptr; };
std::mutex g_vpa_mu;
std::map<int, VpaWs> g_vpa_ws;
}  // namespace

// workspace shared by the two sample-preparation stages (they run one after the other on one stream): scratch first, then a
// staging area for host-buffer calls
static size_t sample_ws_need(int w, int h, int d, int channels) {
    const size_t stage = (size_t(channels) + 1) * size_t(w) * h * d * 4;
    return std::max(u3d::vpa_workspace_bytes(w, h, d, channels), u3d::simulate_workspace_bytes(w, h, d)) + stage;
}

// Two workspaces per handle: one for the calls that run on the main stream, one for the prefetch stream (stream3), so a prefetch in
// flight never shares scratch (displacement field, min/max cells, tissue planes) with a main-stream sample call.
static int ensure_model_ws(Model* m, size_t need, bool prefetch = false) {
    void*& ws = prefetch ? m->pf_ws : m->vpa_ws;
    size_t& bytes = prefetch ? m->pf_ws_bytes : m->vpa_ws_bytes;
    if (bytes >= need) return 0;
    cudaStreamSynchronize(m->stream);
    if (m->stream3) cudaStreamSynchronize(m->stream3);
    if (ws) cudaFree(ws);
    ws = nullptr;
    bytes = 0;
    if (cudaMalloc(&ws, need) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
    bytes = need;
    return 0;
}

// simulate_modality (train.cpp:43-180): label == nullptr selects the image-only overload
static int sim_impl(float* t1w, const float* label, unsigned max_label, unsigned seed, int w, int h, int d, int where, void* ws,
                    cudaStream_t stream, long long* launches) {
    u3d::SimPlan plan;
    if (u3d::simulate_make_plan(label != nullptr, max_label, seed, w, h, d, plan)) return 1;
    const size_t V = size_t(w) * h * d;
    float* dimg = t1w;
    const float* dlab = label;
    if (where == 0) {
        float* stage = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + u3d::simulate_workspace_bytes(w, h, d));
        dimg = stage;
        cudaMemcpyAsync(dimg, t1w, V * 4, cudaMemcpyHostToDevice, stream);
        if (label) {
            cudaMemcpyAsync(stage + V, label, V * 4, cudaMemcpyHostToDevice, stream);
            dlab = stage + V;
        }
    }
    int rc = u3d::simulate_run(plan, dimg, dlab, ws, stream, launches);
    if (where == 0) {
        if (!rc) cudaMemcpyAsync(t1w, dimg, V * 4, cudaMemcpyDeviceToHost, stream);
        cudaError_t e = cudaStreamSynchronize(stream);
        if (!rc && e != cudaSuccess) { set_error(std::string("simulate_modality: ") + cudaGetErrorString(e)); rc = 1; }
    }
    return rc;
}

// the fused / prefetched sample calls run simulate_modality first when the handle asks for it (train.cpp:459-462)
static int sim_stage(Model* m, float* din, const float* dlab, uint64_t seed, cudaStream_t stream, void* ws) {
    if (!m->sim_mode) return 0;
    if (m->in_count != 1) { set_error("simulate_modality needs a single-channel image"); return 1; }
    return sim_impl(din, m->sim_mode == 1 ? dlab : nullptr, unsigned(m->out_count), unsigned(seed), m->dim[0], m->dim[1], m->dim[2], 1,
                    ws, stream, &m->launches);
}

static int vpa_impl(const char* const* keys, const float* vals, int n_opts, float* image, float* label, int is_label, int w, int h,
                    int d, int channels, uint64_t seed, int where, void* ws, cudaStream_t stream, long long* launches) {
    u3d::VpaPlan plan;
    if (u3d::vpa_make_plan(keys, vals, n_opts, is_label, w, h, d, channels, seed, plan)) return 1;
    const size_t V = size_t(w) * h * d;
    float* dimg = image;
    float* dlab = label;
    if (where == 0) {
        // host buffers: stage through the tail of the persistent workspace (no per-call cudaMalloc / cudaFree)
        float* stage = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + u3d::vpa_workspace_bytes(w, h, d, channels));
        dimg = stage;
        dlab = stage + size_t(channels) * V;
        cudaMemcpyAsync(dimg, image, size_t(channels) * V * 4, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(dlab, label, V * 4, cudaMemcpyHostToDevice, stream);
    }
    int rc = u3d::vpa_run(plan, dimg, dlab, ws, stream, launches);
    if (where == 0) {
        if (!rc) {
            cudaMemcpyAsync(image, dimg, size_t(channels) * V * 4, cudaMemcpyDeviceToHost, stream);
            cudaMemcpyAsync(label, dlab, V * 4, cudaMemcpyDeviceToHost, stream);
        }
        cudaError_t e = cudaStreamSynchronize(stream);
        if (!rc && e != cudaSuccess) { set_error(std::string("vpa: ") + cudaGetErrorString(e)); rc = 1; }
    }
    return rc;
}

int vpa_augment(const char* const* keys, const float* vals, int n_opts, float* image, float* label, int is_label, int w, int h, int d,
                int channels, uint64_t seed, int where, int gpu) {
    GUARD_BEGIN
    if (cudaSetDevice(gpu) != cudaSuccess) { set_error("no CUDA device: vpa_augment has no CPU fallback"); return 1; }
    std::lock_guard<std::mutex> lock(g_vpa_mu);
    VpaWs& W = g_vpa_ws[gpu];
    const size_t need = sample_ws_need(w, h, d, channels);
    if (!W.s && cudaStreamCreateWithFlags(&W.s, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return 1; }
    if (W.bytes < need) {
        if (W.p) cudaFree(W.p);
        W.bytes = 0;
        if (cudaMalloc(&W.p, need) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
        W.bytes = need;
    }
    int rc = vpa_impl(keys, vals, n_opts, image, label, is_label, w, h, d, channels, seed, where, W.p, W.s, nullptr);
    if (!rc && where != 0 && cudaStreamSynchronize(W.s) != cudaSuccess) { set_error("vpa: kernel failed"); rc = 1; }
    return rc;
    GUARD_END
}

int unet3d_vpa_augment(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, float* image, float* label, int is_label,
                       int w, int hgt, int d, int channels, uint64_t seed, int where) {
    GUARD_BEGIN NEED(h)
    Model* m = h->m;
    cudaSetDevice(m->device);
    if (ensure_model_ws(m, sample_ws_need(w, hgt, d, channels))) return 1;
    return vpa_impl(keys, vals, n_opts, image, label, is_label, w, hgt, d, channels, seed, where, m->vpa_ws, m->stream, &m->launches);
    GUARD_END
}

int unet3d_train_microbatch_augmented(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, const float* image_host,
                                      const float* label_host, uint64_t seed, int collapse_before, int use_ce, int use_dice, int use_mse,
                                      float loss_out3[3]) {
    GUARD_BEGIN NEED(h)
    Model* m = h->m;
    cudaSetDevice(m->device);
    float* din = nullptr;
    float* dlab = nullptr;
    if (m->staging(&din, &dlab)) return 1;
    const int w = m->dim[0], hgt = m->dim[1], d = m->dim[2], channels = m->in_count;
    const size_t V = size_t(w) * hgt * d;
    if (ensure_model_ws(m, sample_ws_need(w, hgt, d, channels))) return 1;
    // one upload of the raw sample; augmentation and the micro-batch run stream-ordered on it without leaving HBM
    if (cudaMemcpyAsync(din, image_host, size_t(channels) * V * 4, cudaMemcpyHostToDevice, m->stream) != cudaSuccess ||
        cudaMemcpyAsync(dlab, label_host, V * 4, cudaMemcpyHostToDevice, m->stream) != cudaSuccess) {
        set_error("unet3d_train_microbatch_augmented: upload failed");
        return 1;
    }
    if (sim_stage(m, din, dlab, seed, m->stream, m->vpa_ws)) return 1;
    if (vpa_impl(keys, vals, n_opts, din, dlab, 1, w, hgt, d, channels, seed, 1, m->vpa_ws, m->stream, &m->launches)) return 1;
    return m->train_microbatch(din, dlab, collapse_before, use_ce, use_dice, use_mse, loss_out3, nullptr, 1);
    GUARD_END
}

int unet3d_prefetch_augmented(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, const float* image, const float* label,
                              uint64_t seed, int where) {
    GUARD_BEGIN NEED(h)
    Model* m = h->m;
    cudaSetDevice(m->device);
    float* din = nullptr;
    float* dlab = nullptr;
    int slot = 0;
    if (m->prefetch_slot(&din, &dlab, &slot)) return 1;
    const int w = m->dim[0], hgt = m->dim[1], d = m->dim[2], channels = m->in_count;
    const size_t V = size_t(w) * hgt * d;
    if (ensure_model_ws(m, sample_ws_need(w, hgt, d, channels), true)) return 1;
    const cudaMemcpyKind kind = where == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (cudaMemcpyAsync(din, image, size_t(channels) * V * 4, kind, m->stream3) != cudaSuccess ||
        cudaMemcpyAsync(dlab, label, V * 4, kind, m->stream3) != cudaSuccess) {
        set_error("unet3d_prefetch_augmented: copy failed");
        return 1;
    }
    if (sim_stage(m, din, dlab, seed, m->stream3, m->pf_ws)) return 1;
    if (vpa_impl(keys, vals, n_opts, din, dlab, 1, w, hgt, d, channels, seed, 1, m->pf_ws, m->stream3, &m->launches)) return 1;
    if (cudaEventRecord(m->ev_sample[slot], m->stream3) != cudaSuccess) { set_error("unet3d_prefetch_augmented: event"); return 1; }
    m->pf_mark(slot);
    return 0;
    GUARD_END
}

int unet3d_train_microbatch_prefetched(unet3d_t* h, int collapse_before, int use_ce, int use_dice, int use_mse, float loss_out3[3]) {
    GUARD_BEGIN NEED(h)
    Model* m = h->m;
    cudaSetDevice(m->device);
    float* din = nullptr;
    float* dlab = nullptr;
    if (m->consume_prefetched(&din, &dlab)) return 1;
    return m->train_microbatch(din, dlab, collapse_before, use_ce, use_dice, use_mse, loss_out3, nullptr, 1);
    GUARD_END
}

int simulate_modality(float* t1w, const float* label, unsigned max_label, unsigned seed, int w, int h, int d, int where, int gpu) {
    GUARD_BEGIN
    if (cudaSetDevice(gpu) != cudaSuccess) { set_error("no CUDA device: simulate_modality has no CPU fallback"); return 1; }
    if (!t1w) { set_error("simulate_modality: null image"); return 1; }
    std::lock_guard<std::mutex> lock(g_vpa_mu);
    VpaWs& W = g_vpa_ws[gpu];
    const size_t need = sample_ws_need(w, h, d, 1);
    if (!W.s && cudaStreamCreateWithFlags(&W.s, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return 1; }
    if (W.bytes < need) {
        if (W.p) cudaFree(W.p);
        W.bytes = 0;
        if (cudaMalloc(&W.p, need) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
        W.bytes = need;
    }
    int rc = sim_impl(t1w, label, max_label, seed, w, h, d, where, W.p, W.s, nullptr);
    if (!rc && where != 0 && cudaStreamSynchronize(W.s) != cudaSuccess) { set_error("simulate_modality: kernel failed"); rc = 1; }
    return rc;
    GUARD_END
}

int unet3d_simulate_modality(unet3d_t* h, float* t1w, const float* label, unsigned max_label, unsigned seed, int w, int hgt, int d,
                             int where) {
    GUARD_BEGIN NEED(h)
    Model* m = h->m;
    cudaSetDevice(m->device);
    if (!t1w) { set_error("simulate_modality: null image"); return 1; }
    if (ensure_model_ws(m, sample_ws_need(w, hgt, d, 1))) return 1;
    return sim_impl(t1w, label, max_label, seed, w, hgt, d, where, m->vpa_ws, m->stream, &m->launches);
    GUARD_END
}

int vpa_plan_describe(const char* const* keys, const float* vals, int n_opts, int is_label, int w, int h, int d, int channels,
                      uint64_t seed, float M12[12], float persp3[3], int* nfoci, float* foci5) {
    GUARD_BEGIN
    u3d::VpaPlan plan;
    if (u3d::vpa_make_plan(keys, vals, n_opts, is_label, w, h, d, channels, seed, plan)) return 1;
    if (M12)
        for (int i = 0; i < 12; ++i) M12[i] = plan.M[i];
    if (persp3)
        for (int i = 0; i < 3; ++i) persp3[i] = plan.has_persp ? plan.persp[i] : 0.f;
    if (nfoci) *nfoci = plan.nfoci;
    if (foci5)
        for (int f = 0; f < plan.nfoci; ++f) {
            foci5[5 * f + 0] = float(plan.foci[f].loc[0]);
            foci5[5 * f + 1] = float(plan.foci[f].loc[1]);
            foci5[5 * f + 2] = float(plan.foci[f].loc[2]);
            foci5[5 * f + 3] = plan.foci[f].radius;
            foci5[5 * f + 4] = plan.foci[f].mag;
        }
    return 0;
    GUARD_END
}

int vpa_plan_perlin(const char* const* keys, const float* vals, int n_opts, int is_label, int w, int h, int d, int channels,
                    uint64_t seed, int* applies, int* perm512, float* zoom) {
    GUARD_BEGIN
    u3d::VpaPlan plan;
    if (u3d::vpa_make_plan(keys, vals, n_opts, is_label, w, h, d, channels, seed, plan)) return 1;
    if (applies) *applies = plan.perlin;
    if (perm512 && plan.perlin)
        for (int i = 0; i < 512; ++i) perm512[i] = int(plan.perm[i]);
    if (zoom) *zoom = plan.perlin ? plan.zoom : 0.f;
    return 0;
    GUARD_END
}

int simulate_modality_plan(int labelled, unsigned max_label, unsigned seed, float* lut_out, float* terms_out, float* gamma_out) {
    GUARD_BEGIN
    u3d::SimPlan plan;
    if (u3d::simulate_make_plan(labelled, max_label, seed, 1, 1, 1, plan)) return 1;
    if (lut_out)
        for (int i = 0; i < plan.n_lut; ++i) lut_out[i] = plan.lut[i];
    if (terms_out)
        for (int t = 0; t < u3d::kSimTerms; ++t) {
            terms_out[5 * t + 0] = float(plan.a[t]);
            terms_out[5 * t + 1] = float(plan.b[t]);
            terms_out[5 * t + 2] = float(plan.c[t]);
            terms_out[5 * t + 3] = float(plan.d[t]);
            terms_out[5 * t + 4] = plan.w[t];
        }
    if (gamma_out) *gamma_out = plan.gamma;
    return 0;
    GUARD_END
}

int unet3d_set_simulate_modality(unet3d_t* h, int mode) {
    GUARD_BEGIN NEED(h)
    if (mode < 0 || mode > 2) { set_error("unet3d_set_simulate_modality: mode must be 0 (off), 1 (labelled template) or 2 (image only)"); return 1; }
    h->m->sim_mode = mode;
    return 0;
    GUARD_END
}

int unet3d_nccl_unique_id(void* id128) {
    GUARD_BEGIN return u3d::nccl_unique_id(id128);
    GUARD_END
}
int unet3d_nccl_comm_init(void** comm, int nranks, int rank, const void* id128) {
    GUARD_BEGIN return u3d::nccl_comm_init(comm, nranks, rank, id128);
    GUARD_END
}
int unet3d_nccl_comm_destroy(void* comm) { return u3d::nccl_comm_destroy(comm); }

int unet3d_attach_comm(unet3d_t* h, void* comm, int microbatches_per_step) {
    GUARD_BEGIN NEED(h) return h->m->attach_comm(comm, microbatches_per_step);
    GUARD_END
}

}  // extern "C"
