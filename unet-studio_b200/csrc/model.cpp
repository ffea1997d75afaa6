#include "model.h"
#include "postproc.h"

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace u3d {

#define M_CHECK(x)                  \
    do {                            \
        if ((x) != 0) return 1;     \
    } while (0)
#define M_CUDA(x)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            set_error(std::string(#x) + ": " + cudaGetErrorString(e_));                \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// feature string (grammar of unet.cpp:24-101, graph of unet.cpp:103-166)
// ------------------------------------------------------------------------------------------------
std::string default_feature(int out_count) {
    const std::string out = "conv" + std::to_string(out_count) + ",ks1,stride1";
    auto blk = [](int c, int s) {
        const std::string cs = std::to_string(c);
        return "conv" + cs + ",ks3,stride" + std::to_string(s) + "+norm,leaky_relu+conv" + cs + ",ks3,stride1+norm,leaky_relu";
    };
    std::string f;
    f += blk(16, 1) + "\n" + blk(32, 2) + "\n" + blk(64, 2) + "\n" + blk(128, 2) + "\n" + blk(256, 2) + "\n";
    f += blk(256, 2) + "+conv_trans256,ks2,stride2\n";
    f += blk(256, 1) + "+" + out + "+conv_trans128,ks2,stride2\n";
    f += blk(128, 1) + "+" + out + "+conv_trans64,ks2,stride2\n";
    f += blk(64, 1) + "+" + out + "+conv_trans32,ks2,stride2\n";
    f += blk(32, 1) + "+" + out + "+conv_trans16,ks2,stride2\n";
    f += blk(16, 1) + "+" + out;
    return f;
}

static std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream in(s);
    while (std::getline(in, cur, sep)) out.push_back(cur);
    return out;
}

static std::vector<std::string> split_lines(const std::string& s) {
    std::vector<std::string> out;
    for (auto& l : split(s, '\n')) {
        std::string t = l;
        while (!t.empty() && (t.back() == '\r' || t.back() == ' ')) t.pop_back();
        if (!t.empty()) out.push_back(t);
    }
    return out;
}

static int to_int(const std::string& s) {
    try {
        return std::stoi(s);
    } catch (...) {
        throw std::runtime_error("stoi");  // std::stoi's own what() in the reference
    }
}

// ---- token grammar -------------------------------------------------------------------------------------------------------------
// A token is a comma-separated attribute list; an attribute is NAME or NAME<digits...> (everything from the first digit on is the
// value; a bare name has the value "1").  The first rule of kLayerRules whose key is present decides the module, then the first
// activation key present appends an activation (same precedence as unet.cpp:38-98).
namespace {

class Token {
  public:
    explicit Token(const std::string& text) : text_(text) {
        for (const std::string& a : split(text, ',')) {
            const size_t digit = a.find_first_of("0123456789");
            if (digit == std::string::npos) attr_[a] = "1";
            else attr_[a.substr(0, digit)] = a.substr(digit);
        }
    }
    bool has(const char* key) const { return attr_.find(key) != attr_.end(); }
    int number(const char* key, int fallback) const {
        const auto it = attr_.find(key);
        return it == attr_.end() ? fallback : to_int(it->second);
    }
    // what the reference names in "unknown layer: ..." (unet.cpp:87): the first key of its std::unordered_map, or the raw text
    std::string some_key() const { return attr_.empty() ? text_ : attr_.begin()->first; }

  private:
    std::string text_;
    std::unordered_map<std::string, std::string> attr_;   // same container as the reference so begin() names the same key
};

struct LayerRule {
    const char* key;
    ModuleDef::Kind kind;
    bool has_channels;          // the key's number is the output channel count
    int ks, stride;             // defaults when the token does not say
    bool (*geometry_ok)(int ks, int stride);
    const char* geometry_error;
};
bool convt_geometry(int ks, int stride) { return ks == 2 && stride == 2; }
bool conv_geometry(int ks, int stride) { return (ks == 1 && stride == 1) || (ks == 3 && (stride == 1 || stride == 2)); }
const LayerRule kLayerRules[] = {
    {"max_pool", ModuleDef::MAXPOOL, false, 0, 0, nullptr, nullptr},
    {"upsample", ModuleDef::UPSAMPLE, false, 0, 0, nullptr, nullptr},
    {"conv_trans", ModuleDef::CONVT, true, 2, 2, convt_geometry, "conv_trans supports only ks2 stride2"},
    {"conv", ModuleDef::CONV, true, 3, 1, conv_geometry, "conv supports only ks1 stride1, ks3 stride1, and ks3 stride2"},
    {"norm", ModuleDef::NORM, false, 0, 0, nullptr, nullptr},
    {"bnorm", ModuleDef::BNORM, false, 0, 0, nullptr, nullptr},
};
const struct { const char* key; ModuleDef::Kind kind; } kActivationRules[] = {
    {"relu", ModuleDef::RELU}, {"leaky_relu", ModuleDef::LEAKY}, {"elu", ModuleDef::ELU}};

// Appends the modules one token stands for; returns the channel count behind it.
int append_token(BlockDef& blk, const std::string& text, int in_c) {
    const Token tok(text);
    const LayerRule* rule = nullptr;
    for (const LayerRule& r : kLayerRules)
        if (tok.has(r.key)) { rule = &r; break; }
    if (!rule) throw std::runtime_error("unknown layer: " + tok.some_key());
    ModuleDef m;
    m.kind = rule->kind;
    m.cin = in_c;
    m.cout = rule->has_channels ? tok.number(rule->key, 0) : in_c;
    if (rule->geometry_ok) {
        m.ks = tok.number("ks", rule->ks);
        m.stride = tok.number("stride", rule->stride);
        if (!rule->geometry_ok(m.ks, m.stride)) throw std::runtime_error(rule->geometry_error);
    }
    blk.mods.push_back(m);
    for (const auto& a : kActivationRules)
        if (tok.has(a.key)) {
            ModuleDef act;
            act.kind = a.kind;
            act.cin = act.cout = m.cout;
            blk.mods.push_back(act);
            break;
        }
    return m.cout;
}

}  // namespace

// Line i of the feature string is encoder level i for the first n/2+1 lines; the remaining lines are the decoder levels from the
// deepest up (unet.cpp:103-166).  In a decoder line the token equal to the LAST token of the LAST line is that level's output
// head; tokens in front of it form the level's decoder block (fed by cat{skip, x}), tokens behind it the up-sampling tail.
Model::Model(int in_c, int out_c, const std::string& feature, bool host_only_)
    : host_only(host_only_), in_count(in_c), out_count(out_c), architecture(feature) {
    const std::vector<std::string> lines = split_lines(feature);
    if (lines.size() < 3) throw std::runtime_error("invalid u-net structure");
    const int n_enc = int(lines.size()) / 2 + 1, n_dec = int(lines.size()) - n_enc;
    const std::vector<std::string> last_line = split(lines.back(), '+');
    if (last_line.empty()) throw std::runtime_error("invalid u-net structure");
    const std::string head_token = last_line.back();
    encoding.resize(n_enc);
    decoding.resize(n_dec);
    output.resize(n_dec);
    tail.resize(n_dec);
    std::vector<int> width_at(n_enc, 0);   // channels leaving encoder level l (= the skip connection's width)
    int width = in_c;
    for (int li = 0; li < int(lines.size()); ++li) {
        const std::vector<std::string> pieces = split(lines[li], '+');
        if (li < n_enc) {
            encoding[li].name = "encode" + std::to_string(li);
            for (const std::string& t : pieces) width = append_token(encoding[li], t, width);
            width_at[li] = width;
            continue;
        }
        const int level = int(lines.size()) - 1 - li;
        const std::string tag = std::to_string(level);
        decoding[level].name = "decode" + tag;
        output[level].name = "output" + tag;
        tail[level].name = "decode_tail" + tag;
        width += width_at[level];
        BlockDef* sink = &decoding[level];
        for (const std::string& t : pieces) {
            if (t == head_token) {
                append_token(output[level], t, width);   // the head does not change the width of the trunk
                sink = &tail[level];
            } else
                width = append_token(*sink, t, width);
        }
    }
    // registration order == parameters() order == tensorN order (unet.cpp:130,160-164)
    long long off = 0;
    auto reg = [&](BlockDef& blk) {
        for (size_t i = 0; i < blk.mods.size(); ++i) {
            ModuleDef& m = blk.mods[i];
            std::vector<std::vector<int64_t>> shapes;
            if (m.kind == ModuleDef::CONV) shapes = {{m.cout, m.cin, m.ks, m.ks, m.ks}, {m.cout}};
            else if (m.kind == ModuleDef::CONVT) shapes = {{m.cin, m.cout, 2, 2, 2}, {m.cout}};
            else if (m.kind == ModuleDef::NORM || m.kind == ModuleDef::BNORM) shapes = {{m.cin}, {m.cin}};
            else continue;
            m.p0 = int(params.size());
            for (int k = 0; k < 2; ++k) {
                ParamInfo p;
                p.name = blk.name + "." + std::to_string(i) + (k ? ".bias" : ".weight");
                p.shape = shapes[k];
                p.numel = 1;
                for (auto d : p.shape) p.numel *= d;
                p.offset = off;
                off += (p.numel + 3) / 4 * 4;
                p.decay = !(p.name.find("bias") != std::string::npos || p.shape.size() <= 1);
                params.push_back(p);
            }
            if (m.kind == ModuleDef::BNORM) {
                m.buf0 = n_buffers;
                n_buffers += 2;
            }
        }
    };
    for (auto& b : encoding) reg(b);
    for (int l = n_dec - 1; l >= 0; --l) {
        reg(decoding[l]);
        if (!output[l].mods.empty()) reg(output[l]);
        if (!tail[l].mods.empty()) reg(tail[l]);
    }
    flat_n = off;

    if (host_only) return;
    cudaGetDevice(&device);
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess)
        throw std::runtime_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(cudaGetLastError()) +
                                 " (libunet3d_b200 needs a CUDA device; there is no CPU fallback)");
    if (cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev_pack, cudaEventDisableTiming) != cudaSuccess)
        throw std::runtime_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(cudaGetLastError()));
    const size_t fb = size_t(flat_n) * sizeof(float);
    if (cudaMalloc(&d_params, fb) != cudaSuccess || cudaMalloc(&d_grads, fb) != cudaSuccess || cudaMalloc(&d_mom, fb) != cudaSuccess)
        throw std::runtime_error("cudaMalloc of the parameter buffers failed");
    cudaMemsetAsync(d_params, 0, fb, stream);
    cudaMemsetAsync(d_grads, 0, fb, stream);
    cudaMemsetAsync(d_mom, 0, fb, stream);
    // BatchNorm running stats: mean 0, var 1
    auto bufs = [&](BlockDef& blk) {
        for (auto& m : blk.mods)
            if (m.kind == ModuleDef::BNORM)
                for (int k = 0; k < 2; ++k) {
                    float* b = nullptr;
                    cudaMalloc(&b, size_t(m.cin) * 4);
                    std::vector<float> h(m.cin, k ? 1.f : 0.f);
                    cudaMemcpy(b, h.data(), size_t(m.cin) * 4, cudaMemcpyHostToDevice);
                    d_buffers.push_back(b);
                    buffer_len.push_back(m.cin);
                }
    };
    for (auto& b : encoding) bufs(b);
    for (int l = n_dec - 1; l >= 0; --l) { bufs(decoding[l]); bufs(output[l]); bufs(tail[l]); }
    // optimizer work list
    std::vector<SgdChunk> chunks;
    for (auto& p : params)
        for (long long o = 0; o < p.numel; o += 8192)
            chunks.push_back(SgdChunk{p.offset + o, int(std::min<long long>(8192, p.numel - o)), p.decay ? 3e-5f : 0.f});
    n_chunks = int(chunks.size());
    cudaMalloc(&d_chunks, chunks.size() * sizeof(SgdChunk));
    cudaMemcpy(d_chunks, chunks.data(), chunks.size() * sizeof(SgdChunk), cudaMemcpyHostToDevice);
    cudaMalloc(&d_status, sizeof(SgdStatus));
    cudaMalloc(&d_counter, 16);
    cudaMemsetAsync(d_counter, 0, 16, stream);
    const size_t nl = std::max<size_t>(output.size(), 1);   // per-level loss scratch, sized from the number of deep-supervision heads
    cudaMalloc(&d_loss_acc, sizeof(double) * nl * 80);
    cudaMalloc(&d_loss_part, sizeof(float) * nl * size_t(loss_part_rows()) * loss_part_cols());
    cudaMalloc(&d_losses, sizeof(float) * nl * 3 * 2);
}

Model::~Model() {
    if (host_only) return;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (stream2) cudaStreamSynchronize(stream2);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_pack) cudaEventDestroy(ev_pack);
    if (stream4) { cudaStreamSynchronize(stream4); cudaStreamDestroy(stream4); }
    if (ev_ar_ready) cudaEventDestroy(ev_ar_ready);
    if (ev_ar_done) cudaEventDestroy(ev_ar_done);
    if (h_val) cudaFreeHost(h_val);
    if (ev_status) cudaEventDestroy(ev_status);
    if (h_status) cudaFreeHost(h_status);
    free_plan();
    cudaFree(d_params); cudaFree(d_grads); cudaFree(d_mom);
    if (vpa_ws) cudaFree(vpa_ws);
    if (pf_ws) cudaFree(pf_ws);
    for (auto b : d_buffers) cudaFree(b);
    cudaFree(d_counter);
    cudaFree(d_chunks); cudaFree(d_status); cudaFree(d_loss_acc); cudaFree(d_loss_part); cudaFree(d_losses);
    if (stream3) { cudaStreamSynchronize(stream3); cudaStreamDestroy(stream3); }
    for (cudaEvent_t e : ev_slabs) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { if (ev_sample[i]) cudaEventDestroy(ev_sample[i]); if (pf_in[i]) cudaFree(pf_in[i]); }
    for (int i = 0; i < 2; ++i) { if (ew_in[i]) cudaFree(ew_in[i]); if (ew_out[i]) cudaFree(ew_out[i]); }
    if (ev_buf) cudaFree(ev_buf);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
}

int Model::alloc(void** p, size_t bytes) {
    M_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    owned.push_back(*p);
    return 0;
}

void Model::drop_graphs() {
    for (auto& g : graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
}

int Model::graph_run(const std::vector<uint64_t>& key, const std::function<int()>& body) {
    static const bool disabled = std::getenv("U3D_NO_GRAPH") != nullptr;
    if (disabled || prof_on) return body();
    GraphEntry* e = nullptr;
    for (auto& g : graphs)
        if (g.key == key) { e = &g; break; }
    if (!e) {   // first sight: enqueue normally (lazy allocations, function attributes, tensor-map encodes happen here)
        if (graphs.size() >= 16) drop_graphs();
        GraphEntry n;
        n.key = key;
        graphs.push_back(n);
        return body();
    }
    if (e->bad) return body();
    if (!e->exec) {
        const long long l0 = launches;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); e->bad = true; return body(); }
        const int rc = body();
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(stream, &g);
        if (rc != 0 || ce != cudaSuccess || g == nullptr) {
            if (std::getenv("U3D_GRAPH_DEBUG")) std::fprintf(stderr, "u3d: graph capture failed (rc %d, %s)\n", rc, cudaGetErrorString(ce));
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            e->bad = true;
            launches = l0;
            if (rc != 0) return 1;
            return body();   // nothing ran during the failed capture
        }
        const cudaError_t ie = cudaGraphInstantiate(&e->exec, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { cudaGetLastError(); e->exec = nullptr; e->bad = true; launches = l0; return body(); }
        e->n_launches = launches - l0;
        launches = l0;
        if (std::getenv("U3D_GRAPH_DEBUG")) std::fprintf(stderr, "u3d: captured a graph of %lld launches (kind %llu)\n", e->n_launches, (unsigned long long)key[0]);
    }
    M_CUDA(cudaGraphLaunch(e->exec, stream));
    launches += e->n_launches;
    return 0;
}

void Model::free_plan() {
    drop_graphs();
    if (stream2) cudaStreamSynchronize(stream2);   // an asynchronous re-pack may still be writing the blobs freed below
    pack_pending = false;
    for (void* p : owned) cudaFree(p);
    owned.clear();
    tens.clear();
    steps.clear();
    logits.clear();
    dlogits.clear();
    level_dims.clear();
    d_in_f32 = d_label = d_partials = d_sums = nullptr;
    d_pack_descs = nullptr; d_pack_first = nullptr; n_pack_jobs = n_pack_blocks = 0;
    d_scratch = nullptr;
    d_wgrad_partial = nullptr; wgrad_partial_bytes = 0;
    d_splitk = nullptr;
    splitk_bytes = 0;
    planned = false;
}

// ------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------
int Model::init_params(uint64_t seed) {
    // torch default init rule (kaiming_uniform a=sqrt(5) => U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and
    // bias; norm gamma 1, beta 0).  Own RNG stream: bit-parity with torch::manual_seed is not a goal, the
    // parity tests load the reference's dumped tensors through set_param instead.
    std::mt19937_64 rng(seed);
    std::vector<float> h(size_t(flat_n), 0.f);
    for (size_t i = 0; i < params.size(); ++i) {
        const ParamInfo& p = params[i];
        float* dst = h.data() + p.offset;
        if (p.shape.size() == 5) {
            const double fan_in = double(p.shape[1] * p.shape[2] * p.shape[3] * p.shape[4]);
            const double bound = 1.0 / std::sqrt(fan_in);
            std::uniform_real_distribution<double> U(-bound, bound);
            for (long long k = 0; k < p.numel; ++k) dst[k] = float(U(rng));
            const ParamInfo& b = params[i + 1];
            float* bd = h.data() + b.offset;
            for (long long k = 0; k < b.numel; ++k) bd[k] = float(U(rng));
            ++i;
        } else {
            const bool gamma = p.name.size() >= 6 && p.name.compare(p.name.size() - 6, 6, "weight") == 0;
            for (long long k = 0; k < p.numel; ++k) dst[k] = gamma ? 1.f : 0.f;
        }
    }
    cudaSetDevice(device);
    M_CUDA(cudaMemcpyAsync(d_params, h.data(), h.size() * 4, cudaMemcpyHostToDevice, stream));
    M_CUDA(cudaStreamSynchronize(stream));
    packs_dirty = true;
    return 0;
}

int Model::get_flat(const float* base, int i, float* host, float scale) {
    if (i < 0 || i >= int(params.size())) { set_error("parameter index out of range"); return 1; }
    cudaSetDevice(device);
    M_CUDA(cudaMemcpyAsync(host, base + params[i].offset, size_t(params[i].numel) * 4, cudaMemcpyDeviceToHost, stream));
    M_CUDA(cudaStreamSynchronize(stream));
    if (scale != 1.f)
        for (long long k = 0; k < params[i].numel; ++k) host[k] *= scale;
    return 0;
}

int Model::set_param(int i, const float* host) {
    if (i < 0 || i >= int(params.size())) { set_error("parameter index out of range"); return 1; }
    cudaSetDevice(device);
    M_CUDA(cudaMemcpyAsync(d_params + params[i].offset, host, size_t(params[i].numel) * 4, cudaMemcpyHostToDevice, stream));
    M_CUDA(cudaStreamSynchronize(stream));
    packs_dirty = true;
    return 0;
}

int Model::set_momentum(int i, const float* host) {
    if (i < 0 || i >= int(params.size())) { set_error("parameter index out of range"); return 1; }
    cudaSetDevice(device);
    M_CUDA(cudaMemcpyAsync(d_mom + params[i].offset, host, size_t(params[i].numel) * 4, cudaMemcpyHostToDevice, stream));
    M_CUDA(cudaStreamSynchronize(stream));
    mom_initialized = true;
    return 0;
}

int Model::set_dim(int w, int h, int d) {
    if (w <= 0 || h <= 0 || d <= 0) { set_error("invalid dimension"); return 1; }
    if (w != dim[0] || h != dim[1] || d != dim[2]) {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        if (stream3) cudaStreamSynchronize(stream3);
        pf_pending[0] = pf_pending[1] = false;   // prefetched samples of the old grid are dropped
        pf_next = pf_head = 0;
        free_plan();
    }
    dim[0] = w; dim[1] = h; dim[2] = d;
    return 0;
}

int Model::set_mode(int mode) {
    if (mode < 0 || mode > 2) { set_error("set_mode: 0 = prepare_for_inference, 1 = train, 2 = eval"); return 1; }
    const bool train = mode == 1;
    cudaSetDevice(device);
    if (train != training) {
        cudaStreamSynchronize(stream);
        free_plan();
    }
    training = train;
    bn_running = mode == 2;
    if (mode == 0) {   // unet.cpp:14-21: running_mean.zero_(), running_var.fill_(1)
        for (size_t i = 0; i < d_buffers.size(); ++i) {
            std::vector<float> h(size_t(buffer_len[i]), (i & 1) ? 1.f : 0.f);
            M_CUDA(cudaMemcpyAsync(d_buffers[i], h.data(), h.size() * 4, cudaMemcpyHostToDevice, stream));
            M_CUDA(cudaStreamSynchronize(stream));
        }
    }
    return 0;
}

// U3D_TRACE_LAUNCHES=<file>: one line per tensor-kernel launch site (layer, pass, kernel family, problem count, algorithmic FLOPs), in
// launch order -- joined with an ncu launch list of the same run (one stream, no graphs) it gives the per-layer tensor-pipe table
void Model::trace_launch(const Step& s, const char* pass, int kind, int nprob, double flops) {
    static const char* path = std::getenv("U3D_TRACE_LAUNCHES");
    if (!path) return;
    static FILE* f = std::fopen(path, "w");
    if (!f) return;
    static const char* fam[] = {"conv_igemm", "conv_wgrad", "conv_s2", "conv_wgrad_band", "conv_tma", "conv_band", "conv_wgrad_quad", "?"};
    const ParamInfo& p = params[size_t(s.p_w)];
    std::fprintf(f, "%s\t%s\t%s\t%d\t%.6g\tcin=%d+%d cout=%d k%d s%d%s out=%dx%dx%d\n", p.name.c_str(), pass, fam[kind & 7], nprob, flops, s.g.cin[0],
                 s.g.cin[1], s.g.cout, s.g.ks, s.g.stride, s.g.transposed ? " transposed" : "", s.g.out_w, s.g.out_h, s.g.out_d);
    std::fflush(f);
    trace_marker_launch(std::strcmp(pass, "wgrad") == 0 && stream2 && !std::getenv("U3D_ONE_STREAM") ? stream2 : stream);
}

void Model::prof_begin(int kind, double flops, cudaStream_t on) {
    if (!prof_on) return;
    if (prof_used + 2 > prof_ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        prof_ev.push_back(a);
        prof_ev.push_back(b);
        prof_kind.push_back(0);
        prof_flops.push_back(0);
    }
    prof_kind[prof_used / 2] = kind;
    prof_flops[prof_used / 2] = flops;
    cudaEventRecord(prof_ev[prof_used], on ? on : stream);
}
void Model::prof_end(cudaStream_t on) {
    if (!prof_on) return;
    cudaEventRecord(prof_ev[prof_used + 1], on ? on : stream);
    prof_used += 2;
}
int Model::prof_read(double out[24], int reset) {
    cudaSetDevice(device);
    M_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 24; ++i) out[i] = 0;
    for (size_t i = 0; i + 1 < prof_used; i += 2) {
        float ms = 0.f;
        M_CUDA(cudaEventElapsedTime(&ms, prof_ev[i], prof_ev[i + 1]));
        const int k = 3 * std::min(prof_kind[i / 2], 7);
        out[k] += ms;
        out[k + 1] += 1;
        out[k + 2] += prof_flops[i / 2];
    }
    if (reset) prof_used = 0;
    return 0;
}

int Model::timer_start() {
    cudaSetDevice(device);
    if (!ev0) { M_CUDA(cudaEventCreate(&ev0)); M_CUDA(cudaEventCreate(&ev1)); }
    M_CUDA(cudaEventRecord(ev0, stream));
    return 0;
}
int Model::timer_stop(float* ms) {
    cudaSetDevice(device);
    if (!ev0) { set_error("timer_stop without timer_start"); return 1; }
    M_CUDA(cudaEventRecord(ev1, stream));
    M_CUDA(cudaEventSynchronize(ev1));
    M_CUDA(cudaEventElapsedTime(ms, ev0, ev1));
    return 0;
}

int Model::sync() {
    cudaSetDevice(device);
    M_CUDA(cudaStreamSynchronize(stream));
    const unsigned int code = read_device_error();
    if (code) { set_error("device pipeline timeout code " + std::to_string(code)); return 1; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// graph -> steps
// ------------------------------------------------------------------------------------------------
int Model::build_steps() {
    tens.clear();
    steps.clear();
    auto new_ten = [&](int C, int d, int h, int w, bool grad) {
        Ten t;
        t.C = C; t.Cp = pad16(C); t.d = d; t.h = h; t.w = w; t.needs_grad = grad;
        tens.push_back(t);
        return int(tens.size()) - 1;
    };
    int cur = new_ten(in_count, dim[2], dim[1], dim[0], false);
    bool fail = false;
    std::string why;
    auto run_block = [&](const BlockDef& blk, int in_a, int in_b, int head_level) -> int {
        int x = in_a, x2 = in_b;
        for (size_t i = 0; i < blk.mods.size() && !fail; ++i) {
            const ModuleDef& m = blk.mods[i];
            const Ten a = tens[x];
            if (x2 >= 0 && m.kind != ModuleDef::CONV) {
                fail = true;
                why = "a decoder level must start with a conv token (the channel concat is folded into it)";
                break;
            }
            if (m.kind == ModuleDef::CONV || m.kind == ModuleDef::CONVT) {
                Step s;
                s.kind = Step::CONV;
                s.in0 = x; s.in1 = x2;
                LayerGeom& g = s.g;
                g.transposed = m.kind == ModuleDef::CONVT;
                g.ks = m.ks; g.stride = m.stride;
                g.cin[0] = a.C; g.cin[1] = x2 >= 0 ? tens[x2].C : 0;
                g.cout = m.cout;
                g.in_d = a.d; g.in_h = a.h; g.in_w = a.w;
                if (x2 >= 0 && (tens[x2].d != a.d || tens[x2].h != a.h || tens[x2].w != a.w)) {
                    fail = true;
                    why = "skip connection and up-sampled tensor differ in size (torch::cat would throw): use a grid that is a multiple of 2^levels";
                    break;
                }
                if (g.cin[0] + g.cin[1] != m.cin) { fail = true; why = "internal channel mismatch"; break; }
                if (g.transposed) { g.out_d = 2 * a.d; g.out_h = 2 * a.h; g.out_w = 2 * a.w; }
                else {
                    const int pad = (m.ks - 1) / 2;
                    g.out_d = (a.d + 2 * pad - m.ks) / m.stride + 1;
                    g.out_h = (a.h + 2 * pad - m.ks) / m.stride + 1;
                    g.out_w = (a.w + 2 * pad - m.ks) / m.stride + 1;
                }
                if (g.out_d <= 0 || g.out_h <= 0 || g.out_w <= 0) { fail = true; why = "volume too small for the network"; break; }
                s.p_w = m.p0; s.p_b = m.p0 + 1;
                const bool next_norm = i + 1 < blk.mods.size() &&
                                       (blk.mods[i + 1].kind == ModuleDef::NORM || (blk.mods[i + 1].kind == ModuleDef::BNORM && training));
                s.stats = next_norm && !g.transposed;
                s.drop_bias = head_level < 0 && i + 1 < blk.mods.size() && blk.mods[i + 1].kind == ModuleDef::NORM;
                if (head_level >= 0) {
                    if (blk.mods.size() != 1) { fail = true; why = "the output token must be a single conv"; break; }
                    s.head_level = head_level;
                    s.out = -1;
                    level_dims[3 * head_level] = g.out_d; level_dims[3 * head_level + 1] = g.out_h; level_dims[3 * head_level + 2] = g.out_w;
                } else {
                    s.out = new_ten(m.cout, g.out_d, g.out_h, g.out_w, true);
                    x = s.out;
                }
                x2 = -1;
                steps.push_back(s);
            } else if (m.kind == ModuleDef::MAXPOOL || m.kind == ModuleDef::UPSAMPLE) {
                Step s;
                s.kind = m.kind == ModuleDef::MAXPOOL ? Step::MAXPOOL : Step::UPSAMPLE;
                s.in0 = x;
                if (m.kind == ModuleDef::MAXPOOL) {
                    if (a.d / 2 <= 0 || a.h / 2 <= 0 || a.w / 2 <= 0) { fail = true; why = "volume too small for max_pool"; break; }
                    s.out = new_ten(a.C, a.d / 2, a.h / 2, a.w / 2, true);
                } else
                    s.out = new_ten(a.C, a.d * 2, a.h * 2, a.w * 2, true);
                x = s.out;
                steps.push_back(s);
            } else {
                Step s;
                s.kind = Step::NORMACT;
                s.in0 = x;
                auto act_of = [](ModuleDef::Kind k) { return k == ModuleDef::RELU ? ACT_RELU : k == ModuleDef::LEAKY ? ACT_LEAKY : ACT_ELU; };
                if (m.kind == ModuleDef::NORM || m.kind == ModuleDef::BNORM) {
                    s.norm = m.kind == ModuleDef::NORM ? 1 : 2;
                    s.p_g = m.p0;
                    s.buf0 = m.buf0;
                    if (i + 1 < blk.mods.size() && blk.mods[i + 1].kind >= ModuleDef::RELU) {
                        s.act = act_of(blk.mods[i + 1].kind);
                        ++i;
                    }
                    s.stats_from_conv = !steps.empty() && steps.back().kind == Step::CONV && steps.back().stats && steps.back().out == x;
                } else
                    s.act = act_of(m.kind);
                s.out = new_ten(a.C, a.d, a.h, a.w, true);
                x = s.out;
                steps.push_back(s);
            }
        }
        return x;
    };
    const int E = int(encoding.size());
    level_dims.assign(3 * output.size(), 0);
    std::vector<int> skips(E, -1);
    for (int l = 0; l < E && !fail; ++l) {
        cur = run_block(encoding[l], cur, -1, -1);
        if (l < E - 1) skips[l] = cur;
    }
    for (int l = E - 2; l >= 0 && !fail; --l) {
        cur = run_block(decoding[l], skips[l], cur, -1);
        if (!fail && !output[l].mods.empty()) run_block(output[l], cur, -1, l);
        if (!fail && !tail[l].mods.empty()) cur = run_block(tail[l], cur, -1, -1);
    }
    if (fail) { set_error(why); return 1; }
    // first-layer precision: when every consumer of the network input is a conv reading it as its first source and
    // the padded channels have room, the input is stored as fp16 hi + lo pairs and the consumer's weights are packed to match
    static const bool no_split = std::getenv("U3D_NO_SPLIT_INPUT") != nullptr;
    split_input = !no_split && 3 * in_count <= tens[0].Cp;
    for (const Step& s : steps) {
        if (s.in1 == 0) split_input = false;
        if (s.in0 == 0 && s.kind != Step::CONV) split_input = false;
    }
    return 0;
}

int Model::ensure_plan() {
    cudaSetDevice(device);
    if (planned && planned_training == training) return 0;
    free_plan();
    M_CHECK(build_steps());
    const bool tr = training;
    size_t max_bytes = 0;
    int max_cp = 16;
    for (auto& t : tens) {
        M_CHECK(alloc(&t.p, t.bytes()));
        if (tr && t.needs_grad) M_CHECK(alloc(&t.grad, t.bytes()));
        max_bytes = std::max(max_bytes, t.bytes());
        max_cp = std::max(max_cp, t.Cp);
    }
    // padded channels of the network input must be zero; every other tensor is fully written by its producer
    const long long V0 = tens[0].V();
    M_CHECK(alloc(reinterpret_cast<void**>(&d_in_f32), size_t(in_count) * V0 * 4));
    M_CHECK(alloc(reinterpret_cast<void**>(&d_label), size_t(V0) * 4));
    const int rows = std::max(reduce_rows_max(), device_sm_count());
    M_CHECK(alloc(reinterpret_cast<void**>(&d_partials), size_t(rows) * 2 * std::max(max_cp, 512) * 4));
    M_CHECK(alloc(reinterpret_cast<void**>(&d_sums), size_t(2) * max_cp * 4));
    if (tr) {
        M_CHECK(alloc(&d_scratch, max_bytes));
        scratch_bytes = max_bytes;
        wgrad_partial_bytes = std::max(conv_wgrad_band_scratch_bytes(), size_t(64) << 20);   // (the generic kernel: items x 128 x ntile x 4 bytes)
        M_CHECK(alloc(reinterpret_cast<void**>(&d_wgrad_partial), wgrad_partial_bytes));
    }
    splitk_bytes = size_t(96) << 20;   // 16 slices x (< 74 boxes x 120 voxels) x 256 channels x 4 B fits with room to spare
    M_CHECK(alloc(reinterpret_cast<void**>(&d_splitk), splitk_bytes));
    const int L = int(output.size());
    logits.assign(L, nullptr);
    dlogits.assign(L, nullptr);
    const int ocp = pad16(out_count);
    for (int l = 0; l < L; ++l) {
        const long long v = 1LL * level_dims[3 * l] * level_dims[3 * l + 1] * level_dims[3 * l + 2];
        if (v == 0) continue;
        M_CHECK(alloc(reinterpret_cast<void**>(&logits[l]), size_t(out_count) * v * 4));
        if (tr) M_CHECK(alloc(&dlogits[l], size_t(ocp) * v * 2));
    }
    head_step.assign(L, -1);
    for (size_t si = 0; si < steps.size(); ++si) {
        Step& s = steps[si];
        if (s.kind == Step::CONV && s.head_level >= 0) {
            head_step[s.head_level] = int(si);
            static const bool nofuse = std::getenv("U3D_NO_HEAD_FUSE") != nullptr;
            const int xcp = tens[s.in0].Cp;
            s.head_fwd_fused = !nofuse && s.g.ks == 1 && s.in1 < 0 && !s.g.transposed && head_fwd_supported(s.g.cout, xcp);
            s.head_bwd_fused = s.head_fwd_fused && tr && tens[s.in0].needs_grad && head_bwd_supported(s.g.cout, xcp);
        }
    }
    for (auto& s : steps) {
        if (s.kind == Step::CONV) {
            {
                const long long vox = 1LL * s.g.out_d * s.g.out_h * s.g.out_w;
                const bool halo = s.head_level < 0 && conv_band_wants_kc16(s.g.ks, s.g.stride, s.g.transposed,
                                                                           pad16(s.g.cin[0]) + (s.g.cin[1] ? pad16(s.g.cin[1]) : 0), pad16(s.g.cout), vox);
                const bool s2 = s.head_level < 0 && conv_s2_wants_kc16(s.g.ks, s.g.stride, s.g.transposed, pad16(s.g.cin[0]), s.g.cin[1] ? 2 : 1,
                                                                       pad16(s.g.cout), vox);
                plan_forward(s.g, s.fprobs, s.fpacks, s.fkc, (halo || s2) ? 16 : 0);
            }
            s.flops = 2.0 * double(s.g.cin[0] + s.g.cin[1]) * s.g.cout * (s.g.transposed ? 1.0 : double(s.g.ks * s.g.ks * s.g.ks)) *
                      double(s.g.out_d) * s.g.out_h * s.g.out_w;
            for (size_t i = 0; i < s.fprobs.size(); ++i) {
                void* blob = nullptr;
                M_CHECK(alloc(&blob, pack_bytes(s.fpacks[i])));
                s.pack_bufs.push_back(blob);
                s.fpacks[i].w = param_ptr(s.p_w);
                s.fpacks[i].out = blob;
                s.fpacks[i].out_bf16 = 0;
                if (split_input && s.in0 == 0) s.fpacks[i].split_k = in_count;
                ConvProblem& P = s.fprobs[i];
                P.src0 = tens[s.in0].p;
                P.c0p = tens[s.in0].Cp;
                if (s.in1 >= 0) { P.src1 = tens[s.in1].p; P.c1p = tens[s.in1].Cp; }
                P.dst = s.head_level >= 0 ? static_cast<void*>(logits[s.head_level]) : tens[s.out].p;
                P.wpack = blob;
                P.bias = s.drop_bias ? nullptr : param_ptr(s.p_b);
            }
            if (tr) {
                void* dy = s.head_level >= 0 ? dlogits[s.head_level] : tens[s.out].grad;
                const int ins[2] = {s.in0, s.in1};
                for (int src = 0; src < 2; ++src) {
                    if (ins[src] < 0) continue;
                    WgradProblem W;
                    plan_wgrad(s.g, src, W);
                    if (!s.g.transposed) { W.T = tens[ins[src]].p; W.t_cp = tens[ins[src]].Cp; W.U = dy; }
                    else { W.T = dy; W.U = tens[s.in0].p; W.u_cp = tens[s.in0].Cp; }
                    W.dw = grad_ptr(s.p_w);
                    s.wg.push_back(W);
                    if (!tens[ins[src]].needs_grad) continue;
                    Step::DG& D = s.dg[src];
                    {
                        const long long vox = 1LL * s.g.in_d * s.g.in_h * s.g.in_w;
                        const bool halo = conv_band_wants_kc16(s.g.ks, s.g.stride, s.g.transposed, pad16(s.g.cout), pad16(s.g.cin[src]), vox);
                        plan_dgrad(s.g, src, D.probs, D.packs, D.kc, halo ? 16 : 0);
                    }
                    for (size_t i = 0; i < D.probs.size(); ++i) {
                        void* blob = nullptr;
                        M_CHECK(alloc(&blob, pack_bytes(D.packs[i])));
                        s.pack_bufs.push_back(blob);
                        D.packs[i].w = param_ptr(s.p_w);
                        D.packs[i].out = blob;
                        D.packs[i].out_bf16 = 0;
                        D.probs[i].src0 = dy;
                        D.probs[i].dst = tens[ins[src]].grad;
                        D.probs[i].dst_cp = tens[ins[src]].Cp;
                        D.probs[i].wpack = blob;
                    }
                }
            }
        } else if (s.kind == Step::NORMACT) {
            if (s.norm) {
                M_CHECK(alloc(reinterpret_cast<void**>(&s.mean), size_t(tens[s.in0].Cp) * 4));
                M_CHECK(alloc(reinterpret_cast<void**>(&s.rstd), size_t(tens[s.in0].Cp) * 4));
            }
        } else if (s.kind == Step::MAXPOOL) {
            M_CHECK(alloc(reinterpret_cast<void**>(&s.idx), tens[s.out].V() * tens[s.out].Cp * sizeof(int)));
        }
    }
    // Norm + activation folded into the output head: when the only reader of an activated tensor is the CUDA-core head
    // kernel, the norm_act_fwd pass over that full-resolution tensor is dropped; head_fwd takes the raw conv output and applies
    // scale/shift/activation itself (in training it also stores the activated voxels: the fused head backward reads them).
    // The same fold into the tensor-core consumers (conv_band / conv_s2 producers) was built and measured slower than the separate pass
    // -- those kernels are bound by shared-memory operand fetch and producer latency, see DESIGN.md 10.
    static const bool no_xf = std::getenv("U3D_NO_XF") != nullptr;
    for (size_t ni = 0; ni < steps.size() && !no_xf; ++ni) {
        Step& n = steps[ni];
        if (n.kind != Step::NORMACT) continue;
        int reader = -1, nreaders = 0;
        for (size_t ci = 0; ci < steps.size(); ++ci)
            if (steps[ci].in0 == n.out || steps[ci].in1 == n.out) { reader = int(ci); ++nreaders; }
        if (nreaders != 1) continue;
        Step& c = steps[reader];
        if (c.kind != Step::CONV || c.head_level < 0 || !c.head_fwd_fused || c.in0 != n.out) continue;
        n.elided = true;
        c.xf_from = int(ni);
    }
    {   // concat convs whose source 0 (the skip tensor) is also consumed by another conv: their skip gradient can be deferred
        skip_has_other_consumer.assign(steps.size(), 0);
        for (size_t ci = 0; ci < steps.size(); ++ci) {
            const Step& c = steps[ci];
            if (c.kind != Step::CONV || c.in1 < 0) continue;
            for (size_t oi = 0; oi < steps.size(); ++oi)
                if (oi != ci && steps[oi].kind == Step::CONV && oi < ci && (steps[oi].in0 == c.in0 || steps[oi].in1 == c.in0)) skip_has_other_consumer[ci] = 1;
        }
    }
    {   // job table of the one-launch weight re-pack
        std::vector<PackDesc> jobs;
        std::vector<int> first(1, 0);
        for (auto& s : steps) {
            if (s.kind != Step::CONV) continue;
            for (auto& k : s.fpacks) jobs.push_back(k);
            for (int src = 0; src < 2; ++src)
                for (auto& k : s.dg[src].packs) jobs.push_back(k);
        }
        for (auto& k : jobs) {
            k.force_elementwise = pack_force_elementwise() ? 1 : 0;
            first.push_back(first.back() + pack_job_blocks(k));
        }
        n_pack_jobs = int(jobs.size());
        n_pack_blocks = first.back();
        d_pack_descs = nullptr; d_pack_first = nullptr;
        if (n_pack_jobs) {
            M_CHECK(alloc(reinterpret_cast<void**>(&d_pack_descs), jobs.size() * sizeof(PackDesc)));
            M_CHECK(alloc(reinterpret_cast<void**>(&d_pack_first), first.size() * sizeof(int)));
            M_CUDA(cudaMemcpyAsync(d_pack_descs, jobs.data(), jobs.size() * sizeof(PackDesc), cudaMemcpyHostToDevice, stream));
            M_CUDA(cudaMemcpyAsync(d_pack_first, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            M_CUDA(cudaStreamSynchronize(stream));   // the host vectors go out of scope
        }
    }
    grad_written.assign(tens.size(), 0);
    // data-parallel overlap split: the first conv (forward order) that has >= 5 % of the parameters in front of it.  In the default net
    // that is encode4's first conv: levels 0-3 hold 6 % of the parameters but the last ~20 % of the backward time.
    dp_split_step = -1;
    dp_split = 0;
    for (size_t si = 0; si < steps.size(); ++si) {
        if (steps[si].kind != Step::CONV || steps[si].head_level >= 0) continue;
        const long long off = params[steps[si].p_w].offset;
        if (off * 20 >= flat_n && off < flat_n) { dp_split_step = int(si); dp_split = off; break; }
    }
    if (loss_scale == 0.f) {
        int e = int(std::floor(std::log2(double(std::max<long long>(V0, 1))))) - 2;
        e = std::max(4, std::min(e, 24));
        loss_scale = std::ldexp(1.f, e);
    }
    planned = true;
    planned_training = tr;
    packs_dirty = true;
    return 0;
}

int Model::repack_on(cudaStream_t on) {
    M_CHECK(pack_all_launch(d_pack_descs, d_pack_first, n_pack_jobs, n_pack_blocks, on));
    ++launches;
    return 0;
}

int Model::repack() {
    if (pack_pending) {   // the re-pack that Model::step started on the side stream (it overlaps the next sample's augmentation)
        M_CUDA(cudaStreamWaitEvent(stream, ev_pack, 0));
        pack_pending = false;
    }
    if (!packs_dirty) return 0;
    M_CHECK(repack_on(stream));
    packs_dirty = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// forward (unet.cpp:168-193)
// ------------------------------------------------------------------------------------------------
int Model::run_forward(int levels_wanted, bool bn_eval) {
    for (auto& s : steps) {
        if (s.kind == Step::CONV) {
            if (s.head_level >= levels_wanted) continue;
            if (s.head_fwd_fused) {
                const Ten& a = tens[s.in0];
                SrcTransform xf{};
                const void* x = a.p;
                if (s.xf_from >= 0) {
                    const Step& n = steps[s.xf_from];
                    xf.enabled = 1;
                    xf.C = tens[n.in0].C;
                    xf.has_norm = n.cur_has_norm;
                    xf.act = n.act;
                    xf.mean = n.cur_mean;
                    xf.rstd = n.cur_rstd;
                    xf.gamma = n.norm ? param_ptr(n.p_g) : nullptr;
                    xf.beta = n.norm ? param_ptr(n.p_g + 1) : nullptr;
                    xf.writeback = training ? a.p : nullptr;
                    x = tens[n.in0].p;
                }
                M_CHECK(head_fwd_launch(x, a.C, a.Cp, param_ptr(s.p_w), param_ptr(s.p_b), logits[s.head_level], s.g.cout, a.V(), stream,
                                        s.xf_from >= 0 ? &xf : nullptr));
                ++launches;
                continue;
            }
            ConvLaunch cfg{};
            cfg.kc = s.fkc;
            cfg.epi = s.head_level >= 0 ? EPI_PLANAR32 : EPI_STORE16;
            int rows = 0;
            cfg.stats_grid_out = &rows;
            if (s.stats) cfg.stats_partials = d_partials;
            cfg.splitk_scratch = d_splitk; cfg.splitk_scratch_bytes = splitk_bytes;
            trace_launch(s, "fwd", conv_kernel_kind(s.fprobs, cfg), int(s.fprobs.size()), s.flops);
            prof_begin(conv_kernel_kind(s.fprobs, cfg), s.flops);
            M_CHECK(conv_launch(s.fprobs, cfg, stream));
            prof_end();
            ++launches;
            if (s.stats) { last_stat_rows = rows; last_stat_ntot = s.fprobs[0].ntile * s.fprobs[0].ntiles; }
        } else if (s.kind == Step::NORMACT) {
            const Ten& a = tens[s.in0];
            const float* gamma = s.norm ? param_ptr(s.p_g) : nullptr;
            const float* beta = s.norm ? param_ptr(s.p_g + 1) : nullptr;
            if (s.norm == 2 && (bn_eval || bn_running)) {
                // eval(): y = gamma*(x - running_mean)/sqrt(running_var) + beta, eps 0 (unet.cpp:80-84; validation, train.cpp:836)
                M_CHECK(rstd_from_var_launch(d_buffers[s.buf0 + 1], s.rstd, a.C, 0.f, stream));
                s.cur_mean = d_buffers[s.buf0]; s.cur_rstd = s.rstd; s.cur_has_norm = 1;
                if (!s.elided) M_CHECK(norm_act_fwd_launch(a.p, tens[s.out].p, a.V(), a.C, a.Cp, 1, s.act, d_buffers[s.buf0], s.rstd, gamma, beta, stream));
                launches += s.elided ? 1 : 2;
            } else if (s.norm == 1 || (s.norm == 2 && training)) {
                int rows = last_stat_rows, ntot = last_stat_ntot;
                if (!s.stats_from_conv) {
                    M_CHECK(channel_stats_launch(a.p, a.V(), a.C, a.Cp, d_partials, &rows, stream));
                    ntot = a.Cp;
                    ++launches;
                }
                float* rm = (s.norm == 2) ? d_buffers[s.buf0] : nullptr;
                float* rv = (s.norm == 2) ? d_buffers[s.buf0 + 1] : nullptr;
                M_CHECK(finalize_stats_launch(d_partials, rows, ntot, a.C, double(a.V()), s.norm == 1 ? 1e-5f : 0.f, s.mean, s.rstd,
                                              rm, rv, 0.1f, stream));
                s.cur_mean = s.mean; s.cur_rstd = s.rstd; s.cur_has_norm = 1;
                if (!s.elided) M_CHECK(norm_act_fwd_launch(a.p, tens[s.out].p, a.V(), a.C, a.Cp, 1, s.act, s.mean, s.rstd, gamma, beta, stream));
                launches += s.elided ? 1 : 2;
            } else {
                // no norm, or BatchNorm after prepare_for_inference (mean 0, var 1, eps 0 => y = gamma*x + beta; unet.cpp:7-22)
                s.cur_mean = nullptr; s.cur_rstd = nullptr; s.cur_has_norm = s.norm ? 1 : 0;
                if (!s.elided) {
                    M_CHECK(norm_act_fwd_launch(a.p, tens[s.out].p, a.V(), a.C, a.Cp, s.norm ? 1 : 0, s.act, nullptr, nullptr, gamma, beta, stream));
                    ++launches;
                }
            }
        } else if (s.kind == Step::MAXPOOL) {
            const Ten& o = tens[s.out];
            M_CHECK(maxpool_fwd_launch(tens[s.in0].p, o.p, s.idx, o.Cp, o.d, o.h, o.w, stream));
            ++launches;
        } else {
            const Ten& a = tens[s.in0];
            M_CHECK(upsample_fwd_launch(a.p, tens[s.out].p, a.Cp, a.d, a.h, a.w, stream));
            ++launches;
        }
    }
    return 0;
}

int Model::upload_input(const float* in, int where) {
    const long long V0 = tens[0].V();
    const float* src = in;
    if (where == 0) {
        M_CUDA(cudaMemcpyAsync(d_in_f32, in, size_t(in_count) * V0 * 4, cudaMemcpyHostToDevice, stream));
        src = d_in_f32;
    }
    M_CHECK(pack_act_launch(src, tens[0].p, in_count, tens[0].Cp, V0, false, stream, split_input ? 1 : 0));
    ++launches;
    return 0;
}

int Model::forward(const float* in, float* const* out_levels, int n_levels_wanted, int where) {
    M_CHECK(ensure_plan());
    M_CHECK(repack());
    const int L = int(output.size());
    if (n_levels_wanted < 1 || n_levels_wanted > L) { set_error("levels wanted out of range"); return 1; }
    const float* fsrc = in;
    if (where == 0) {
        M_CUDA(cudaMemcpyAsync(d_in_f32, in, size_t(in_count) * tens[0].V() * 4, cudaMemcpyHostToDevice, stream));
        fsrc = d_in_f32;
    }
    M_CHECK(graph_run({2u, uint64_t(reinterpret_cast<uintptr_t>(fsrc)), uint64_t(n_levels_wanted), uint64_t(bn_running ? 1 : 0)}, [&]() -> int {
        M_CHECK(upload_input(fsrc, 1));
        return run_forward(n_levels_wanted);
    }));
    for (int l = 0; l < n_levels_wanted; ++l) {
        if (!out_levels || !out_levels[l]) continue;
        if (!logits[l]) { set_error("undefined deep supervision output at level " + std::to_string(l)); return 1; }
        const long long v = 1LL * level_dims[3 * l] * level_dims[3 * l + 1] * level_dims[3 * l + 2];
        M_CUDA(cudaMemcpyAsync(out_levels[l], logits[l], size_t(out_count) * v * 4,
                               where == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, stream));
    }
    if (where == 0) return sync();
    return 0;
}

int Model::evaluate_windows(const float* const* in_windows, float* const* out_windows, int n_windows, int where) {
    if (n_windows <= 0) return 0;
    if (where != 0 || n_windows == 1) {
        for (int i = 0; i < n_windows; ++i) {
            float* outs[1] = {out_windows[i]};
            M_CHECK(forward(in_windows[i], outs, 1, where));
        }
        return 0;
    }
    M_CHECK(ensure_plan());
    M_CHECK(repack());
    if (!logits[0]) { set_error("undefined output at level 0"); return 1; }
    const long long V0 = tens[0].V();
    const size_t in_bytes = size_t(in_count) * V0 * 4, out_bytes = size_t(out_count) * V0 * 4;
    if (ew_in_bytes < in_bytes || ew_out_bytes < out_bytes) {
        M_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < 2; ++i) {
            if (ew_in[i]) cudaFree(ew_in[i]);
            if (ew_out[i]) cudaFree(ew_out[i]);
            ew_in[i] = ew_out[i] = nullptr;
            M_CUDA(cudaMalloc(reinterpret_cast<void**>(&ew_in[i]), in_bytes));
            M_CUDA(cudaMalloc(reinterpret_cast<void**>(&ew_out[i]), out_bytes));
        }
        ew_in_bytes = in_bytes;
        ew_out_bytes = out_bytes;
    }
    if (!stream3) M_CUDA(cudaStreamCreateWithFlags(&stream3, cudaStreamNonBlocking));
    cudaStream_t up = stream3, down = stream2;
    cudaEvent_t ev_up[2], ev_packed[2], ev_out[2], ev_down[2];
    for (int i = 0; i < 2; ++i) {
        M_CUDA(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
        M_CUDA(cudaEventCreateWithFlags(&ev_packed[i], cudaEventDisableTiming));
        M_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        M_CUDA(cudaEventCreateWithFlags(&ev_down[i], cudaEventDisableTiming));
    }
    int rc = 0;
    auto upload = [&](int i) {
        const int s = i & 1;
        if (i >= 2) cudaStreamWaitEvent(up, ev_packed[s], 0);   // the slot's previous window has been packed into the NDHWC input
        cudaMemcpyAsync(ew_in[s], in_windows[i], in_bytes, cudaMemcpyHostToDevice, up);
        cudaEventRecord(ev_up[s], up);
    };
    upload(0);
    for (int i = 0; i < n_windows && !rc; ++i) {
        const int s = i & 1;
        if (i + 1 < n_windows) upload(i + 1);
        cudaStreamWaitEvent(stream, ev_up[s], 0);
        // packing the slot into the NDHWC input + the forward: one captured graph per staging slot
        rc = graph_run({3u, uint64_t(reinterpret_cast<uintptr_t>(ew_in[s])), uint64_t(bn_running ? 1 : 0)}, [&]() -> int {
            M_CHECK(pack_act_launch(ew_in[s], tens[0].p, in_count, tens[0].Cp, V0, false, stream, split_input ? 1 : 0));
            ++launches;
            return run_forward(1);
        });
        cudaEventRecord(ev_packed[s], stream);
        if (rc) break;
        if (i >= 2) cudaStreamWaitEvent(stream, ev_down[s], 0);  // the slot's previous window has left for the host
        cudaMemcpyAsync(ew_out[s], logits[0], out_bytes, cudaMemcpyDeviceToDevice, stream);
        cudaEventRecord(ev_out[s], stream);
        cudaStreamWaitEvent(down, ev_out[s], 0);
        cudaMemcpyAsync(out_windows[i], ew_out[s], out_bytes, cudaMemcpyDeviceToHost, down);
        cudaEventRecord(ev_down[s], down);
    }
    cudaStreamSynchronize(up);
    cudaStreamSynchronize(down);
    if (!rc) rc = sync();
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_up[i]); cudaEventDestroy(ev_packed[i]); cudaEventDestroy(ev_out[i]); cudaEventDestroy(ev_down[i]); }
    if (!rc && cudaGetLastError() != cudaSuccess) { set_error("evaluate_windows: copy failed"); rc = 1; }
    return rc;
}

// evaluate.cpp:195-230 + 274 for one volume: cut windows (handle_fov_pre), forward()[0] per window, softmax, re-assemble
// (handle_fov_post), create_mask, argmax.  The TIPL side of this is un-vendored: see postproc.cu for the assumed semantics.
int Model::evaluate_volume(const float* volume, int vw, int vh, int vd, int sx, int sy, int sz, float threshold, uint8_t* label_out,
                           float* fg_out, float* prob_out, int where, int* n_windows_out) {
    if (vw < 1 || vh < 1 || vd < 1 || !volume || !label_out) { set_error("evaluate_volume: invalid argument"); return 1; }
    M_CHECK(ensure_plan());
    M_CHECK(repack());
    if (!logits[0]) { set_error("undefined output at level 0"); return 1; }
    const int ww = dim[0], wh = dim[1], wd = dim[2], C = out_count;
    const long long VV = 1LL * vw * vh * vd, WV = 1LL * ww * wh * wd;
    // device layout: [volume in_count*VV] [window in_count*WV] [acc C*VV] [cnt VV] [fg VV] [labels VV bytes]
    const size_t need = (size_t(in_count) * VV + size_t(in_count) * WV + size_t(C) * VV + 2 * size_t(VV)) * 4 + size_t(VV) + 256;
    if (ev_bytes < need) {
        M_CUDA(cudaStreamSynchronize(stream));
        if (ev_buf) cudaFree(ev_buf);
        ev_buf = nullptr; ev_bytes = 0;
        M_CUDA(cudaMalloc(reinterpret_cast<void**>(&ev_buf), need));
        ev_bytes = need;
    }
    float* d_vol = ev_buf;
    float* d_win = d_vol + size_t(in_count) * VV;
    float* d_acc = d_win + size_t(in_count) * WV;
    float* d_cnt = d_acc + size_t(C) * VV;
    float* d_fg = d_cnt + VV;
    uint8_t* d_lab = reinterpret_cast<uint8_t*>(d_fg + VV);
    const float* vol = volume;
    const std::vector<int> ox = window_origins(vw, ww, sx), oy = window_origins(vh, wh, sy), oz = window_origins(vd, wd, sz);
    // Host volume: uploaded in z slabs on the copy stream, one slab per window z-origin (the slab ends where that origin's windows end),
    // so the first windows run while the rest of the volume is still on its way (the 131 MB of a 320^3 volume are 5 ms of PCIe time).
    std::vector<int> slab_end;        // exclusive z end of slab k (ascending); windows at oz[k] read z < slab_end[k]
    if (where == 0) {
        if (!stream3) M_CUDA(cudaStreamCreateWithFlags(&stream3, cudaStreamNonBlocking));
        while (ev_slabs.size() < oz.size()) {
            cudaEvent_t e;
            M_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_slabs.push_back(e);
        }
        M_CUDA(cudaEventRecord(ev_fork, stream));             // the previous call's kernels have finished with the volume buffer
        M_CUDA(cudaStreamWaitEvent(stream3, ev_fork, 0));
        int z0 = 0;
        for (size_t k = 0; k < oz.size(); ++k) {
            const int z1 = k + 1 == oz.size() ? vd : std::min(vd, std::max(z0, oz[k] + wd));
            if (z1 > z0)
                M_CUDA(cudaMemcpy2DAsync(d_vol + size_t(z0) * vh * vw, size_t(VV) * 4, volume + size_t(z0) * vh * vw, size_t(VV) * 4,
                                         size_t(z1 - z0) * vh * vw * 4, size_t(in_count), cudaMemcpyHostToDevice, stream3));
            M_CUDA(cudaEventRecord(ev_slabs[k], stream3));
            slab_end.push_back(z1);
            z0 = z1;
        }
        vol = d_vol;
    }
    M_CUDA(cudaMemsetAsync(d_acc, 0, (size_t(C) + 1) * VV * 4, stream));
    int n = 0;
    for (size_t zi = 0; zi < oz.size(); ++zi) {
        const int z = oz[zi];
        if (where == 0) M_CUDA(cudaStreamWaitEvent(stream, ev_slabs[zi], 0));
        for (int y : oy)
            for (int x : ox) {
                M_CHECK(crop_window_launch(vol, d_win, in_count, vw, vh, vd, ww, wh, wd, x, y, z, stream));
                // pack + forward of the window buffer: the same launches for every window, replayed as a graph from the second one on
                M_CHECK(graph_run({4u, uint64_t(reinterpret_cast<uintptr_t>(d_win)), uint64_t(bn_running ? 1 : 0)}, [&]() -> int {
                    M_CHECK(pack_act_launch(d_win, tens[0].p, in_count, tens[0].Cp, WV, false, stream, split_input ? 1 : 0));
                    return run_forward(1);
                }));
                M_CHECK(softmax_accumulate_launch(logits[0], d_acc, d_cnt, C, vw, vh, vd, ww, wh, wd, x, y, z, stream));
                launches += 3;
                ++n;
            }
    }
    M_CHECK(mask_argmax_launch(d_acc, d_cnt, d_lab, d_fg, C, VV, threshold, prob_out != nullptr, stream));
    ++launches;
    if (n_windows_out) *n_windows_out = n;
    const cudaMemcpyKind k = where == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    M_CUDA(cudaMemcpyAsync(label_out, d_lab, size_t(VV), k, stream));
    if (fg_out) M_CUDA(cudaMemcpyAsync(fg_out, d_fg, size_t(VV) * 4, k, stream));
    if (prob_out) M_CUDA(cudaMemcpyAsync(prob_out, d_acc, size_t(C) * VV * 4, k, stream));
    if (where == 0) return sync();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// backward (autograd of the step body, train.cpp:706)
// ------------------------------------------------------------------------------------------------
int Model::run_backward() {
    static const bool no_side = std::getenv("U3D_ONE_STREAM") != nullptr;
    const bool two_streams = !no_side && !prof_on && stream2 != nullptr;   // the per-family event profile needs serial kernels
    bool dp_tail_launched = false;
    const bool dp_overlap_now = dp_comm != nullptr && stream4 != nullptr && dp_split_step >= 0 && dp_seen == dp_microbatches && !no_side;
    std::fill(grad_written.begin(), grad_written.end(), 0);
    // fused heads: the loss-gradient kernel (launched before this function) already stored dL/dx of the head input
    for (const Step& s : steps)
        if (s.kind == Step::CONV && s.head_bwd_fused) grad_written[s.in0] = 1;
    // data gradient of conv step s wrt its source src (store, or read-add-store when the tensor already holds a contribution)
    auto launch_dgrad = [&](Step& s, int src) -> int {
        const int t = src ? s.in1 : s.in0;
        ConvLaunch cfg{};
        cfg.kc = s.dg[src].kc;
        cfg.epi = grad_written[t] ? EPI_ACCUM16 : EPI_STORE16;
        cfg.splitk_scratch = d_splitk; cfg.splitk_scratch_bytes = splitk_bytes;
        const double fl = s.flops * double(s.g.cin[src]) / double(s.g.cin[0] + s.g.cin[1]);
        trace_launch(s, src ? "dgrad1" : "dgrad0", conv_kernel_kind(s.dg[src].probs, cfg), int(s.dg[src].probs.size()), fl);
        prof_begin(conv_kernel_kind(s.dg[src].probs, cfg), fl);
        M_CHECK(conv_launch(s.dg[src].probs, cfg, stream));
        prof_end();
        ++launches;
        grad_written[t] = 1;
        return 0;
    };
    // A skip tensor receives two data gradients: from the decoder conv that consumes it (early in this loop) and from the next encoder
    // level's stride-2 conv (late).  The stride-2 data gradient scatters 32-byte rows (parity-stacked epilogue) and is bound by that
    // traffic; as a read-add-store it moves it twice.  So the decoder conv's skip gradient is DEFERRED until just before the skip
    // tensor's producer runs its backward: the scatter kernel stores, the dense kernel (MMA-bound, the extra read is free) accumulates.
    static const bool no_defer = std::getenv("U3D_NO_DEFER_SKIP") != nullptr;
    std::vector<int> deferred;   // conv steps whose source-0 data gradient is pending
    for (int si = int(steps.size()) - 1; si >= 0; --si) {
        Step& s = steps[si];
        if (s.out >= 0 && !deferred.empty()) {
            for (size_t k = 0; k < deferred.size();) {
                Step& c = steps[size_t(deferred[k])];
                if (c.in0 == s.out) {
                    M_CHECK(launch_dgrad(c, 0));
                    deferred.erase(deferred.begin() + long(k));
                } else
                    ++k;
            }
        }
        if (s.kind == Step::CONV && s.head_bwd_fused) continue;
        if (s.kind == Step::CONV) {
            const bool head = s.head_level >= 0;
            if (!head && !grad_written[s.out]) continue;  // output never used downstream
            const void* dy = head ? dlogits[s.head_level] : tens[s.out].grad;
            const long long Vout = 1LL * s.g.out_d * s.g.out_h * s.g.out_w;
            if (!s.stats) {
                // bias gradient = sum_v dy (a conv feeding a norm has an exactly-zero bias gradient: the norm removes the mean)
                M_CHECK(colsum_accumulate_launch(dy, Vout, s.g.cout, pad16(s.g.cout), d_partials, grad_ptr(s.p_b), stream));
                launches += 2;
            }
            // weight gradient on the side stream: it only reads x and dy (both final here) and adds into this layer's slice of the
            // flat gradient, so it can overlap the data gradient below (the small deep-level launches fill the SMs the other leaves idle)
            WgradLaunch wc{};
            wc.partial_scratch = d_wgrad_partial; wc.partial_scratch_bytes = wgrad_partial_bytes;
            bool all_rows = !s.wg.empty(), all_quad = !s.wg.empty();
            for (const auto& wp : s.wg) {
                all_quad = all_quad && conv_wgrad_quad_eligible(wp);
                all_rows = all_rows && (conv_wgrad_band_eligible(wp) || conv_wgrad_quad_eligible(wp));
            }
            cudaStream_t ws = two_streams ? stream2 : stream;
            if (two_streams) {
                M_CUDA(cudaEventRecord(ev_fork, stream));
                M_CUDA(cudaStreamWaitEvent(stream2, ev_fork, 0));
            }
            trace_launch(s, "wgrad", all_quad ? 6 : all_rows ? 3 : 1, int(s.wg.size()), s.flops);
            prof_begin(all_quad ? 6 : all_rows ? 3 : 1, s.flops, ws);
            int nl = 0;
            M_CHECK(conv_wgrad_dispatch(s.wg, wc, ws, &nl));
            prof_end(ws);
            launches += nl;
            const int ins[2] = {s.in0, s.in1};
            for (int src = 0; src < 2; ++src) {
                if (ins[src] < 0 || !tens[ins[src]].needs_grad) continue;
                if (src == 0 && s.in1 >= 0 && !no_defer && !grad_written[s.in0] && skip_has_other_consumer[size_t(si)]) {
                    deferred.push_back(si);
                    continue;
                }
                M_CHECK(launch_dgrad(s, src));
            }
            if (dp_overlap_now && si == dp_split_step) {
                // every contribution to the tail bucket has been issued: weight gradients on stream2, bias / norm / head gradients on
                // the main stream.  Reduce it on stream4 while the (parameter-poor, time-rich) first encoder levels run their backward.
                M_CUDA(cudaEventRecord(ev_ar_ready, stream));
                M_CUDA(cudaStreamWaitEvent(stream4, ev_ar_ready, 0));
                if (two_streams) {
                    M_CUDA(cudaEventRecord(ev_join, stream2));
                    M_CUDA(cudaStreamWaitEvent(stream4, ev_join, 0));
                }
                const ncclResult_t nr = ncclAllReduce(d_grads + dp_split, d_grads + dp_split, size_t(flat_n - dp_split), ncclFloat, ncclSum,
                                                      static_cast<ncclComm_t>(dp_comm), stream4);
                if (nr != ncclSuccess) { set_error(std::string("ncclAllReduce: ") + ncclGetErrorString(nr)); return 1; }
                dp_tail_launched = true;
            }
        } else {
            if (!grad_written[s.out]) continue;
            const Ten& a = tens[s.in0];
            const Ten& o = tens[s.out];
            if (!a.needs_grad) continue;
            void* target = grad_written[s.in0] ? d_scratch : a.grad;
            if (s.kind == Step::NORMACT) {
                const float* gamma = s.norm ? param_ptr(s.p_g) : nullptr;
                const float* beta = s.norm ? param_ptr(s.p_g + 1) : nullptr;
                M_CHECK(norm_act_bwd_launch(a.p, o.grad, target, a.V(), a.C, a.Cp, s.norm ? 1 : 0, s.act, s.mean, s.rstd, gamma, beta,
                                            d_partials, d_sums, s.norm ? grad_ptr(s.p_g) : nullptr,
                                            s.norm ? grad_ptr(s.p_g + 1) : nullptr, stream, d_counter));
                launches += s.norm ? 2 : 1;
            } else if (s.kind == Step::MAXPOOL) {
                M_CHECK(maxpool_bwd_launch(o.grad, s.idx, target, o.Cp, o.d, o.h, o.w, stream));
                ++launches;
            } else {
                M_CHECK(upsample_bwd_launch(o.grad, target, a.Cp, a.d, a.h, a.w, stream));
                ++launches;
            }
            if (grad_written[s.in0]) {
                M_CHECK(add16_launch(a.grad, d_scratch, a.V() * (a.Cp / 8), stream));
                ++launches;
            }
            grad_written[s.in0] = 1;
        }
    }
    for (int k : deferred) M_CHECK(launch_dgrad(steps[size_t(k)], 0));   // (a producer that was never reached)
    if (dp_tail_launched) {
        // the tail bucket's all-reduce joins the main stream at the end of the backward pass (it ran beside the first encoder levels'
        // backward).  Joining here rather than in unet3d_step keeps the whole micro-batch one fork/join region: capturable as a graph.
        M_CUDA(cudaEventRecord(ev_ar_done, stream4));
        M_CUDA(cudaStreamWaitEvent(stream, ev_ar_done, 0));
    }
    if (two_streams) {   // the update (and the next forward) must see every weight gradient
        M_CUDA(cudaEventRecord(ev_join, stream2));
        M_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    }
    return 0;
}

// one N=1 micro-batch: forward, 5-level deep supervision, backward (train.cpp:628-706)
int Model::train_microbatch(const float* in, const float* label, int collapse_before, int use_ce, int use_dice, int use_mse,
                            float* loss_out3, float* all_levels, int where) {
    if (!training) { set_error("train_microbatch needs training mode (unet3d_set_mode(h, 1))"); return 1; }
    M_CHECK(resolve_status());   // the loss scale may change with the last update's overflow flag
    M_CHECK(ensure_plan());
    M_CHECK(repack());
    const int L = int(output.size());
    for (int l = 0; l < L; ++l) {
        if (!logits[l]) { set_error("undefined deep supervision output at level " + std::to_string(l)); return 1; }
        if (level_dims[3 * l] != (dim[2] >> l) || level_dims[3 * l + 1] != (dim[1] >> l) || level_dims[3 * l + 2] != (dim[0] >> l) ||
            ((dim[2] >> l) << l) != dim[2] || ((dim[1] >> l) << l) != dim[1] || ((dim[0] >> l) << l) != dim[0]) {
            set_error("deep supervision needs level k to be exactly 1/2^k of the grid (train.cpp:645-662)");
            return 1;
        }
    }
    if (collapse_before < 0 || collapse_before >= std::max(out_count, 1)) { set_error("invalid collapse_before"); return 1; }
    const long long V0 = tens[0].V();
    const float* src = in;
    const float* lab = label;
    if (where == 0) {
        M_CUDA(cudaMemcpyAsync(d_in_f32, in, size_t(in_count) * V0 * 4, cudaMemcpyHostToDevice, stream));
        M_CUDA(cudaMemcpyAsync(d_label, label, size_t(V0) * 4, cudaMemcpyHostToDevice, stream));
        src = d_in_f32;
        lab = d_label;
    }
    ++dp_seen;
    if (dp_comm != nullptr && dp_seen > dp_microbatches) {
        set_error("more micro-batches in this step than declared in unet3d_attach_comm (the tail gradient bucket is already reduced)");
        return 1;
    }
    auto enqueue = [&]() -> int {
        M_CHECK(upload_input(src, 1));
        M_CHECK(run_forward(L));
        float weight_sum = 0.f;
        for (int k = 0; k < L; ++k) weight_sum += 1.0f / float(1 << k);
        const float inv_weight_sum = 1.0f / weight_sum;
        const bool any = use_ce || use_dice || use_mse;
        for (int k = 0; k < L; ++k) {
            LossLevel Q{};
            Q.logits = logits[k]; Q.label = lab; Q.dlogits = dlogits[k];
            Q.C = out_count; Q.Cp = pad16(out_count); Q.collapse_before = collapse_before;
            Q.d = level_dims[3 * k]; Q.h = level_dims[3 * k + 1]; Q.w = level_dims[3 * k + 2];
            Q.H0 = dim[1]; Q.W0 = dim[0]; Q.shift = k;
            const float nw = (1.0f / float(1 << k)) * inv_weight_sum;
            Q.w_ce = (use_ce || !any) ? nw : 0.f;   // "if(!level_loss.defined()) level_loss = ce" (train.cpp:696-697)
            Q.w_dice = use_dice ? nw : 0.f;
            Q.w_mse = use_mse ? nw : 0.f;
            Q.loss_scale = loss_scale;
            Q.acc = d_loss_acc + 80 * k;
            Q.part = d_loss_part + size_t(k) * loss_part_rows() * loss_part_cols();
            Q.out3 = d_losses + 3 * k;
            HeadFuse Hd{};
            const Step* hs = head_step[k] >= 0 ? &steps[head_step[k]] : nullptr;
            if (hs && hs->head_bwd_fused) {
                const Ten& a = tens[hs->in0];
                Hd.x = a.p; Hd.xc = a.C; Hd.xcp = a.Cp;
                Hd.w = param_ptr(hs->p_w);
                Hd.dx = a.grad; Hd.dx_accum = 0;
                Hd.dw = grad_ptr(hs->p_w); Hd.db = grad_ptr(hs->p_b);
                M_CHECK(loss_level_launch(Q, &Hd, stream));
            } else
                M_CHECK(loss_level_launch(Q, stream));
            launches += 4;
        }
        return run_backward();
    };
    // with an attached communicator the LAST micro-batch of a step carries the tail bucket's all-reduce on a forked stream (NCCL
    // collectives are capturable; every rank captures and replays the same sequence): a different graph than the other micro-batches
    static const bool no_side_dp = std::getenv("U3D_ONE_STREAM") != nullptr;
    static const bool no_nccl_graph = std::getenv("U3D_NO_NCCL_GRAPH") != nullptr;
    const bool overlap_now = dp_comm != nullptr && stream4 != nullptr && dp_split_step >= 0 && dp_seen == dp_microbatches && !no_side_dp;
    if (dp_comm == nullptr || !no_nccl_graph) {
        uint32_t ls_bits;
        std::memcpy(&ls_bits, &loss_scale, 4);
        const std::vector<uint64_t> key = {1u, uint64_t(reinterpret_cast<uintptr_t>(src)), uint64_t(reinterpret_cast<uintptr_t>(lab)),
                                           uint64_t(collapse_before), uint64_t((use_ce ? 1 : 0) | (use_dice ? 2 : 0) | (use_mse ? 4 : 0)), ls_bits,
                                           uint64_t(overlap_now ? 1 : 0), uint64_t(reinterpret_cast<uintptr_t>(dp_comm))};
        M_CHECK(graph_run(key, enqueue));
    } else
        M_CHECK(enqueue());
    if (overlap_now) dp_tail_reduced = true;   // (host state: also on a graph replay)
    std::vector<float> h(size_t(3) * L);
    M_CUDA(cudaMemcpyAsync(h.data(), d_losses, h.size() * 4, cudaMemcpyDeviceToHost, stream));
    M_CHECK(sync());
    if (loss_out3) std::memcpy(loss_out3, h.data(), 12);
    if (all_levels) std::memcpy(all_levels, h.data(), h.size() * 4);
    return 0;
}

int Model::prefetch_slot(float** in_dev, float** label_dev, int* slot) {
    cudaSetDevice(device);
    const size_t V = size_t(dim[0]) * dim[1] * dim[2];
    const size_t need = (size_t(in_count) + 1) * V * 4;
    if (!stream3) {
        M_CUDA(cudaStreamCreateWithFlags(&stream3, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) M_CUDA(cudaEventCreateWithFlags(&ev_sample[i], cudaEventDisableTiming));
    }
    if (pf_bytes < need) {
        M_CUDA(cudaStreamSynchronize(stream3));
        M_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < 2; ++i) {
            if (pf_in[i]) cudaFree(pf_in[i]);
            pf_in[i] = nullptr;
            M_CUDA(cudaMalloc(reinterpret_cast<void**>(&pf_in[i]), need));
            pf_label[i] = pf_in[i] + size_t(in_count) * V;
        }
        pf_bytes = need;
        pf_pending[0] = pf_pending[1] = false;
        pf_next = pf_head = 0;
    }
    if (pf_pending[pf_next]) { set_error("prefetch: both staging slots hold samples that have not been consumed"); return 1; }
    *slot = pf_next;
    *in_dev = pf_in[pf_next];
    *label_dev = pf_label[pf_next];
    return 0;
}

int Model::consume_prefetched(float** in_dev, float** label_dev) {
    if (!pf_pending[pf_head]) { set_error("no prefetched sample (call unet3d_prefetch_augmented first)"); return 1; }
    M_CUDA(cudaStreamWaitEvent(stream, ev_sample[pf_head], 0));
    *in_dev = pf_in[pf_head];
    *label_dev = pf_label[pf_head];
    pf_pending[pf_head] = false;
    pf_head ^= 1;
    return 0;
}

int Model::staging(float** in_dev, float** label_dev) {
    M_CHECK(ensure_plan());
    *in_dev = d_in_f32;
    *label_dev = d_label;
    return 0;
}

// validation forward + level-0 losses (train.cpp:826-851)
int Model::validate(const float* in, const float* label, int collapse_before, float* loss_out3, int where) {
    M_CHECK(ensure_plan());
    M_CHECK(repack());
    const long long V0 = tens[0].V();
    M_CHECK(upload_input(in, where));
    const float* lab = label;
    if (where == 0) {
        M_CUDA(cudaMemcpyAsync(d_label, label, size_t(V0) * 4, cudaMemcpyHostToDevice, stream));
        lab = d_label;
    }
    M_CHECK(run_forward(1, true));   // output_model->eval() (train.cpp:836): BatchNorm3d reads its running statistics, nothing is updated
    LossLevel Q{};
    Q.logits = logits[0]; Q.label = lab; Q.dlogits = nullptr;
    Q.C = out_count; Q.Cp = pad16(out_count); Q.collapse_before = collapse_before;
    Q.d = level_dims[0]; Q.h = level_dims[1]; Q.w = level_dims[2];
    Q.H0 = dim[1]; Q.W0 = dim[0]; Q.shift = 0;
    Q.acc = d_loss_acc; Q.part = d_loss_part; Q.out3 = d_losses;
    M_CHECK(loss_level_launch(Q, stream));
    launches += 3;
    if (!h_val) M_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_val), 16));
    M_CUDA(cudaMemcpyAsync(h_val, d_losses, 12, cudaMemcpyDeviceToHost, stream));
    val_pending = true;
    if (loss_out3 == nullptr) return 0;   // asynchronous form: the caller collects the losses with validate_result()
    return validate_result(loss_out3);
}

// second half of the asynchronous validation: waits for this handle's stream only (a training handle on the same GPU keeps running)
int Model::validate_result(float* loss_out3) {
    if (!val_pending) { set_error("validate_result without a pending unet3d_validate_async"); return 1; }
    cudaSetDevice(device);
    M_CHECK(sync());
    val_pending = false;
    if (loss_out3) std::memcpy(loss_out3, h_val, 12);
    return 0;
}

int Model::copy_from(const Model& src) {
    // unet.cpp:195-222: parameters (and buffers) with identical sizes are copied, others are left alone
    cudaSetDevice(device);
    const size_t n = std::min(params.size(), src.params.size());
    cudaStreamSynchronize(src.stream);
    for (size_t i = 0; i < n; ++i) {
        if (params[i].shape != src.params[i].shape) continue;
        M_CUDA(cudaMemcpyPeerAsync(d_params + params[i].offset, device, src.d_params + src.params[i].offset, src.device,
                                   size_t(params[i].numel) * 4, stream));
    }
    for (size_t i = 0; i < d_buffers.size() && i < src.d_buffers.size(); ++i) {
        if (buffer_len[i] != src.buffer_len[i]) continue;
        M_CUDA(cudaMemcpyPeerAsync(d_buffers[i], device, src.d_buffers[i], src.device, size_t(buffer_len[i]) * 4, stream));
    }
    for (int k = 0; k < 3; ++k) { voxel_size[k] = src.voxel_size[k]; }
    fov_strategy = src.fov_strategy; postproc = src.postproc; preproc = src.preproc;   // unet.cpp:217-221
    M_CHECK(set_dim(src.dim[0], src.dim[1], src.dim[2]));
    M_CUDA(cudaStreamSynchronize(stream));
    packs_dirty = true;
    return 0;
}

}  // namespace u3d
