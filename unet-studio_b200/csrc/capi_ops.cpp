// Operator-level C-ABI entry points (include/unet3d_b200.h, "u3d_op_*"): one reference layer through
// the production planner + tcgen05 kernels with HOST fp32 NCDHW buffers.  Used by the parity tests to
// localise errors per layer; the model-level API (capi.cpp) drives the same planner and kernels.
#include <cstring>
#include <vector>

#include "../../include/unet3d_b200.h"
#include "elementwise.h"
#include "plan.h"

namespace u3d {
namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16) == cudaSuccess ? 0 : 1; }
};

#define OP_CHECK(x)                                                         \
    do {                                                                    \
        if ((x) != 0) return 1;                                             \
    } while (0)
#define OP_CUDA(x)                                                          \
    do {                                                                    \
        cudaError_t e_ = (x);                                               \
        if (e_ != cudaSuccess) {                                            \
            set_error(std::string(#x) + ": " + cudaGetErrorString(e_));     \
            return 1;                                                       \
        }                                                                   \
    } while (0)

int finish(cudaStream_t s) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        unsigned int code = 0;
        set_error(std::string("kernel failed: ") + cudaGetErrorString(e));
        (void)code;
        return 1;
    }
    const unsigned int code = read_device_error();
    if (code) {
        set_error("device pipeline timeout code " + std::to_string(code));
        return 1;
    }
    return 0;
}

LayerGeom make_geom(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d) {
    LayerGeom g{};
    g.transposed = transposed; g.ks = ks; g.stride = stride;
    g.cin[0] = cin0; g.cin[1] = cin1; g.cout = cout;
    g.in_d = d; g.in_h = h; g.in_w = w;
    if (transposed) { g.out_d = 2 * d; g.out_h = 2 * h; g.out_w = 2 * w; }
    else {
        const int pad = (ks - 1) / 2;
        g.out_d = (d + 2 * pad - ks) / stride + 1;
        g.out_h = (h + 2 * pad - ks) / stride + 1;
        g.out_w = (w + 2 * pad - ks) / stride + 1;
    }
    return g;
}

int upload_act(DevBuf& dst, const float* host, int C, long long V, bool bf16, cudaStream_t s) {
    DevBuf tmp;
    if (tmp.alloc(size_t(C) * V * 4) || dst.alloc(size_t(pad16(C)) * V * 2)) { set_error("cudaMalloc failed"); return 1; }
    OP_CUDA(cudaMemcpyAsync(tmp.p, host, size_t(C) * V * 4, cudaMemcpyHostToDevice, s));
    OP_CHECK(pack_act_launch(static_cast<const float*>(tmp.p), dst.p, C, pad16(C), V, bf16, s));
    OP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int download_act(const void* dev, float* host, int C, long long V, bool bf16, cudaStream_t s) {
    DevBuf tmp;
    if (tmp.alloc(size_t(C) * V * 4)) { set_error("cudaMalloc failed"); return 1; }
    OP_CHECK(unpack_act_launch(dev, static_cast<float*>(tmp.p), C, pad16(C), V, bf16, s));
    OP_CUDA(cudaMemcpyAsync(host, tmp.p, size_t(C) * V * 4, cudaMemcpyDeviceToHost, s));
    OP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace
}  // namespace u3d

using namespace u3d;

extern "C" int u3d_op_conv_forward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                                   const float* x0, const float* x1, const float* weight, const float* bias, float* y,
                                   double* stats_sum_sumsq, int planar_fp32) {
    cudaStream_t s = 0;
    const LayerGeom g = make_geom(transposed, ks, stride, cin0, cin1, cout, w, h, d);
    const long long Vin = 1LL * w * h * d, Vout = 1LL * g.out_w * g.out_h * g.out_d;
    std::vector<ConvProblem> probs;
    std::vector<PackDesc> packs;
    int kc = 0;
    plan_forward(g, probs, packs, kc, (!planar_fp32 && (conv_band_wants_kc16(ks, stride, transposed, pad16(cin0) + (cin1 ? pad16(cin1) : 0), pad16(cout), Vout) ||
                                                        conv_s2_wants_kc16(ks, stride, transposed, pad16(cin0), cin1 ? 2 : 1, pad16(cout), Vout))) ? 16 : 0);
    DevBuf dx0, dx1, dw, db, dy, dstats, dyf;
    OP_CHECK(upload_act(dx0, x0, cin0, Vin, false, s));
    if (cin1) OP_CHECK(upload_act(dx1, x1, cin1, Vin, false, s));
    const size_t wcount = size_t(cout) * (cin0 + cin1) * (transposed ? 8 : ks * ks * ks);
    if (dw.alloc(wcount * 4) || db.alloc(size_t(cout) * 4) || dy.alloc(size_t(pad16(cout)) * Vout * 2) ||
        dyf.alloc(size_t(cout) * Vout * 4)) { set_error("cudaMalloc failed"); return 1; }
    OP_CUDA(cudaMemcpy(dw.p, weight, wcount * 4, cudaMemcpyHostToDevice));
    if (bias) OP_CUDA(cudaMemcpy(db.p, bias, size_t(cout) * 4, cudaMemcpyHostToDevice));
    OP_CUDA(cudaMemset(dy.p, 0xff, size_t(pad16(cout)) * Vout * 2));  // poison: every voxel must be written
    std::vector<DevBuf> wp(packs.size());
    for (size_t i = 0; i < packs.size(); ++i) {
        if (wp[i].alloc(pack_bytes(packs[i]))) { set_error("cudaMalloc failed"); return 1; }
        packs[i].w = static_cast<const float*>(dw.p);
        packs[i].out = wp[i].p;
        packs[i].out_bf16 = 0;
        OP_CHECK(pack_weights_launch(packs[i], s));
        probs[i].src0 = dx0.p; probs[i].src1 = dx1.p;
        probs[i].dst = planar_fp32 ? dyf.p : dy.p;
        probs[i].wpack = wp[i].p;
        probs[i].bias = bias ? static_cast<const float*>(db.p) : nullptr;
    }
    ConvLaunch cfg{};
    cfg.kc = kc; cfg.a_bf16 = 0; cfg.b_bf16 = 0; cfg.out_bf16 = 0;
    cfg.epi = planar_fp32 ? EPI_PLANAR32 : EPI_STORE16;
    int grid = 0;
    cfg.stats_grid_out = &grid;
    DevBuf dsplit;   // scratch for the deterministic split-K of the small deep-level shapes
    if (Vout <= 16384 && !dsplit.alloc(size_t(48) << 20)) { cfg.splitk_scratch = static_cast<float*>(dsplit.p); cfg.splitk_scratch_bytes = size_t(48) << 20; }
    const int ntot = probs[0].ntile * probs[0].ntiles;
    if (stats_sum_sumsq && (probs.size() == 1 || probs[0].band_pass) && !planar_fp32) {
        if (dstats.alloc(size_t(device_sm_count()) * 2 * ntot * 4)) { set_error("cudaMalloc failed"); return 1; }
        cfg.stats_partials = static_cast<float*>(dstats.p);
    }
    OP_CHECK(conv_launch(probs, cfg, s));
    OP_CHECK(finish(s));
    if (planar_fp32)
        OP_CUDA(cudaMemcpy(y, dyf.p, size_t(cout) * Vout * 4, cudaMemcpyDeviceToHost));
    else
        OP_CHECK(download_act(dy.p, y, cout, Vout, false, s));
    if (cfg.stats_partials) {
        std::vector<float> part(size_t(grid) * 2 * ntot);
        OP_CUDA(cudaMemcpy(part.data(), dstats.p, part.size() * 4, cudaMemcpyDeviceToHost));
        for (int c = 0; c < cout; ++c) {
            double a = 0, q = 0;
            for (int b = 0; b < grid; ++b) {
                a += part[size_t(b) * 2 * ntot + c];
                q += part[size_t(b) * 2 * ntot + ntot + c];
            }
            stats_sum_sumsq[c] = a;
            stats_sum_sumsq[cout + c] = q;
        }
    }
    return 0;
}

extern "C" int u3d_op_conv_backward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                                    const float* x0, const float* x1, const float* weight, const float* dy,
                                    float* gx0, float* gx1, float* gw, int flags) {
    const int accumulate_gx0 = flags & 1;
    cudaStream_t s = 0;
    const LayerGeom g = make_geom(transposed, ks, stride, cin0, cin1, cout, w, h, d);
    const long long Vin = 1LL * w * h * d, Vout = 1LL * g.out_w * g.out_h * g.out_d;
    DevBuf dx[2], ddy, dw, dgw, dgx[2];
    OP_CHECK(upload_act(dx[0], x0, cin0, Vin, false, s));
    if (cin1) OP_CHECK(upload_act(dx[1], x1, cin1, Vin, false, s));
    OP_CHECK(upload_act(ddy, dy, cout, Vout, false, s));
    const size_t wcount = size_t(cout) * (cin0 + cin1) * (transposed ? 8 : ks * ks * ks);
    if (dw.alloc(wcount * 4) || dgw.alloc(wcount * 4)) { set_error("cudaMalloc failed"); return 1; }
    OP_CUDA(cudaMemcpy(dw.p, weight, wcount * 4, cudaMemcpyHostToDevice));
    OP_CUDA(cudaMemset(dgw.p, 0, wcount * 4));
    const int cin[2] = {cin0, cin1};
    float* gx[2] = {gx0, gx1};
    // data gradients
    for (int src = 0; src < 2; ++src) {
        if (!cin[src] || !gx[src]) continue;
        std::vector<ConvProblem> probs;
        std::vector<PackDesc> packs;
        int kc = 0;
        plan_dgrad(g, src, probs, packs, kc, conv_band_wants_kc16(ks, stride, transposed, pad16(cout), pad16(cin[src]), Vin) ? 16 : 0);
        const bool acc = src == 0 && (accumulate_gx0 & 1);
        if (acc) OP_CHECK(upload_act(dgx[src], gx[src], cin[src], Vin, false, s));
        else {
            if (dgx[src].alloc(size_t(pad16(cin[src])) * Vin * 2)) { set_error("cudaMalloc failed"); return 1; }
            OP_CUDA(cudaMemset(dgx[src].p, 0xff, size_t(pad16(cin[src])) * Vin * 2));
        }
        std::vector<DevBuf> wp(packs.size());
        for (size_t i = 0; i < packs.size(); ++i) {
            if (wp[i].alloc(pack_bytes(packs[i]))) { set_error("cudaMalloc failed"); return 1; }
            packs[i].w = static_cast<const float*>(dw.p);
            packs[i].out = wp[i].p;
            packs[i].out_bf16 = 0;
            OP_CHECK(pack_weights_launch(packs[i], s));
            probs[i].src0 = ddy.p;
            probs[i].dst = dgx[src].p;
            probs[i].wpack = wp[i].p;
        }
        ConvLaunch cfg{};
        cfg.kc = kc;
        cfg.epi = acc ? EPI_ACCUM16 : EPI_STORE16;
        DevBuf dsplit;
        if (Vin <= 16384 && !dsplit.alloc(size_t(48) << 20)) { cfg.splitk_scratch = static_cast<float*>(dsplit.p); cfg.splitk_scratch_bytes = size_t(48) << 20; }
        OP_CHECK(conv_launch(probs, cfg, s));
        OP_CHECK(finish(s));
        OP_CHECK(download_act(dgx[src].p, gx[src], cin[src], Vin, false, s));
    }
    // weight gradient
    if (gw) {
        std::vector<WgradProblem> wprobs;
        for (int src = 0; src < 2; ++src) {
            if (!cin[src]) continue;
            WgradProblem W;
            plan_wgrad(g, src, W);
            if (!transposed) { W.T = dx[src].p; W.U = ddy.p; }
            else { W.T = ddy.p; W.U = dx[0].p; }
            W.dw = static_cast<float*>(dgw.p);
            wprobs.push_back(W);
        }
        WgradLaunch wc{};
        DevBuf dpart;   // the partial-block scratch the model gives its weight gradients (U3D_WBAND_ATOMICS / U3D_WGRAD_ATOMICS fall back to atomics)
        if (!dpart.alloc(size_t(64) << 20)) { wc.partial_scratch = static_cast<float*>(dpart.p); wc.partial_scratch_bytes = size_t(64) << 20; }
        OP_CHECK(conv_wgrad_dispatch(wprobs, wc, s, nullptr));
        OP_CHECK(finish(s));
        OP_CUDA(cudaMemcpy(gw, dgw.p, wcount * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// MaxPool3d(2,2) with argmax (unet.cpp:38-39): y [C][D/2][H/2][W/2] and torch-style int64 indices (flat D*H*W offset)
extern "C" int u3d_op_maxpool_forward(int c, int w, int h, int d, const float* x, float* y, int64_t* indices) {
    cudaStream_t s = 0;
    const long long Vin = 1LL * w * h * d;
    const int ow = w / 2, oh = h / 2, od = d / 2;
    const long long Vout = 1LL * ow * oh * od;
    const int cp = pad16(c);
    DevBuf dx, dy, didx;
    OP_CHECK(upload_act(dx, x, c, Vin, false, s));
    if (dy.alloc(size_t(cp) * Vout * 2) || didx.alloc(size_t(cp) * Vout * 4)) { set_error("cudaMalloc failed"); return 1; }
    OP_CHECK(maxpool_fwd_launch(dx.p, dy.p, static_cast<int*>(didx.p), cp, od, oh, ow, s));
    OP_CHECK(finish(s));
    OP_CHECK(download_act(dy.p, y, c, Vout, false, s));
    std::vector<int> hi(size_t(cp) * Vout);
    OP_CUDA(cudaMemcpy(hi.data(), didx.p, hi.size() * 4, cudaMemcpyDeviceToHost));
    for (int ch = 0; ch < c; ++ch)
        for (long long v = 0; v < Vout; ++v) indices[size_t(ch) * Vout + v] = hi[size_t(v) * cp + ch];
    return 0;
}
