// Host-side UNet3d: feature_string parser, layer graph, memory plan and the forward / backward /
// optimizer executors that drive the sm_100a kernels.  Mirrors UNet3dImpl (/root/reference/unet.hpp:13-70).
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "elementwise.h"
#include "plan.h"

namespace u3d {

struct ModuleDef {
    enum Kind { CONV, CONVT, MAXPOOL, UPSAMPLE, NORM, BNORM, RELU, LEAKY, ELU } kind;
    int cin = 0, cout = 0, ks = 0, stride = 0;
    int p0 = -1;    // first parameter index (weight|gamma); p0+1 = bias|beta
    int buf0 = -1;  // BNORM: running_mean index; buf0+1 = running_var
};
struct BlockDef {
    std::string name;
    std::vector<ModuleDef> mods;
};
struct ParamInfo {
    std::string name;
    std::vector<int64_t> shape;
    long long offset = 0;  // element offset in the flat buffers (4-element aligned)
    long long numel = 0;
    bool decay = false;    // unet.cpp:252-258
};

struct Ten {  // NDHWC fp16 activation
    int C = 0, Cp = 0, d = 0, h = 0, w = 0;
    void* p = nullptr;
    void* grad = nullptr;
    bool needs_grad = false;
    long long V() const { return 1LL * d * h * w; }
    size_t bytes() const { return size_t(V()) * Cp * 2; }
};

struct Step {
    enum Kind { CONV, NORMACT, MAXPOOL, UPSAMPLE } kind = CONV;
    int in0 = -1, in1 = -1, out = -1;
    int head_level = -1;       // >= 0: this conv writes logits[head_level] (fp32 planar)
    bool head_fwd_fused = false;   // head on CUDA cores (head_fwd_kernel) instead of the tensor path
    bool head_bwd_fused = false;   // head dgrad/wgrad/bias gradient fused into the loss-gradient kernel
    // CONV
    LayerGeom g{};
    int p_w = -1, p_b = -1;
    bool stats = false;        // epilogue emits the statistics of the norm that follows
    bool drop_bias = false;    // InstanceNorm3d follows: the bias is not added (the norm removes any per-channel constant and the bias
                               // gradient is exactly zero), so the fp16 raw output is not rounded relative to an information-free offset
    std::vector<ConvProblem> fprobs;
    std::vector<PackDesc> fpacks;
    int fkc = 0;
    struct DG {
        std::vector<ConvProblem> probs;
        std::vector<PackDesc> packs;
        int kc = 0;
    } dg[2];
    std::vector<WgradProblem> wg;
    std::vector<void*> pack_bufs;  // owned device blobs (fwd then dgrad)
    double flops = 0;              // algorithmic 2*Cin*Cout*k^3*V_out of this layer
    int xf_from = -1;              // fused head: its input is the output of an ELIDED norm/activation step -- read that step's raw input and
                                   // apply scale/shift/activation in head_fwd_kernel (SrcTransform)
    // NORMACT
    int norm = 0;              // 0 none, 1 InstanceNorm3d (eps 1e-5), 2 BatchNorm3d (eps 0)
    int act = ACT_NONE;
    int p_g = -1;
    float* mean = nullptr;
    float* rstd = nullptr;
    int buf0 = -1;
    bool stats_from_conv = false;
    bool elided = false;       // the only reader is the fused head, which applies scale/shift/activation itself (SrcTransform): statistics only
    const float* cur_mean = nullptr;   // coefficients of the current forward pass (mode dependent), read by the consumers of an elided step
    const float* cur_rstd = nullptr;
    int cur_has_norm = 0;
    // MAXPOOL
    int* idx = nullptr;
};

class Model {
  public:
    // throws std::runtime_error (unet.cpp:53,66,88,117); host_only = structure without any device state
    Model(int in_c, int out_c, const std::string& feature, bool host_only = false);
    bool host_only = false;
    ~Model();

    // ---- reference-visible state (unet.hpp:16-18,37-38) ----
    int in_count, out_count;
    std::string architecture;
    int dim[3] = {192, 224, 192};  // W, H, D
    float voxel_size[3] = {1.f, 1.f, 1.f};
    // model-file metadata (unet.hpp:18-23; defaults of the constructor, unet.cpp:110-112)
    std::string preproc, postproc = "softmax+create_mask+argmax", orientation, fov_strategy = "align_top";
    std::vector<float> testing_errors, training_errors;   // 3 floats (ce, dice, mse) per step
    std::vector<int> single_component_label;
    const float* params_base() const { return d_params; }
    const float* momentum_base() const { return d_mom; }
    bool training = true;
    bool bn_running = false;       // eval() without prepare_for_inference: BatchNorm3d normalises with its running statistics (eps 0)

    std::vector<BlockDef> encoding, decoding, output, tail;
    std::vector<ParamInfo> params;
    int n_buffers = 0;
    long long flat_n = 0;

    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // weight gradients of a layer run here, concurrently with its data gradient on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pack = nullptr;
    // sample prefetch (train.cpp:446-485: augmentation workers run beside the trainer): upload + augmentation of the NEXT sample on a
    // third stream into one of two staging slots while the current micro-batch computes
    // data-parallel overlap: with an attached communicator the gradient bucket of everything behind the first (cheap-in-parameters,
    // expensive-in-time) encoder levels is all-reduced on a fourth stream as soon as its last contribution has been issued in the
    // backward pass of the step's LAST micro-batch; Model::step then only reduces the small prefix
    void* dp_comm = nullptr;
    int dp_microbatches = 1;         // micro-batches this rank runs per step
    int dp_seen = 0;                 // micro-batches since the last step
    bool dp_tail_reduced = false;    // the tail bucket [dp_split, flat_n) of this step is already in flight / reduced
    long long dp_split = 0;          // first element of the tail bucket
    int dp_split_step = -1;          // conv step (forward order) whose backward completes the tail bucket
    cudaStream_t stream4 = nullptr;
    cudaEvent_t ev_ar_ready = nullptr, ev_ar_done = nullptr;
    int attach_comm(void* comm, int microbatches_per_step);
    cudaStream_t stream3 = nullptr;
    std::vector<cudaEvent_t> ev_slabs;   // evaluate_volume: one per z slab of the host volume upload
    cudaEvent_t ev_sample[2] = {nullptr, nullptr};
    float* pf_in[2] = {nullptr, nullptr};
    float* pf_label[2] = {nullptr, nullptr};
    size_t pf_bytes = 0;
    int pf_next = 0;                 // slot the next prefetch writes
    int pf_head = 0;                 // oldest pending slot
    bool pf_pending[2] = {false, false};   // slot holds a prefetched sample that has not been consumed
    void pf_mark(int slot) { pf_pending[slot] = true; pf_next = slot ^ 1; }
    int prefetch_slot(float** in_dev, float** label_dev, int* slot);   // allocates / returns the staging slot to fill on stream3
    int consume_prefetched(float** in_dev, float** label_dev);          // main stream waits for the slot; returns its buffers
    bool pack_pending = false;       // a weight re-pack launched on stream2 after the last update has not been joined yet
    float* d_params = nullptr;
    float* d_grads = nullptr;
    float* d_mom = nullptr;
    std::vector<float*> d_buffers;   // BatchNorm running stats [C] each
    std::vector<int> buffer_len;
    int good_steps = 0;
    float loss_scale_max = 0.f;
    bool optimizer_created = false;
    bool mom_initialized = false;
    float lr0 = 0.f;
    float loss_scale = 0.f;          // 0 = choose from the volume size at plan time
    double last_grad_norm = 0.0;
    int last_step_skipped = 0;
    long long launches = 0;          // kernels launched by this handle (bench "gpu_launches")
    void* vpa_ws = nullptr;          // augmentation workspace (unet3d_vpa_augment)
    size_t vpa_ws_bytes = 0;
    void* pf_ws = nullptr;           // the same for the prefetch stream (unet3d_prefetch_augmented)
    size_t pf_ws_bytes = 0;
    int sim_mode = 0;                // simulate_modality before augmentation in the fused / prefetched sample calls: 0 off, 1 labelled, 2 image only

    int init_params(uint64_t seed);
    int get_flat(const float* base, int i, float* host, float scale);
    int set_param(int i, const float* host);
    int set_momentum(int i, const float* host);
    int set_dim(int w, int h, int d);
    int set_mode(int mode);   // 1 train(), 0 prepare_for_inference() (BatchNorm buffers reset to (0,1)), 2 eval() (running statistics)
    // where: 0 = host pointers, 1 = device pointers
    int forward(const float* in, float* const* out_levels, int n_levels, int where);
    int train_microbatch(const float* in, const float* label, int collapse_before, int use_ce, int use_dice, int use_mse,
                         float* loss_out3, float* all_levels, int where);
    int validate(const float* in, const float* label, int collapse_before, float* loss_out3, int where);   // loss_out3 == nullptr: asynchronous
    int validate_result(float* loss_out3);
    float* h_val = nullptr;          // pinned result of the last validation
    bool val_pending = false;
    // evaluate.cpp:223-230 over a list of windows; host pointers: upload of window i+1 and download of window i-1 overlap the
    // forward of window i (two staging slots each way, copy streams = the side streams that are idle during inference)
    int evaluate_windows(const float* const* in_windows, float* const* out_windows, int n_windows, int where);
    // evaluate one whole volume (any size): windows of the model grid, forward()[0] per window, softmax + re-assembly +
    // create_mask + argmax on the device (postproc.cu); label_out = 1 byte per voxel, fg_out / prob_out optional
    int evaluate_volume(const float* volume, int vw, int vh, int vd, int stride_x, int stride_y, int stride_z, float threshold,
                        uint8_t* label_out, float* fg_out, float* prob_out, int where, int* n_windows_out);
    float* ev_buf = nullptr;         // device: volume in, window in, accumulators, count, fg, labels
    size_t ev_bytes = 0;
    float* ew_in[2] = {nullptr, nullptr};
    float* ew_out[2] = {nullptr, nullptr};
    size_t ew_in_bytes = 0, ew_out_bytes = 0;
    // device staging buffers of the network input ([in][D][H][W] fp32) and label ([D][H][W] fp32) of the current plan
    int staging(float** in_dev, float** label_dev);
    int step(int batch_size, double lr, void* nccl_comm);
    int copy_from(const Model& src);
    int sync();
    int timer_start();            // CUDA events on this handle's stream
    int timer_stop(float* ms);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // per-launch CUDA-event profile of the tensor-core kernels (bench.py roofline):
    // kind 0 = conv_igemm, 1 = conv_wgrad, 2 = conv_s2, 3 = conv_wgrad_band, 4 = conv_tma, 5 = conv_band, 6 = conv_wgrad_quad
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;
    std::vector<double> prof_flops;
    size_t prof_used = 0;
    void trace_launch(const Step& s, const char* pass, int kind, int nprob, double flops);
    void prof_begin(int kind, double flops, cudaStream_t on = nullptr);
    void prof_end(cudaStream_t on = nullptr);
    int prof_read(double out[24], int reset);  // per kind (8): {ms, launches, algorithmic FLOPs}
    // CUDA graphs: the kernel sequence of a micro-batch / an inference forward is fixed for a plan, so its second occurrence with the
    // same buffers is captured (stream capture of the ordinary enqueue code, side stream included) and replayed afterwards: one
    // graph launch instead of ~250 kernel launches whose host cost paced the small deep-level kernels.  Keyed by the device pointers
    // and flags baked into the kernel arguments; dropped with the plan.  U3D_NO_GRAPH=1 disables.
    struct GraphEntry {
        std::vector<uint64_t> key;
        cudaGraphExec_t exec = nullptr;
        long long n_launches = 0;
        bool bad = false;     // capture failed once: keep enqueuing normally
    };
    std::vector<GraphEntry> graphs;
    int graph_run(const std::vector<uint64_t>& key, const std::function<int()>& body);
    void drop_graphs();
    int n_levels() const { return int(output.size()); }
    bool planned_for_pack_blobs() const { return planned; }   // the plan's weight blobs exist
    int resolve_status();            // waits for the last update's status read-back (lazy; see step.cpp)
    SgdStatus* h_status = nullptr;   // pinned
    cudaEvent_t ev_status = nullptr;
    bool status_pending = false;

  private:
    std::vector<Ten> tens;
    std::vector<Step> steps;
    std::vector<float*> logits;      // per level, fp32 planar
    std::vector<void*> dlogits;      // per level, fp16 [v][Cp]
    std::vector<int> level_dims;     // d,h,w per level
    std::vector<int> head_step;      // per level: index of the head conv in `steps` (-1 = none)
    bool planned = false, planned_training = false;
    bool packs_dirty = true;
    PackDesc* d_pack_descs = nullptr;   // every weight blob of the plan (forward then data-gradient packs of each conv), one launch
    int* d_pack_first = nullptr;
    int n_pack_jobs = 0, n_pack_blocks = 0;
    bool split_input = false;        // network input stored as fp16 [hi | lo | hi] in the padded channels (PackDesc::split_k)
    float* d_in_f32 = nullptr;
    float* d_label = nullptr;
    float* d_partials = nullptr;
    float* d_sums = nullptr;
    void* d_scratch = nullptr;       // gradient staging for multi-consumer tensors
    float* d_wgrad_partial = nullptr;   // per-CTA weight-gradient blocks of conv_wgrad_band (partial-block mode)
    size_t wgrad_partial_bytes = 0;
    float* d_splitk = nullptr;       // fp32 slices of the deterministic split-K convs of the deep levels (conv_tma.cu)
    size_t splitk_bytes = 0;
    size_t scratch_bytes = 0;
    double* d_loss_acc = nullptr;
    float* d_loss_part = nullptr;    // per level: [loss_part_rows()][loss_part_cols()] partial sums
    float* d_losses = nullptr;
    SgdChunk* d_chunks = nullptr;
    int n_chunks = 0;
    SgdStatus* d_status = nullptr;
    unsigned int* d_counter = nullptr;   // "last block finalizes" ticket of the norm-backward reduction (reset by that block)
    int last_stat_rows = 0, last_stat_ntot = 0;
    std::vector<char> grad_written;
    std::vector<char> skip_has_other_consumer;   // per step: concat conv whose skip tensor feeds an earlier conv as well (deferred skip gradient)
    std::vector<void*> owned;        // every cudaMalloc of the plan

    void free_plan();
    int ensure_plan();
    int build_steps();
    int alloc(void** p, size_t bytes);
    int repack();
    int repack_on(cudaStream_t s);
    int run_forward(int levels_wanted, bool bn_eval = false);   // bn_eval: BatchNorm3d uses (and does not update) its running statistics
    int run_backward();
    int upload_input(const float* in, int where);
    float* param_ptr(int i) { return d_params + params[i].offset; }
    float* grad_ptr(int i) { return d_grads + params[i].offset; }
};

std::string default_feature(int out_count);  // train.cpp:1054-1069

}  // namespace u3d
