/* C-ABI of libunet3d_b200.so — the B200-native drop-in for UNet-Studio's libtorch/cuDNN model path.
 *
 * The reference has no FFI layer: its boundary is the C++ class UNet3dImpl (/root/reference/unet.hpp:13-70)
 * plus the training-step body (train.cpp:628-706,755-766), the inference window loop
 * (evaluate.cpp:223-230) and visual_perception_augmentation (train.hpp:43-48).  Every entry point below
 * cites the reference interface it replaces.  Plain pointers and sizes only; no torch types.
 * include/unet3d.hpp wraps this ABI in a C++ class with the reference's member names.
 *
 * Conventions
 *  - every int function returns 0 on success, non-zero on failure; unet3d_last_error() then returns a
 *    thread-local message (constructor errors carry the reference's std::runtime_error text,
 *    unet.cpp:53,66,88,117).  Nothing throws or aborts (the reference turns worker exceptions into
 *    error_msg + aborted, train.cpp:709-721).
 *  - tensors are fp32, NCDHW with x fastest (train.cpp:619-621); labels are float-stored integers
 *    (train.cpp:615-617).  `where` = 0: host pointers (copies happen inside the call, the call returns
 *    after the results are in the caller's buffers); 1: device pointers on the handle's GPU
 *    (stream-ordered, call unet3d_sync() before reading).
 *  - a handle owns its device memory and one CUDA stream, belongs to one GPU, is not internally locked;
 *    distinct handles may be driven from distinct threads (train.cpp:592-600, validation replica
 *    train.cpp:834-840).  The library never keeps a caller pointer past the call (train.cpp:615-621).
 *  - arithmetic: fp16 operands with fp32 accumulation on tcgen05 tensor cores, fp32 statistics, losses,
 *    parameters, gradients and optimizer state (see DESIGN.md "precision").
 */
#ifndef UNET3D_B200_H
#define UNET3D_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct unet3d unet3d_t;

const char* unet3d_last_error(void);

/* default_feature(out_count), train.cpp:1054-1069.  Returns 0, or the needed buffer size if too small. */
int unet3d_default_feature(int out_count, char* buf, size_t buflen);

/* Host-only structure query (no GPU needed): parses the feature_string exactly like the constructor and
 * writes {"levels": L, "params": [{"name","shape","decay"}...]} (tensorN order) as JSON.  Same error
 * behaviour as unet3d_create.  Returns 0, 1 on a parse error, or the needed size if the buffer is too small. */
int unet3d_describe(int in_count, int out_count, const char* feature_string, char* json, size_t json_len);

/* UNet3d(in_count, out_count, feature_string), unet.cpp:103-166 (+ create_layer, unet.cpp:24-101).
 * The module starts in training mode like a fresh torch module. */
int unet3d_create(int in_count, int out_count, const char* feature_string, int gpu, unet3d_t** out);
void unet3d_destroy(unet3d_t* h);

/* public fields of UNet3dImpl: in_count, out_count, architecture, dim, voxel_size (unet.hpp:16-18,37-38) */
int unet3d_in_count(const unet3d_t* h);
int unet3d_out_count(const unet3d_t* h);
int unet3d_levels(const unet3d_t* h);                   /* output.size(): number of deep-supervision heads */
const char* unet3d_architecture(const unet3d_t* h);
int unet3d_set_dim(unet3d_t* h, int w, int hgt, int d); /* model->dim = {w,h,d} (train.cpp:1131) */
int unet3d_get_dim(const unet3d_t* h, int dim3[3]);
int unet3d_set_voxel_size(unet3d_t* h, float x, float y, float z);

/* parameters() in registration order == tensorN order of the .nz model file (main.cpp:193-204,225-231):
 * native contiguous element order, fp32. */
int unet3d_param_count(const unet3d_t* h);
long long unet3d_param_total(const unet3d_t* h);
int unet3d_param_shape(const unet3d_t* h, int i, int64_t dims[5], int* ndim);
const char* unet3d_param_name(const unet3d_t* h, int i);   /* named_parameters() key, e.g. "encode0.0.weight" */
int unet3d_param_decay(const unet3d_t* h, int i);          /* 1 if in the weight-decay group (unet.cpp:252-258) */
int unet3d_get_param(unet3d_t* h, int i, float* host);
int unet3d_set_param(unet3d_t* h, int i, const float* host);
int unet3d_get_grad(unet3d_t* h, int i, float* host);      /* accumulated .grad() of parameter i */
int unet3d_get_momentum(unet3d_t* h, int i, float* host);  /* SGD momentum_buffer (the .opt file, train.cpp:787) */
int unet3d_set_momentum(unet3d_t* h, int i, const float* host);
int unet3d_init_params(unet3d_t* h, uint64_t seed);        /* torch's default init rule, own RNG stream */

/* train(bool) / eval() (unet.hpp:58-62) and prepare_for_inference (unet.cpp:7-22):
 *   mode 1 = train(): BatchNorm3d uses batch statistics and updates running_mean / running_var (momentum 0.1, unbiased variance);
 *   mode 2 = eval():  BatchNorm3d normalises with its running statistics, eps 0 (what the validation replica runs, train.cpp:836);
 *   mode 0 = prepare_for_inference(): eval() + every BatchNorm3d reset to running_mean 0 / running_var 1, i.e. y = gamma*x + beta.
 * InstanceNorm3d always uses the statistics of the sample. */
int unet3d_set_mode(unet3d_t* h, int mode);

/* forward(Tensor{1,in,D,H,W}) -> vector<Tensor> (unet.hpp:51, unet.cpp:168-193).  out_levels[k] receives
 * results[k] ({1,out,D>>k,H>>k,W>>k}); only the first n_levels heads are computed (inference uses
 * n_levels = 1, evaluate.cpp:226). */
int unet3d_forward(unet3d_t* h, const float* in, float* const* out_levels, int n_levels, int where);

/* the inference window loop, evaluate.cpp:223-230: out_windows[i] = forward(in_windows[i])[0] */
int unet3d_evaluate_windows(unet3d_t* h, const float* const* in_windows, float* const* out_windows, int n_windows, int where);

/* One whole volume through evaluate (evaluate.cpp:195-230, 274): the volume ({in,d,hgt,w} fp32, any size) is covered by windows of
 * the model grid (unet3d_set_dim) at the given stride per axis (<= 0: the window size, i.e. no overlap; the last window of an axis
 * is shifted inward to end at the border; a smaller volume is zero-padded at the far end), every window runs forward()[0], and the
 * default postproc "softmax+create_mask+argmax" (unet.cpp:112) runs on the GPU over the re-assembled volume: probabilities of
 * overlapping windows are averaged, fg_prob = 1 - p(channel 0), label = fg_prob > mask_threshold ? arg-max channel : 0
 * (evaluate.cpp:315-319).  label_out: 1 byte per voxel; fg_prob_out ({d,hgt,w} fp32) and label_prob_out ({out,d,hgt,w} fp32) may
 * be NULL.  The window cutting / re-assembly / postproc arithmetic of the reference lives in TIPL (not vendored): parity unpinned,
 * semantics restated in oracle/postproc_oracle.py.  out_count <= 32. */
int unet3d_evaluate_volume(unet3d_t* h, const float* volume, int w, int hgt, int d, int stride_x, int stride_y, int stride_z,
                           float mask_threshold, uint8_t* label_out, float* fg_prob_out, float* label_prob_out, int where, int* n_windows);
/* host only: the window start positions unet3d_evaluate_volume uses along one axis; returns their number */
int unet3d_window_origins(int volume_dim, int window_dim, int stride, int* origins, int max_origins);
/* standalone "softmax+create_mask+argmax" on host logits ({channels, voxels} fp32 planar) */
int u3d_postproc(const float* logits, int channels, long long voxels, float mask_threshold, uint8_t* label_out, float* fg_prob_out,
                 float* label_prob_out, int gpu);
/* resampling to / from the model grid (the step in front of the windows; tipl::scale-style index mapping dst*src_dim/dst_dim):
 * trilinear for images, nearest for label maps.  Host buffers {channels, d, h, w}. */
int u3d_resample(const float* src, int channels, int sw, int sh, int sd, float* dst, int dw, int dh, int dd, int nearest, int gpu);

/* one N=1 micro-batch of the training step body, train.cpp:628-706: forward, calc_losses on every
 * deep-supervision level (train.cpp:501-552), level weights (1/2^k)/sum, backward.  Gradients ACCUMULATE
 * across calls until unet3d_step.  loss_out = level-0 {ce, dice, mse} (what the reference logs,
 * train.cpp:675-681); all_level_losses (optional) = 3 floats per level.
 * Limit of this build: the fused loss head holds one voxel's class vector in registers and supports out_count <= 32 (the reference
 * has no such limit; its shipped models use 2..6 classes).  Larger models can be created, loaded and run through
 * unet3d_forward / unet3d_evaluate_windows, but unet3d_train_microbatch / unet3d_validate fail with
 * "loss head supports 1..32 output channels". */
int unet3d_train_microbatch(unet3d_t* h, const float* in, const float* label, int collapse_before, int use_ce, int use_dice,
                            int use_mse, float loss_out[3], float* all_level_losses, int where);

/* validation forward + level-0 calc_losses, no gradient (train.cpp:826-851).  Runs with eval() semantics whatever the handle's mode
 * (BatchNorm3d reads its running statistics, nothing is updated). */
int unet3d_validate(unet3d_t* h, const float* in, const float* label, int collapse_before, float loss_out[3], int where);

/* The reference validates on its own replica (output_model) in its own thread, beside the trainer (train.cpp:826-851, 773-776).
 * Same structure here: a second handle (unet3d_copy_from under the caller's mutex) owns its own CUDA stream, so its validation
 * runs beside the training handle's kernels on the same GPU.  unet3d_validate_async only enqueues (host pointers with where = 0
 * must stay valid until unet3d_validate_result returns; pinned memory keeps the upload asynchronous); unet3d_validate_result waits
 * for that handle's stream alone and returns the level-0 {ce, dice, mse}. */
int unet3d_validate_async(unet3d_t* h, const float* in, const float* label, int collapse_before, int where);
int unet3d_validate_result(unet3d_t* h, float loss_out[3]);

/* create_optimizer(lr) (unet.cpp:246-277): SGD momentum .99, Nesterov, weight decay 3e-5 | 0 groups */
int unet3d_create_optimizer(unet3d_t* h, float learning_rate);

/* update, train.cpp:755-766: [sum of the replicas' gradients: ONE ncclAllReduce over the flat gradient
 * buffer when nccl_comm != NULL, replacing add_gradient_from + the copy_from weight broadcast],
 * grad /= batch_size, clip_grad_norm_(12), optimizer.step(), zero_grad().  lr = the step's learning rate
 * (train.cpp:566-571). */
int unet3d_step(unet3d_t* h, int batch_size, double lr, void* nccl_comm);
double unet3d_last_grad_norm(const unet3d_t* h);   /* pre-clip global norm of the last step */
int unet3d_last_step_skipped(const unet3d_t* h);   /* 1 if the fp16 gradient path overflowed and the update was skipped */
float unet3d_loss_scale(const unet3d_t* h);
int unet3d_set_loss_scale(unet3d_t* h, float s);
long long unet3d_launch_count(const unet3d_t* h);  /* kernels launched by this handle so far */

/* ---- model file (.nz): load_from_file / save_to_file, main.cpp:157-233 -------------------------------------------------------
 * A gzip stream of MATLAB Level-4 MAT matrices: channels, architecture, dimension, voxel_size, fov_strategy, preproc, orientation,
 * postproc, training_errors, testing_errors, [single_component_label], tensor{i} = parameters()[i] as rows = numel/size(0),
 * cols = size(0) in native element order.  The Level-4 container is pinned against scipy.io (tests/test_modelfile_cpu.py); TIPL's
 * "sloped" integer encoding of tensor{i} is NOT vendored (parity unpinned): the reader converts any numeric type and applies
 * "<name>.slope" / "<name>.inter" companions if present, the writer stores fp32 exactly.
 * unet3d_load_from_file constructs the model from the file's channels + architecture (same errors as unet3d_create;
 * "invalid format" / "tensor size mismatch at tensorN ..." like main.cpp:166,177,199) and leaves it in training mode (main.cpp:193). */
int unet3d_load_from_file(const char* file_name, int gpu, unet3d_t** out);
int unet3d_save_to_file(unet3d_t* h, const char* file_name);
/* train.cpp:787,945-957 keep the SGD state next to the model as <model>.opt via torch::save (a libtorch pickle archive); here the
 * momentum buffers go into the same .nz container (momentum{i} + sgd_state). */
int unet3d_save_optimizer(unet3d_t* h, const char* file_name);
int unet3d_load_optimizer(unet3d_t* h, const char* file_name);
/* raw export of the same logical layout: <directory>/tensor{i}.bin (fp32, native order) + <directory>/model.json */
int unet3d_export_raw(unet3d_t* h, const char* directory);
/* metadata strings of UNet3dImpl (unet.hpp:18): key = "preproc" | "postproc" | "orientation" | "fov_strategy" */
int unet3d_set_info(unet3d_t* h, const char* key, const char* value);
int unet3d_get_info(unet3d_t* h, const char* key, char* buf, size_t buflen);
/* training_errors / testing_errors (unet.hpp:22; 3 floats ce, dice, mse per step).  get returns the number of steps stored. */
int unet3d_set_errors(unet3d_t* h, int testing, const float* ce_dice_mse, int n_steps);
int unet3d_get_errors(unet3d_t* h, int testing, float* ce_dice_mse, int max_steps);

/* the container itself, host only (no GPU): type = Level-4 code (0 f64, 10 f32, 20 i32, 30 i16, 40 u16, 50 u8, +1 = text) */
typedef struct u3d_nz u3d_nz_t;
int u3d_nz_create(u3d_nz_t** out);
int u3d_nz_load(const char* path, u3d_nz_t** out);
int u3d_nz_save(const u3d_nz_t* f, const char* path);
void u3d_nz_free(u3d_nz_t* f);
int u3d_nz_count(const u3d_nz_t* f);
int u3d_nz_info(const u3d_nz_t* f, int i, char* name, size_t name_len, int* type, int* rows, int* cols);
int u3d_nz_add(u3d_nz_t* f, const char* name, int type, int rows, int cols, const void* data);   /* column-major data */
int u3d_nz_read_f32(const u3d_nz_t* f, const char* name, float* out, size_t n);

/* copy_from (unet.cpp:195-222): parameters/buffers of identical size, dim, voxel_size; works across GPUs */
int unet3d_copy_from(unet3d_t* dst, const unet3d_t* src);
int unet3d_sync(unet3d_t* h);
/* CUDA-event timer on the handle's own stream (the stream every kernel of the handle is launched on) */
int unet3d_timer_start(unet3d_t* h);
int unet3d_timer_stop(unet3d_t* h, float* ms);

/* Per-launch CUDA-event profile of the tensor-core kernels on the handle's stream (for the bench roofline):
 * out24 = {ms, launches, algorithmic FLOPs} for each of conv_igemm, conv_wgrad, conv_s2, conv_wgrad_band, conv_tma, conv_band,
 * conv_wgrad_quad and one reserved family (in this order)
 * since the last reset; algorithmic FLOPs = 2*Cin*Cout*k^3*V_out per layer. */
int unet3d_profile(unet3d_t* h, int enable);
int unet3d_profile_read(unet3d_t* h, double out24[24], int reset);

/* visual_perception_augmentation(options, image, label, is_label, shape, seed) (train.hpp:43-48,
 * visual_perception_augmentation.cpp:163-438): in place on `image` ({channels,D,H,W} fp32 = tipl::image<3> with the
 * channels stacked along z) and `label` ({D,H,W} fp32).  options = parallel arrays of option ids (options.txt:1-39)
 * and values; an id that is not listed reads as 0 like unordered_map::operator[].  The random scalars follow the
 * reference's draw order from std::mt19937(seed).  where = 0 host pointers, 1 device pointers.
 * vpa_augment is standalone (own stream on `gpu`, returns when done); unet3d_vpa_augment runs stream-ordered on the
 * handle's stream so the augmented sample feeds unet3d_train_microbatch(where=1) without leaving HBM. */
int vpa_augment(const char* const* keys, const float* vals, int n_opts, float* image, float* label, int is_label, int w, int h, int d,
                int channels, uint64_t seed, int where, int gpu);
int unet3d_vpa_augment(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, float* image, float* label, int is_label,
                       int w, int hgt, int d, int channels, uint64_t seed, int where);

/* host only (no GPU needed): the scalars the augmentation draws for (options, shape, seed) in the reference's draw order
 * (visual_perception_augmentation.cpp:180-320): the output->source affine M12 (3x4, row major), the perspective coefficients,
 * and the local distortion foci (x, y, z, radius, magnitude; up to 12).  Any output pointer may be NULL. */
int vpa_plan_describe(const char* const* keys, const float* vals, int n_opts, int is_label, int w, int h, int d, int channels,
                      uint64_t seed, float M12[12], float persp3[3], int* nfoci, float* foci5);

/* host only: the Perlin background of the same plan (visual_perception_augmentation.cpp:386-418): returns through *applies whether
 * the stage runs for (options, seed); perm512 = the permutation std::shuffle(p, std::mt19937(seed)) produced, zoom = its drawn zoom. */
int vpa_plan_perlin(const char* const* keys, const float* vals, int n_opts, int is_label, int w, int h, int d, int channels,
                    uint64_t seed, int* applies, int* perm512, float* zoom);

/* train.cpp:459-473 + 615-706 in one call: upload the RAW sample once (host pointers, label = float-stored integers), run
 * visual_perception_augmentation on it in HBM (is_label = 1) and feed the result straight into the micro-batch.  Saves the
 * device->host->device round trip of the augmented sample that the two separate where = 0 calls make. */
int unet3d_train_microbatch_augmented(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, const float* image_host,
                                      const float* label_host, uint64_t seed, int collapse_before, int use_ce, int use_dice, int use_mse,
                                      float loss_out3[3]);

/* The reference augments samples in worker threads beside the trainer (train.cpp:446-485, slots in_data[i] / data_ready[i]).
 * unet3d_prefetch_augmented starts upload (where = 0: host pointers, which must stay valid until the matching
 * unet3d_train_microbatch_prefetched returns; where = 1: device pointers) + augmentation of the NEXT sample on a side stream and
 * returns immediately; unet3d_train_microbatch_prefetched makes the micro-batch wait for it and consumes it.  Call order per
 * step: prefetch(i+1), train_prefetched(i), step (after an initial prefetch(0)).  Two staging slots, consumed first-in first-out. */
int unet3d_prefetch_augmented(unet3d_t* h, const char* const* keys, const float* vals, int n_opts, const float* image, const float* label,
                              uint64_t seed, int where);
int unet3d_train_microbatch_prefetched(unet3d_t* h, int collapse_before, int use_ce, int use_dice, int use_mse, float loss_out3[3]);

/* simulate_modality (train.cpp:43-117 labelled-template overload, :119-180 image-only overload; call site train.cpp:459-462):
 * synthesises a random contrast in place on `t1w` ({D,H,W} fp32 in [0,1]).  label = NULL selects the image-only overload;
 * otherwise label holds float-stored integers 0..max_label (the caller passes model->out_count, train.cpp:459).  seed is the
 * sample seed (rand_int = mt19937(seed), rand_float = mt19937(seed+1)).  where = 0 host pointers, 1 device pointers.
 * simulate_modality is standalone (own stream on `gpu`, returns when done); unet3d_simulate_modality runs stream-ordered on the
 * handle's stream.  unet3d_set_simulate_modality makes unet3d_train_microbatch_augmented / unet3d_prefetch_augmented run it on the
 * uploaded sample before the augmentation, like the reference's augmentation thread: mode 0 off (default), 1 labelled template
 * (train_image_is_template), 2 image only. */
int simulate_modality(float* t1w, const float* label, unsigned max_label, unsigned seed, int w, int h, int d, int where, int gpu);
int unet3d_simulate_modality(unet3d_t* h, float* t1w, const float* label, unsigned max_label, unsigned seed, int w, int hgt, int d,
                             int where);
int unet3d_set_simulate_modality(unet3d_t* h, int mode);
/* host only (no GPU needed): the random scalars simulate_modality draws for (overload, max_label, seed) in the reference's draw
 * order (train.cpp:56-58 tissue LUT, :65-78 the 20 terms, :80 gamma): lut_out[max_label + 1] (labelled overload only),
 * terms_out[20][5] = a, b, c, d, w, gamma_out[1].  Any output pointer may be NULL. */
int simulate_modality_plan(int labelled, unsigned max_label, unsigned seed, float* lut_out, float* terms_out, float* gamma_out);

/* NCCL plumbing for the data-parallel step (bootstrap the 128-byte id through any host channel) */
int unet3d_nccl_unique_id(void* id128);
int unet3d_nccl_comm_init(void** comm, int nranks, int rank, const void* id128);
int unet3d_nccl_comm_destroy(void* comm);
/* Optional: tell the handle its communicator and how many micro-batches THIS rank runs per step.  The backward pass of the last
 * micro-batch then all-reduces the gradient bucket of everything behind the first encoder levels (94 % of the parameters of the
 * default net) on its own stream as soon as that bucket is complete, overlapping the rest of the backward pass; unet3d_step (same
 * comm) reduces only the remaining prefix.  Results are identical to the un-attached path up to the order of fp32 additions inside
 * NCCL.  comm = NULL detaches. */
int unet3d_attach_comm(unet3d_t* h, void* comm, int microbatches_per_step);

/* ---- operator level (one reference layer, host buffers) — used by the per-layer parity tests ---------
 * Conv3d k1s1|k3s1|k3s2 pad (k-1)/2 (unet.cpp:59-72) or ConvTranspose3d k2s2 (unet.cpp:46-57), input given
 * as up to two tensors x0|x1 whose channel concat {x0,x1} (unet.cpp:181) is folded into the GEMM K loop.
 * weight/bias use the reference parameter layout ([Cout][Cin][k][k][k]; conv_trans [Cin][Cout][2][2][2]).
 * stats_sum_sumsq (optional, 2*cout doubles): per-channel sum and sum of squares of y, as produced for the
 * following InstanceNorm3d.  planar_fp32 != 0 selects the logits epilogue (fp32 NCDHW, no 16-bit rounding). */
int u3d_op_conv_forward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                        const float* x0, const float* x1, const float* weight, const float* bias, float* y,
                        double* stats_sum_sumsq, int planar_fp32);
/* autograd of the same layer (train.cpp:706): gx0/gx1 data gradients (NULL to skip), gw weight gradient.
 * flags bit 0: accumulate into gx0 (skip connections sum two data gradients). */
int u3d_op_conv_backward(int transposed, int ks, int stride, int cin0, int cin1, int cout, int w, int h, int d,
                         const float* x0, const float* x1, const float* weight, const float* dy, float* gx0,
                         float* gx1, float* gw, int flags);

/* MaxPool3d(2,2) (unet.cpp:38-39): values and torch-convention int64 argmax indices (flat D*H*W input offset,
 * first maximum in d,h,w scan order, NaN propagates) — bit-exact against torch.max_pool3d(return_indices). */
int u3d_op_maxpool_forward(int c, int w, int h, int d, const float* x, float* y, int64_t* indices);

#ifdef __cplusplus
}
#endif
#endif
