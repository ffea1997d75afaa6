// Optimizer step with the data-parallel gradient exchange.
// Reference: the gradient "collective" is add_gradient_from (p2p copy + add_ onto GPU 0, unet.cpp:224-244)
// followed by grad/batch, clip, SGD on GPU 0 and a weight re-broadcast by copy_from at the next step
// (train.cpp:573-579,755-766).  Here every rank owns a replica: ONE ncclAllReduce(sum) over the flat fp32
// gradient buffer, then the identical clip + Nesterov SGD on every rank, so no weight broadcast is needed.
#include <nccl.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "model.h"

namespace u3d {

int Model::attach_comm(void* comm, int microbatches_per_step) {
    if (microbatches_per_step < 1) { set_error("attach_comm: micro-batches per step must be >= 1"); return 1; }
    cudaSetDevice(device);
    if (comm != nullptr && stream4 == nullptr) {
        if (cudaStreamCreateWithFlags(&stream4, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_ar_ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_ar_done, cudaEventDisableTiming) != cudaSuccess) {
            set_error("attach_comm: stream/event creation failed");
            return 1;
        }
    }
    dp_comm = comm;
    dp_microbatches = microbatches_per_step;
    dp_seen = 0;
    dp_tail_reduced = false;
    return 0;
}

// The status of an update (pre-clip gradient norm, overflow flag) is read back lazily: Model::step only enqueues the copy into pinned
// host memory; the first caller that needs it (the next micro-batch for its loss scale, the next step, or the accessors) waits.
int Model::resolve_status() {
    if (!status_pending) return 0;
    cudaSetDevice(device);
    if (cudaEventSynchronize(ev_status) != cudaSuccess) { set_error("step: status read-back failed"); return 1; }
    status_pending = false;
    const unsigned int code = read_device_error();
    if (code) { set_error("device pipeline timeout code " + std::to_string(code)); return 1; }
    const SgdStatus& st = *h_status;
    last_grad_norm = std::sqrt(st.sumsq);
    last_step_skipped = (st.nonfinite != 0 || !std::isfinite(st.sumsq)) ? 1 : 0;
    if (loss_scale_max == 0.f) loss_scale_max = loss_scale;
    if (last_step_skipped) {
        loss_scale = std::max(1.0f, loss_scale / 16.0f);   // fp16 gradient overflow: drop the scale, the update was skipped on the device
        good_steps = 0;
    } else {
        mom_initialized = true;
        if (++good_steps >= 500 && loss_scale < loss_scale_max) { loss_scale *= 2.0f; good_steps = 0; }
    }
    return 0;
}

int Model::step(int batch_size, double lr, void* nccl_comm) {
    if (!optimizer_created) { set_error("create_optimizer has not been called"); return 1; }
    if (batch_size < 1) { set_error("batch_size must be >= 1"); return 1; }
    cudaSetDevice(device);
    if (resolve_status()) return 1;
    if (dp_comm != nullptr && nccl_comm != nullptr && nccl_comm != dp_comm) {
        set_error("unet3d_step: the communicator differs from the one given to unet3d_attach_comm");
        return 1;
    }
    // a rank that ran no micro-batch in this step (batch_size < world: the reference uses min(gpus, batch) workers, train.cpp:592) still
    // takes part in the collectives with its zero gradients
    if (loss_scale == 0.f && ensure_plan()) return 1;
    // (a tail bucket reduced during the backward pass has already joined the main stream at the end of that pass)
    if (nccl_comm != nullptr) {
        // every rank issues the SAME collective sequence whatever it did in its backward passes.  Attached handles: tail bucket
        // [dp_split, flat_n) then prefix [0, dp_split); the tail may already be in flight on stream4 (Model::run_backward).  A rank that
        // declared more micro-batches than it ran reduces the tail here.  Un-attached handles: one all-reduce of the whole buffer.
        ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
        const bool split = dp_comm != nullptr && dp_split_step >= 0 && dp_split > 0 && dp_split < flat_n;
        ncclResult_t r = ncclSuccess;
        if (split) {
            if (!dp_tail_reduced) r = ncclAllReduce(d_grads + dp_split, d_grads + dp_split, size_t(flat_n - dp_split), ncclFloat, ncclSum, comm, stream);
            if (r == ncclSuccess) r = ncclAllReduce(d_grads, d_grads, size_t(dp_split), ncclFloat, ncclSum, comm, stream);
        } else
            r = ncclAllReduce(d_grads, d_grads, size_t(flat_n), ncclFloat, ncclSum, comm, stream);
        if (r != ncclSuccess) { set_error(std::string("ncclAllReduce: ") + ncclGetErrorString(r)); return 1; }
    }
    dp_tail_reduced = false;
    dp_seen = 0;
    const float inv = 1.0f / (loss_scale * float(batch_size));
    if (sgd_step_launch(d_params, d_grads, d_mom, flat_n, d_chunks, n_chunks, inv, float(lr), 0.99f, 12.0f, mom_initialized ? 0 : 1,
                        d_status, stream))
        return 1;
    launches += 2;
    if (!h_status) {
        if (cudaMallocHost(reinterpret_cast<void**>(&h_status), sizeof(SgdStatus)) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_status, cudaEventDisableTiming) != cudaSuccess) { set_error("step: pinned status"); return 1; }
    }
    if (cudaMemcpyAsync(h_status, d_status, sizeof(SgdStatus), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaEventRecord(ev_status, stream) != cudaSuccess) { set_error("step: status copy"); return 1; }
    status_pending = true;
    // re-pack the fp16 weight blobs right away on the side stream: ordered after the update, overlapping whatever the main stream does
    // next (the next sample's upload and augmentation); the next forward joins it (Model::repack).  A skipped update (gradient
    // overflow) leaves the weights unchanged, so the re-pack is then merely redundant.
    static const bool no_side = std::getenv("U3D_ONE_STREAM") != nullptr;
    packs_dirty = true;
    if (planned_for_pack_blobs() && stream2 != nullptr && !no_side) {
        if (cudaEventRecord(ev_fork, stream) != cudaSuccess || cudaStreamWaitEvent(stream2, ev_fork, 0) != cudaSuccess) { set_error("step: event"); return 1; }
        if (repack_on(stream2)) return 1;
        if (cudaEventRecord(ev_pack, stream2) != cudaSuccess) { set_error("step: event"); return 1; }
        pack_pending = true;
        packs_dirty = false;
    }
    return 0;
}

int nccl_unique_id(void* out128) {
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) { set_error(std::string("ncclGetUniqueId: ") + ncclGetErrorString(r)); return 1; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(out128, &id, 128);
    return 0;
}

int nccl_comm_init(void** comm, int nranks, int rank, const void* id128) {
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t c;
    ncclResult_t r = ncclCommInitRank(&c, nranks, id, rank);
    if (r != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); return 1; }
    *comm = c;
    return 0;
}

int nccl_comm_destroy(void* comm) {
    if (comm) ncclCommDestroy(static_cast<ncclComm_t>(comm));
    return 0;
}

}  // namespace u3d
