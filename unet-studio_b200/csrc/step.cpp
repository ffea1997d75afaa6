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

int Model::step(int batch_size, double lr, void* nccl_comm) {
    if (!optimizer_created) { set_error("create_optimizer has not been called"); return 1; }
    if (batch_size < 1) { set_error("batch_size must be >= 1"); return 1; }
    cudaSetDevice(device);
    if (loss_scale == 0.f) { set_error("step called before any micro-batch"); return 1; }
    if (nccl_comm != nullptr) {
        // the tail bucket may already have been reduced on stream4 during the backward pass (Model::run_backward)
        const size_t count = dp_tail_reduced ? size_t(dp_split) : size_t(flat_n);
        if (count) {
            ncclResult_t r = ncclAllReduce(d_grads, d_grads, count, ncclFloat, ncclSum, static_cast<ncclComm_t>(nccl_comm), stream);
            if (r != ncclSuccess) { set_error(std::string("ncclAllReduce: ") + ncclGetErrorString(r)); return 1; }
        }
        if (dp_tail_reduced && cudaStreamWaitEvent(stream, ev_ar_done, 0) != cudaSuccess) { set_error("step: event"); return 1; }
    }
    dp_tail_reduced = false;
    dp_seen = 0;
    const float inv = 1.0f / (loss_scale * float(batch_size));
    if (sgd_step_launch(d_params, d_grads, d_mom, flat_n, d_chunks, n_chunks, inv, float(lr), 0.99f, 12.0f, mom_initialized ? 0 : 1,
                        d_status, stream))
        return 1;
    launches += 2;
    // re-pack the fp16 weight blobs right away on the side stream: ordered after the update, overlapping whatever the main stream does
    // next (read-back below, the next sample's upload and augmentation); the next forward joins it (Model::repack)
    static const bool no_side = std::getenv("U3D_ONE_STREAM") != nullptr;
    bool async_pack = false;
    if (planned_for_pack() && stream2 != nullptr && !no_side) {
        if (cudaEventRecord(ev_fork, stream) != cudaSuccess || cudaStreamWaitEvent(stream2, ev_fork, 0) != cudaSuccess) { set_error("step: event"); return 1; }
        if (repack_on(stream2)) return 1;
        if (cudaEventRecord(ev_pack, stream2) != cudaSuccess) { set_error("step: event"); return 1; }
        pack_pending = true;
        async_pack = true;
    }
    SgdStatus st{};
    cudaError_t e = cudaMemcpyAsync(&st, d_status, sizeof(st), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return 1; }
    if (sync()) return 1;
    last_grad_norm = std::sqrt(st.sumsq);
    last_step_skipped = (st.nonfinite != 0 || !std::isfinite(st.sumsq)) ? 1 : 0;
    if (loss_scale_max == 0.f) loss_scale_max = loss_scale;
    if (last_step_skipped) {
        loss_scale = std::max(1.0f, loss_scale / 16.0f);   // fp16 gradient overflow: drop the scale, skip this update
        good_steps = 0;
    } else {
        mom_initialized = true;
        if (++good_steps >= 500 && loss_scale < loss_scale_max) { loss_scale *= 2.0f; good_steps = 0; }
        packs_dirty = !async_pack;
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
