"""Host-side data-parallel plumbing (one process per GPU; torch.distributed is only the bootstrap channel).

Replaces the thread-per-GPU / atomic work-stealing scheme of train.cpp:573-594,604: micro-batch b of a step is
owned by rank b % world (static; the gradient sum is order independent up to fp rounding), every rank runs its
micro-batches, then ONE all-reduce(sum) of the flat gradient and the identical update on every rank."""
import ctypes


def shard_microbatches(batch_size, world, rank):
    """Indices of the micro-batches of one optimizer step that `rank` runs (train.cpp:604-608: b in [0,batch_size))."""
    return [b for b in range(batch_size) if b % world == rank]


def sample_seed(step, batch_size, b):
    """Seed of micro-batch b of `step` (train.cpp:394-401,608: data index = step*batch_size + b)."""
    return step * batch_size + b


def bootstrap_nccl(pkg, dist_mod, world, rank):
    """Creates the library's own ncclComm_t: rank 0 makes the unique id, torch.distributed carries the 128 bytes."""
    ids = [None]
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        pkg.check(pkg.lib().unet3d_nccl_unique_id(buf))
        ids = [bytes(buf.raw)]
    dist_mod.broadcast_object_list(ids, src=0)
    comm = ctypes.c_void_p()
    pkg.check(pkg.lib().unet3d_nccl_comm_init(ctypes.byref(comm), world, rank, ids[0]))
    return comm


def shard_windows(n_windows, world, rank):
    """Indices of the inference windows (evaluate.cpp:223-230 iterates model_io[i]) that `rank` forwards: window i -> rank
    i % world.  No collective: every rank writes its own windows' outputs and the host gathers them."""
    return [i for i in range(n_windows) if i % world == rank]
