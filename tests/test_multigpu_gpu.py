"""2-GPU data-parallel step through the C-ABI + NCCL (skipped with fewer than 2 GPUs): two ranks with one micro-batch each
must end with identical parameters, equal (up to fp16/atomic rounding) to one process running both micro-batches."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
FEATURE = ("conv16,ks3,stride1+norm,leaky_relu\nconv32,ks3,stride2+norm,leaky_relu+conv_trans16,ks2,stride2\n"
           "conv16,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
W, H, D = 32, 32, 32


def _data(b):
    rng = np.random.default_rng(50 + b)
    x = rng.random((1, 1, D, H, W), dtype=np.float32)
    t = (rng.random((1, D, H, W)) > 0.5).astype(np.float32)
    return x, t


def _worker(rank, world, port, outdir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from tests._pkg import load
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = load()
    comm = m.dist.bootstrap_nccl(m, dist, world, rank)
    net = m.UNet3d(1, 2, FEATURE, gpu=rank)
    net.init_params(7)
    net.set_dim(W, H, D); net.train(True); net.create_optimizer(0.01)
    net.attach_comm(comm, 1)      # the tail gradient bucket is all-reduced during the backward pass
    for step in range(2):
        for b in m.dist.shard_microbatches(world, world, rank):
            x, t = _data(m.dist.sample_seed(step, world, b))
            net.train_microbatch(x, t)
        net.step(world, 0.01, comm)
    np.save(os.path.join(outdir, f"p{rank}.npy"), np.concatenate([p.ravel() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from tests._pkg import load
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert np.array_equal(p0, p1), "replicas diverged without a weight broadcast"
    m = load()
    net = m.UNet3d(1, 2, FEATURE, gpu=0)
    net.init_params(7)
    net.set_dim(W, H, D); net.train(True); net.create_optimizer(0.01)
    for step in range(2):
        for b in range(2):
            x, t = _data(step * 2 + b)
            net.train_microbatch(x, t)
        net.step(2, 0.01)
    ref = np.concatenate([p.ravel() for p in net.parameters()])
    assert np.linalg.norm(p0 - ref) / np.linalg.norm(ref) < 1e-4


def _worker_uneven(rank, world, port, outdir, batch, declared):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from tests._pkg import load
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = load()
    comm = m.dist.bootstrap_nccl(m, dist, world, rank)
    net = m.UNet3d(1, 2, FEATURE, gpu=rank)
    net.init_params(7)
    net.set_dim(W, H, D); net.train(True); net.create_optimizer(0.01)
    net.attach_comm(comm, declared)
    for step in range(2):
        for b in m.dist.shard_microbatches(batch, world, rank):    # rank 1 runs fewer micro-batches than rank 0 -- or none at all
            x, t = _data(m.dist.sample_seed(step, batch, b))
            net.train_microbatch(x, t)
        net.step(batch, 0.01, comm)
        assert not net.last_step_skipped()
    np.save(os.path.join(outdir, f"q{rank}.npy"), np.concatenate([p.ravel() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("batch,declared", [(3, 2), (1, 1)])
def test_two_gpu_uneven_and_idle_ranks_issue_the_same_collectives(tmp_path, batch, declared):
    """batch 3 on 2 GPUs: rank 0 runs two micro-batches (its tail bucket is reduced during its last backward), rank 1 only one of
    the declared two (it reduces the tail in unet3d_step).  batch 1 on 2 GPUs: rank 1 runs none (the reference uses min(gpus, batch)
    workers, train.cpp:592) and joins with zero gradients.  Every rank must issue tail-then-prefix; the result equals one process."""
    import torch.multiprocessing as mp
    from tests._pkg import load
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_uneven, args=(2, port, str(tmp_path), batch, declared), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "q0.npy"), np.load(tmp_path / "q1.npy")
    assert np.array_equal(p0, p1), "replicas diverged"
    m = load()
    net = m.UNet3d(1, 2, FEATURE, gpu=0)
    net.init_params(7)
    net.set_dim(W, H, D); net.train(True); net.create_optimizer(0.01)
    for step in range(2):
        for b in range(batch):
            x, t = _data(step * batch + b)
            net.train_microbatch(x, t)
        net.step(batch, 0.01)
    ref = np.concatenate([p.ravel() for p in net.parameters()])
    assert np.linalg.norm(p0 - ref) / np.linalg.norm(ref) < 1e-4
