"""world_size-2 gloo test of the data-parallel step protocol (CPU): sharding by b % world + all-reduce(sum) +
identical update on every rank == the reference's single-process step over the same micro-batches
(sum of per-sample gradients / batch_size, train.cpp:756-761).  The arithmetic here is the oracle's; what is under
test is the host-side sharding/seed logic shipped in unet-studio_b200/dist.py and the equivalence claim of DESIGN.md 5."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import unet3d_oracle as O

FEATURE = ("conv4,ks3,stride1+norm,leaky_relu\nconv8,ks3,stride2+norm,leaky_relu+conv_trans4,ks2,stride2\n"
           "conv4,ks3,stride1+norm,leaky_relu+conv2,ks1,stride1")
BATCH = 3


def _data(b):
    g = torch.Generator().manual_seed(100 + b)
    x = torch.rand(1, 1, 8, 8, 8, generator=g)
    t = (torch.rand(1, 8, 8, 8, generator=g) > 0.5).long()
    return x, t


def _worker(rank, world, port, out):
    import importlib.util, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("u3d_dist", os.path.join(root, "unet-studio_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = O.parse_feature(1, 2, FEATURE)
    P = O.init_params(net, 1)
    mom = [None] * len(P)
    for step in range(2):
        Pg = [p.detach().clone().requires_grad_(True) for p in P]
        mine = D.shard_microbatches(BATCH, world, rank)
        for b in mine:
            x, t = _data(D.sample_seed(step, BATCH, b))
            total, _, _ = O.micro_batch_loss(net, Pg, x, t)
            total.backward()
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in Pg])
        dist.all_reduce(flat)            # the one collective of the step
        grads, o = [], 0
        for p in P:
            grads.append(flat[o:o + p.numel()].view_as(p).clone()); o += p.numel()
        O.sgd_update(net, P, grads, mom, BATCH, 0.01)
    if rank == 0:
        np.save(out, torch.cat([p.reshape(-1) for p in P]).numpy())
    else:   # replicas must stay identical without any weight broadcast
        ref = [torch.zeros(sum(p.numel() for p in P))]
    flat = torch.cat([p.reshape(-1) for p in P])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "p.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    net = O.parse_feature(1, 2, FEATURE)
    P = O.init_params(net, 1)
    mom = [None] * len(P)
    for step in range(2):
        xs, ts = zip(*[_data(step * BATCH + b) for b in range(BATCH)])
        O.train_step(net, P, mom, xs, ts, 0.01)
    want = torch.cat([p.reshape(-1) for p in P]).numpy()
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-7)


def test_sharding_covers_every_microbatch_once():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("u3d_dist", os.path.join(root, "unet-studio_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
    for world in (1, 2, 4, 8):
        for batch in (1, 3, 8, 13):
            got = sorted(b for r in range(world) for b in D.shard_microbatches(batch, world, r))
            assert got == list(range(batch))


def _window_worker(rank, world, port, out):
    """Inference sharding: every rank forwards its own windows (oracle forward on CPU), rank 0 gathers; no data-path collective."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("u3d_dist", os.path.join(root, "unet-studio_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = O.parse_feature(1, 2, FEATURE)
    P = O.init_params(net, 3)
    mine = D.shard_windows(5, world, rank)
    with torch.no_grad():
        res = {i: O.forward(net, P, _data(200 + i)[0])[0] for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)          # the HOST gathers the outputs (evaluate.cpp:228-229 copies them into model_io)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        assert sorted(merged) == list(range(5))
        np.save(out, torch.stack([merged[i] for i in range(5)]).numpy())
    dist.destroy_process_group()


def test_window_sharding_equals_the_sequential_window_loop(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "w.npy")
    mp.spawn(_window_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    net = O.parse_feature(1, 2, FEATURE)
    P = O.init_params(net, 3)
    with torch.no_grad():
        want = torch.stack([O.forward(net, P, _data(200 + i)[0])[0] for i in range(5)]).numpy()
    np.testing.assert_array_equal(got, want)
