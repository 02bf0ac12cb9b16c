"""Host logic of the data-parallel step on CPU: world_size 2 over gloo.

Two ranks own disjoint arrays; each computes gradients of the same tiny model on its shard; after ONE all-reduce of the
flat bucket both ranks hold the mean gradient, which equals the single-process gradient of the mean loss over shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from gridnext_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _model():
    torch.manual_seed(3)
    return nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))


def _data(n_arrays=5):
    g = torch.Generator(); g.manual_seed(9)
    return torch.randn(n_arrays, 4, 6, generator=g), torch.randn(n_arrays, 4, 3, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert parallel.is_distributed() and parallel.rank() == rank and parallel.world_size() == world
        model = _model()
        bucket = parallel.GradBucket(model.parameters())
        x, y = _data()
        mine = parallel.shard_indices(x.shape[0])
        assert mine == list(range(rank, x.shape[0], world))
        # two accumulation steps: gradients accumulate IN the bucket (p.grad are views), one all-reduce at the end
        for i in mine:
            ((model(x[i]) - y[i]) ** 2).sum().backward()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        bucket.allreduce_sum()
        n_fg = parallel.allreduce_sum_(torch.tensor([float(len(mine))]))
        out.put((rank, bucket.flat.clone(), float(n_fg)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _model()
    x, y = _data()
    for i in range(x.shape[0]):
        ((model(x[i]) - y[i]) ** 2).sum().backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    for rank, flat, n in res:
        assert n == x.shape[0]
        assert torch.allclose(flat, ref, rtol=1e-5, atol=1e-6), rank


def test_shard_indices_cover_every_array_once():
    for n in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in parallel.shard_indices(n, r, world))
            assert got == list(range(n))


def test_bucket_single_process_is_a_no_op_collective():
    model = _model()
    b = parallel.GradBucket(model.parameters())
    x, y = _data(1)
    ((model(x[0]) - y[0]) ** 2).sum().backward()
    before = b.flat.clone()
    b.allreduce_mean()
    assert torch.equal(before, b.flat) and not parallel.is_distributed()
