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
        # programmatic dependent launch must be off in a process group with more than one rank (the 2-GPU step hung with NCCL's
        # all-reduce between kernels that trigger their dependents early): the first C-ABI call after init_process_group decides it
        from gridnext_b200 import _lib
        lib = _lib.load()
        assert lib.gn_set_pdl(1) in (0, 1)
        _lib._PDL_DECIDED[0] = False
        _lib._decide_pdl(lib)
        assert _lib._PDL_DECIDED[0] and lib.gn_set_pdl(0) == 0
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


# ---- round 2: rank-count agreement, parameters without gradients, train_gridwise under two ranks ---------------------------
class _ToyGrid(nn.Module):
    """A GridNet-shaped toy: patch_classifier attribute, (B, C, H, W) -> (B, C, H, W), one parameter that never gets a gradient."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.patch_classifier = nn.Identity()
        self.n_classes = 3
        self.mix = nn.Conv2d(3, 3, 1)
        self.unused = nn.Parameter(torch.ones(2))

    def forward(self, x):
        return self.mix(x)


def _toy_data(n):
    g = torch.Generator(); g.manual_seed(11)
    return torch.utils.data.TensorDataset(torch.randn(n, 3, 4, 5, generator=g), torch.randint(0, 4, (n, 4, 5), generator=g))


def _train_worker(rank, world, port, out, n_arrays):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gridnext_b200.training import train_gridwise
        ds = _toy_data(n_arrays)
        mine = torch.utils.data.Subset(ds, parallel.shard_indices(len(ds)))
        dls = {'train': torch.utils.data.DataLoader(mine, batch_size=1), 'val': torch.utils.data.DataLoader(mine, batch_size=1)}
        model = _ToyGrid()
        opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=0.1)
        try:
            model, vh, th = train_gridwise(model, dls, nn.CrossEntropyLoss(), opt, num_epochs=2)
            out.put((rank, 'ok', [p.detach().numpy().copy() for p in model.parameters()], vh, th))
        except RuntimeError as exc:
            out.put((rank, 'error', str(exc), None, None))
    finally:
        dist.destroy_process_group()


def _spawn(target, world, *args):
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_train_gridwise_two_ranks_same_weights_and_untouched_dead_parameter():
    """Both ranks end with identical weights (one all-reduce per step), histories agree, and the parameter that receives no
    gradient on any rank is NOT stepped (Adam with weight decay would move it if the bucket handed it a zero gradient)."""
    res = _spawn(_train_worker, 2, 4)
    assert all(r[1] == 'ok' for r in res), res
    (_, _, p0, vh0, th0), (_, _, p1, vh1, th1) = res
    import numpy as np
    for a, b in zip(p0, p1):
        assert np.array_equal(a, b)
    assert vh0 == vh1 and th0 == th1 and len(vh0) == 2
    ref = _ToyGrid()
    names = [n for n, _ in ref.named_parameters()]
    got = dict(zip(names, p0))
    assert np.array_equal(got['unused'], ref.unused.detach().numpy())              # untouched
    assert not np.array_equal(got['mix.weight'], ref.mix.weight.detach().numpy())   # trained


def test_train_gridwise_unequal_batch_counts_fail_loudly():
    """5 arrays over 2 ranks = 3 vs 2 batches: every rank raises before the first collective instead of hanging."""
    res = _spawn(_train_worker, 2, 5)
    assert all(r[1] == 'error' and 'disagree' in r[2] for r in res), res
