"""Data parallelism over Visium arrays: one process per GPU, one flat gradient all-reduce per step.

The reference is single-device (/root/reference/gridnext/training.py:21,111).  Arrays are independent
given the weights, so ranks take disjoint arrays (``shard_indices``) and the only exchange is the sum of
parameter gradients before ``optimizer.step()``: a single NCCL all-reduce over ONE persistent flat fp32
bucket (<= 38 MB for DenseNet-121 + MLP + corrector), whose slices ARE the ``.grad`` tensors, so there
is no pack/unpack copy.  Works with the ``gloo`` backend on CPU for the host-logic tests.
"""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def shard_indices(n_items, rank_=None, world=None):
    """Arrays r, r+R, r+2R, ... for rank r."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    return list(range(rank_, n_items, world))


def allreduce_sum_(t):
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class GradBucket:
    """Flat fp32 gradient storage shared by all parameters; ``p.grad`` are views into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('GradBucket: no trainable parameters')
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()
        self.attach()

    def attach(self):
        """(Re)point every ``p.grad`` at its bucket slice, keeping any gradient already accumulated."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            else:
                continue
            p.grad = v

    def zero_(self):
        self.flat.zero_()

    def allreduce_sum(self):
        self.attach()
        allreduce_sum_(self.flat)

    def allreduce_mean(self):
        self.attach()
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world_size())
