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


class _Config:
    """Data-parallel semantics that have no counterpart in the single-device reference.

    loss_norm  'local'  : every rank normalises its loss by ITS foreground count and gradients are averaged over ranks
                          (mean of per-rank means; equals the reference when every rank is looked at alone).
               'global' : every rank divides its loss sum by the foreground count of ALL ranks (one extra scalar all-reduce)
                          and gradients are summed: exactly the single-process result on the concatenated batch.
    sync_bn    False    : the corrector's BatchNorm2d uses per-replica batch statistics (= the reference at that per-GPU batch).
               True     : statistics (forward) and the two gradient sums (backward) are all-reduced: the single-process
                          global-batch result."""
    loss_norm = 'local'
    sync_bn = False


_CFG = _Config()


def config():
    return _CFG


def configure(loss_norm=None, sync_bn=None):
    if loss_norm is not None:
        if loss_norm not in ('local', 'global'):
            raise ValueError("loss_norm must be 'local' or 'global'")
        _CFG.loss_norm = loss_norm
    if sync_bn is not None:
        _CFG.sync_bn = bool(sync_bn)
    return _CFG


def sync_bn_active():
    return _CFG.sync_bn and is_distributed()


def check_equal_across_ranks(value, what):
    """Raise on every rank if ``value`` (an int) differs between ranks -- e.g. the number of batches a phase will run, which
    must agree because each batch issues a collective (a rank that finished early would pair its next all-reduce with another
    rank's gradient all-reduce and hang or corrupt)."""
    if not is_distributed():
        return
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor([float(value), -float(value)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hi, lo = t[0].item(), -t[1].item()
    if hi != lo:
        raise RuntimeError("data-parallel ranks disagree on the %s (min %d, max %d; this rank %d): shard the arrays so that every "
                           "rank runs the same number of batches (pad or drop the remainder)" % (what, int(lo), int(hi), int(value)))


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
        self.live = None          # per parameter: did ANY rank produce a gradient for it (decided at the first all-reduce)
        self._touched = [False] * len(self.params)
        self._hooks = [p.register_post_accumulate_grad_hook(lambda _p, i=i: self._touched.__setitem__(i, True))
                       for i, p in enumerate(self.params)]
        self.attach()

    def close(self):
        """Detach from the parameters: remove the hooks and give every parameter a gradient tensor of its own again."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() == v.data_ptr():
                p.grad = v.clone()

    def attach(self):
        """(Re)point every ``p.grad`` at its bucket slice, keeping any gradient already accumulated.  A parameter without a
        gradient contributes zeros to the sum."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            else:
                continue
            p.grad = v

    def _drop_dead(self):
        """Parameters that received no gradient on ANY rank keep ``grad = None`` like in a single-process run (Adam / weight
        decay would otherwise start moving them).  The set is static for a model, so it is decided once (one small
        all-reduce + host read at the first step) and reused."""
        if self.live is None:
            flags = torch.tensor([1.0 if h else 0.0 for h in self._touched], device=self.flat.device)
            allreduce_sum_(flags)
            self.live = [f > 0 for f in flags.tolist()]
        for p, alive in zip(self.params, self.live):
            if not alive:
                p.grad = None

    def zero_(self):
        self.flat.zero_()

    def allreduce_sum(self):
        self.attach()
        allreduce_sum_(self.flat)
        self._drop_dead()

    def allreduce_mean(self):
        self.attach()
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world_size())
        self._drop_dead()
