"""Training loops with the reference's signatures and return values.

Mirrors /root/reference/gridnext/training.py: ``train_spotwise`` (:11-98) and ``train_gridwise``
(:101-209).  Behaviour kept on purpose: only ``model.patch_classifier`` is put in
eval mode (so GridNetHexMM's count f keeps train-mode BatchNorm1d), no zero_grad before the first
backward, optimizer steps when ``batch_ind % accum_iters == 0``, best-val checkpointing to ``outfile``
and ``<outfile stem>.opt``.

What changes underneath ``train_gridwise``: the foreground mask + CrossEntropyLoss + argmax is one
fused kernel (``losses.masked_cross_entropy``) when the criterion is a plain ``nn.CrossEntropyLoss``;
loss/accuracy counters stay on the device and are read once per phase instead of once per batch; and,
when ``torch.distributed`` is initialised, gradients are summed over ranks in one flat NCCL all-reduce
per optimizer step (``parallel.GradBucket``) with rank 0 doing the printing and checkpointing.
"""
import os
import time
import copy
import numpy as np
import torch

from .losses import masked_cross_entropy, is_plain_cross_entropy, MAX_FUSED_CLASSES
from . import parallel


def _to_device(inputs, device):
    if isinstance(inputs, (list, tuple)):
        return [x.to(device, non_blocking=True) for x in inputs]
    return inputs.to(device, non_blocking=True)


def _device_batches(loader, device):
    """Iterate ``loader`` with the batches already on ``device``.  On CUDA the host-to-device copy of batch i+1 is issued on a copy
    stream before batch i is handed to the caller, so it runs under the compute of batch i (a 4-array batch of 128 px patches is
    2 GB: 36 ms of PCIe time per step that the reference's ``inputs.to(device)`` at the top of the loop body would serialise).
    Same batches, same order; pinned host tensors make the copy asynchronous, pageable ones still work."""
    if device.type != 'cuda':
        for inputs, labels in loader:
            yield _to_device(inputs, device), labels.to(device)
        return
    copy_stream = torch.cuda.Stream(device)
    it = iter(loader)

    def fetch():
        try:
            inputs, labels = next(it)
        except StopIteration:
            return None
        copy_stream.wait_stream(torch.cuda.current_stream(device))      # the allocator may hand back blocks the compute stream just freed
        with torch.cuda.stream(copy_stream):
            x, y = _to_device(inputs, device), labels.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return x, y, ev

    nxt = fetch()
    while nxt is not None:
        x, y, ev = nxt
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        for t in (list(x) if isinstance(x, (list, tuple)) else [x]) + [y]:
            if t.is_cuda:
                t.record_stream(cur)
        nxt = fetch()
        yield x, y


def _spot_forward(model, inputs):
    """f on one spot batch: the count MLP pattern goes through the tensor-core path (count_mlp.forward_spots), a
    gridnext_b200 DenseNet runs its own kernels in either BatchNorm mode, anything else is called as is."""
    if torch.is_tensor(inputs) and inputs.is_cuda and inputs.dim() == 2 and isinstance(model, torch.nn.Sequential):
        from .count_mlp import compile_count_mlp
        fast = compile_count_mlp(model)
        if fast is not None:
            out = fast.forward_spots(inputs)
            if out is not None:
                return out
    return model(inputs)


def train_spotwise(model, dataloaders, criterion, optimizer, num_epochs=10, outfile=None, display=False):
    """Spot classifier (f) pre-training loop (/root/reference/gridnext/training.py:11-98; row f3 of SURVEY.md section 8).

    Same contract as the reference: phases 'train' / 'val' per epoch, per-epoch mean LOSS appended to the train / val histories,
    the weights of the epoch with the lowest validation loss kept (and saved to ``outfile``) and loaded back at the end; returns
    ``(model, val_history, train_history)``.  f runs with train-mode BatchNorm in the train phase (batch statistics +
    running-stat update, csrc/bn_train.cu) and eval-mode BatchNorm in the val phase.  Running sums stay on the device: one host
    read per phase instead of one ``loss.item()`` per batch."""
    since = time.time()
    history = {'train': [], 'val': []}
    best_model_wts = copy.deepcopy(model.state_dict())
    best_loss = np.inf
    device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    model.to(device)
    for epoch in range(num_epochs):
        print('Epoch {}/{}'.format(epoch, num_epochs - 1), flush=True)
        print('-' * 10, flush=True)
        for phase in ('train', 'val'):
            training = phase == 'train'
            model.train(training)
            totals = torch.zeros(2, device=device, dtype=torch.float64)          # sum of loss * batch size, correct predictions
            batches = dataloaders[phase]
            if display:
                from tqdm import tqdm
                batches = tqdm(batches)
            for inputs, labels in batches:
                inputs = _to_device(inputs, device)
                labels = labels.to(device)
                optimizer.zero_grad()
                with torch.set_grad_enabled(training):
                    outputs = _spot_forward(model, inputs)
                    loss = criterion(outputs, labels)
                    if training:
                        loss.backward()
                        optimizer.step()
                totals[0] += loss.detach().double() * labels.size(0)
                totals[1] += (outputs.detach().argmax(1) == labels).sum()
            n = len(dataloaders[phase].dataset)
            epoch_loss, epoch_acc = (totals / n).tolist()
            print('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, epoch_loss, epoch_acc), flush=True)
            if not training and epoch_loss < best_loss:
                best_loss = epoch_loss
                best_model_wts = copy.deepcopy(model.state_dict())
                if outfile is not None:
                    torch.save(model.state_dict(), outfile)
            history[phase].append(epoch_loss)
        print()
    time_elapsed = time.time() - since
    print('Training complete in {:.0f}m {:.0f}s'.format(time_elapsed // 60, time_elapsed % 60), flush=True)
    print('Best val loss: {:4f}'.format(best_loss), flush=True)
    model.load_state_dict(best_model_wts)
    return model, history['val'], history['train']


def gridwise_step(model, inputs, labels, criterion, accum_iters=1, train=True, n_fg_override=None):
    """One forward (+ backward) of training.py:141-164.  Returns (loss tensor, acc fp64[4] | None, (n_correct, n_fg) | None)."""
    outputs = model(inputs)
    assert outputs.shape[2] == labels.shape[1] and outputs.shape[3] == labels.shape[2], \
        "Output tensor does not match label dimensions!"
    if is_plain_cross_entropy(criterion) and outputs.is_cuda and outputs.shape[1] <= MAX_FUSED_CLASSES:
        loss, acc = masked_cross_entropy(outputs, labels, accum_iters, n_fg_override)
        extra = None
    else:   # user-supplied criterion (or more classes than the fused kernel holds in registers): the reference's generic path
        o = outputs.permute((0, 2, 3, 1))
        o = torch.reshape(o, (-1, o.shape[-1]))
        l = torch.reshape(labels, (-1,))
        o = o[l > 0]
        l = l[l > 0] - 1
        loss = criterion(o, l) / accum_iters
        _, preds = torch.max(o, 1)
        acc, extra = None, (torch.sum(preds == l), len(l))
    if train:
        loss.backward()
    return loss, acc, extra


# ---- CUDA-graph replay of the training step (opt-in: GRIDNEXT_B200_GRAPH=1 or use_cuda_graphs(True)) ----------------------------
_USE_GRAPHS = [os.environ.get('GRIDNEXT_B200_GRAPH', '0') == '1']


def use_cuda_graphs(on=True):
    """Replay the grid-wise training step (forward, masked CE, backward, all-reduce, optimizer step) as ONE captured CUDA graph
    per batch shape.  The count-only configuration (BASELINE configs[0]) is ~60 short kernels per step and launch-bound when
    they are issued one by one from Python; a replay issues them back to back.  Needs ``accum_iters == 1``, a plain
    ``nn.CrossEntropyLoss`` and an optimizer that can be captured (``torch.optim.Adam/AdamW`` are switched to
    ``capturable=True`` when their state is still empty); anything else runs eagerly as before."""
    _USE_GRAPHS[0] = bool(on)


def _make_capturable(opt):
    if opt is None:
        return True
    if not isinstance(opt, (torch.optim.Adam, torch.optim.AdamW)):
        return isinstance(opt, torch.optim.SGD)
    if all(g.get('capturable', False) for g in opt.param_groups):
        return True
    if len(opt.state) == 0:
        for g in opt.param_groups:
            g['capturable'] = True
        return True
    return False


class _GraphedStep:
    """Static input buffers + one CUDA graph per batch signature.  The first batch of a signature runs eagerly (lazy optimizer
    state, workspace plans); the second is captured; later ones copy into the static buffers and replay."""

    def __init__(self, fn):
        self.fn, self.seen, self.graphs = fn, {}, {}

    @staticmethod
    def _sig(inputs, labels):
        xs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        return tuple((tuple(x.shape), x.dtype) for x in xs) + ((tuple(labels.shape), labels.dtype),)

    def __call__(self, inputs, labels):
        sig = self._sig(inputs, labels)
        entry = self.graphs.get(sig)
        if entry is None:
            if sig not in self.seen:
                self.seen[sig] = True
                return self.fn(inputs, labels)
            xs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
            static_x = [torch.empty_like(x) for x in xs]
            static_y = torch.empty_like(labels)
            for a, b in zip(static_x, xs):
                a.copy_(b)
            static_y.copy_(labels)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            arg = static_x if isinstance(inputs, (list, tuple)) else static_x[0]
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                self.fn(arg, static_y)
            self.graphs[sig] = (g, static_x, static_y)
            g.replay()
            return None
        g, static_x, static_y = entry
        xs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        for a, b in zip(static_x, xs):
            a.copy_(b, non_blocking=True)
        static_y.copy_(labels, non_blocking=True)
        g.replay()
        return None

    def release(self):
        self.graphs.clear()


def train_gridwise(model, dataloaders, criterion, optimizer, num_epochs=10, outfile=None,
                   f_opt=None, accum_iters=1):
    """Grid-wise training loop (/root/reference/gridnext/training.py:101-209), same signature and return triple.

    Data parallel (``torch.distributed`` initialised, one process per GPU): every rank iterates ITS shard of the arrays
    (``parallel.shard_indices`` / a DistributedSampler); gradients are summed in one flat all-reduce per optimizer step.  The
    ranks must see the same number of batches per phase (checked up front; a mismatch raises instead of hanging in NCCL).
    ``parallel.configure(loss_norm='global')`` divides every rank's loss sum by the GLOBAL foreground count (= the single-process
    result on the concatenated batch) instead of averaging per-rank means; ``parallel.configure(sync_bn=True)``
    makes the corrector's BatchNorm2d layers use statistics over all ranks' cells."""
    since = time.time()
    train_history, val_history = [], []
    best_model_wts = copy.deepcopy(model.state_dict())
    best_loss = np.inf

    dist_on = parallel.is_distributed()
    rank0 = parallel.rank() == 0
    device = torch.device("cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() else "cpu")
    model.to(device)
    bucket = parallel.GradBucket([p for p in model.parameters() if p.requires_grad]) if dist_on else None
    global_norm = dist_on and parallel.config().loss_norm == 'global'
    fused_ce = is_plain_cross_entropy(criterion) and device.type == 'cuda'

    def say(*a):
        if rank0:
            print(*a, flush=True)

    run = torch.zeros(4, device=device, dtype=torch.float64)   # loss*batch, corrects, foreground, labels out of range

    def one_batch(inputs, labels, phase, step_now):
        batch_size = labels.size(0)
        nfg = None
        if global_norm and fused_ce:
            nfg = (labels > 0).sum().double().reshape(1)
            parallel.allreduce_sum_(nfg)
        with torch.set_grad_enabled(phase == 'train'):
            loss, acc, extra = gridwise_step(model, inputs, labels, criterion, accum_iters, phase == 'train', nfg)
            if phase == 'train' and step_now:
                if bucket is not None:
                    if global_norm and fused_ce:
                        bucket.allreduce_sum()          # every rank's loss is already divided by the global count
                    else:
                        bucket.allreduce_mean()
                # with a gradient bucket the .grad tensors are views of it: zero them in place so that the next backward
                # accumulates straight into the bucket (no per-parameter copy before the all-reduce)
                optimizer.step()
                optimizer.zero_grad(set_to_none=bucket is None)
                if f_opt is not None:
                    f_opt.step()
                    f_opt.zero_grad(set_to_none=bucket is None)
        run[0] += loss.detach().double() * batch_size
        if acc is not None:
            run[1] += acc[2]
            run[2] += acc[1]
            run[3] += acc[3]
        else:
            run[1] += extra[0]
            run[2] += extra[1]

    graphed = None
    if (_USE_GRAPHS[0] and device.type == 'cuda' and accum_iters == 1 and fused_ce
            and _make_capturable(optimizer) and _make_capturable(f_opt)):
        graphed = _GraphedStep(lambda x, y: one_batch(x, y, 'train', True))

    try:
        for epoch in range(num_epochs):
            say('Epoch {}/{}'.format(epoch, num_epochs - 1))
            say('-' * 10)
            for phase in ['train', 'val']:
                model.train() if phase == 'train' else model.eval()
                model.patch_classifier.eval()      # training.py:126 -- f's BN/dropout frozen
                if dist_on:
                    parallel.check_equal_across_ranks(len(dataloaders[phase]), "number of %s batches" % phase)
                run.zero_()
                n_seen = 0
                for batch_ind, (inputs, labels) in enumerate(_device_batches(dataloaders[phase], device)):
                    n_seen += labels.size(0)
                    if graphed is not None and phase == 'train':
                        graphed(inputs, labels)
                    else:
                        one_batch(inputs, labels, phase, batch_ind % accum_iters == 0)
                n_total = len(dataloaders[phase].dataset)
                if dist_on:
                    cnt = torch.tensor([float(n_seen)], device=device, dtype=torch.float64)
                    tot = run.clone()
                    parallel.allreduce_sum_(tot)
                    parallel.allreduce_sum_(cnt)
                    n_total = int(cnt.item())
                    r = tot.tolist()
                    if global_norm and fused_ce:
                        # each rank's loss is (its loss sum / global fg count): the batch loss is their sum over ranks
                        n_total = max(n_total // parallel.world_size(), 1)
                else:
                    r = run.tolist()                       # the one host sync of the phase
                if r[3] > 0:
                    raise IndexError("Target out of bounds: %d labels exceed the %d classes of the model output" % (int(r[3]), model.n_classes))
                epoch_loss = r[0] / max(n_total, 1)
                epoch_acc = r[1] / r[2] if r[2] > 0 else float('nan')
                say('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, epoch_loss, epoch_acc))

                if phase == 'val' and epoch_loss < best_loss:
                    best_loss = epoch_loss
                    best_model_wts = copy.deepcopy(model.state_dict())
                    if outfile is not None and rank0:
                        torch.save(model.state_dict(), outfile)
                        if f_opt is not None:
                            torch.save({'g_opt': optimizer.state_dict(), 'f_opt': f_opt.state_dict()},
                                       os.path.splitext(outfile)[0] + ".opt")
                        else:
                            torch.save(optimizer.state_dict(), os.path.splitext(outfile)[0] + ".opt")
                (val_history if phase == 'val' else train_history).append(epoch_loss)
            say()
    finally:
        if graphed is not None:
            torch.cuda.synchronize()
            graphed.release()          # graphs (and the NCCL work captured in them) go before the process group does
        if bucket is not None:
            bucket.close()

    time_elapsed = time.time() - since
    say('Training complete in {:.0f}m {:.0f}s'.format(time_elapsed // 60, time_elapsed % 60))
    say('Best val loss: {:4f}'.format(best_loss))
    model.load_state_dict(best_model_wts)
    return model, val_history, train_history
