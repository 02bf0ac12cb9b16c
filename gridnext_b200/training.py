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

from .losses import masked_cross_entropy, is_plain_cross_entropy
from . import parallel


def _to_device(inputs, device):
    if isinstance(inputs, (list, tuple)):
        return [x.to(device, non_blocking=True) for x in inputs]
    return inputs.to(device, non_blocking=True)


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
    if is_plain_cross_entropy(criterion) and outputs.is_cuda:
        loss, acc = masked_cross_entropy(outputs, labels, accum_iters, n_fg_override)
        extra = None
    else:   # user-supplied criterion: the reference's generic path
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


def train_gridwise(model, dataloaders, criterion, optimizer, num_epochs=10, outfile=None,
                   f_opt=None, accum_iters=1):
    since = time.time()
    train_history, val_history = [], []
    best_model_wts = copy.deepcopy(model.state_dict())
    best_loss = np.inf

    dist_on = parallel.is_distributed()
    rank0 = parallel.rank() == 0
    device = torch.device("cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() else "cpu")
    model.to(device)
    bucket = parallel.GradBucket([p for p in model.parameters() if p.requires_grad]) if dist_on else None

    def say(*a):
        if rank0:
            print(*a, flush=True)

    for epoch in range(num_epochs):
        say('Epoch {}/{}'.format(epoch, num_epochs - 1))
        say('-' * 10)
        for phase in ['train', 'val']:
            model.train() if phase == 'train' else model.eval()
            model.patch_classifier.eval()      # training.py:126 -- f's BN/dropout frozen

            run = torch.zeros(3, device=device, dtype=torch.float64)   # loss*batch, corrects, foreground
            n_seen = 0
            for batch_ind, (inputs, labels) in enumerate(dataloaders[phase]):
                batch_size = labels.size(0)
                n_seen += batch_size
                inputs = _to_device(inputs, device)
                labels = labels.to(device, non_blocking=True)
                with torch.set_grad_enabled(phase == 'train'):
                    loss, acc, extra = gridwise_step(model, inputs, labels, criterion, accum_iters, phase == 'train')
                    if phase == 'train' and batch_ind % accum_iters == 0:
                        if bucket is not None:
                            bucket.allreduce_mean()
                        optimizer.step()
                        optimizer.zero_grad()
                        if f_opt is not None:
                            f_opt.step()
                            f_opt.zero_grad()
                run[0] += loss.detach().double() * batch_size
                if acc is not None:
                    run[1] += acc[2]
                    run[2] += acc[1]
                else:
                    run[1] += extra[0]
                    run[2] += extra[1]
            n_total = len(dataloaders[phase].dataset)
            if dist_on:
                cnt = torch.tensor([float(n_seen)], device=device, dtype=torch.float64)
                parallel.allreduce_sum_(run)
                parallel.allreduce_sum_(cnt)
                n_total = int(cnt.item())
            r = run.tolist()                       # the one host sync of the phase
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

    time_elapsed = time.time() - since
    say('Training complete in {:.0f}m {:.0f}s'.format(time_elapsed // 60, time_elapsed % 60))
    say('Best val loss: {:4f}'.format(best_loss))
    model.load_state_dict(best_model_wts)
    return model, val_history, train_history
