"""Index helpers and the evaluation loop of the reference's utils module (hot-path subset).

Reference: /root/reference/gridnext/utils.py:20-57 (all_fgd_predictions), :64-79 (index maps).
File parsing / plotting helpers of that module are out of scope.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import ptr, stream, call


def pseudo_hex_to_oddr(col, row):
    """Visium pseudo-hex (array_col, array_row) -> odd-right (x, y)."""
    if row % 2 == 0:
        x = col / 2
    else:
        x = (col - 1) / 2
    y = row
    return int(x), int(y)


def oddr_to_pseudo_hex(col, row):
    if row % 2 == 0:
        x = 2 * col
    else:
        x = 2 * col + 1
    y = row
    return int(x), int(y)


def pseudo_to_true_hex(col, row):
    return col / 2, row * np.sqrt(3) / 2


def fg_predictions(outputs, labels):
    """One batch of the evaluation loop on the device: (B, C, H, W) fp32 logits + (B, H, W) labels (0 = background) ->
    (true labels, predicted labels, softmax vectors) of the foreground spots in grid order, as device tensors.  The mask /
    gather / softmax / argmax chain of utils.py:44-52 is one kernel (csrc/corrector_ops.cu: gn_fg_predictions)."""
    _lib.require_cuda(outputs, labels)
    outputs = outputs.contiguous().float()
    labels = labels.contiguous().long()
    B, C, H, W = outputs.shape
    fg = (labels > 0).reshape(-1)
    incl = torch.cumsum(fg, 0, dtype=torch.int32)
    offsets = (incl - fg.to(torch.int32)).contiguous()
    n = B * H * W
    true_out = torch.empty(n, device=outputs.device, dtype=torch.int64)
    pred_out = torch.empty(n, device=outputs.device, dtype=torch.int64)
    smax_out = torch.empty((n, C), device=outputs.device, dtype=torch.float32)
    call('gn_fg_predictions', ptr(outputs), ptr(labels), ptr(offsets), ptr(true_out), ptr(pred_out), ptr(smax_out), B, C, H * W, stream())
    n_fg = int(incl[-1].item())            # the one host read of the batch (the arrays go to the host next anyway)
    return true_out[:n_fg], pred_out[:n_fg], smax_out[:n_fg]


def all_fgd_predictions(dataloader, model, f_only=False):
    """Flattened (true labels, predicted labels, softmax vectors) over all foreground spots."""
    true_vals, pred_vals, pred_smax = [], [], []
    device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    model.to(device)
    model.eval()
    for x, y in dataloader:
        x = [t.to(device) for t in x] if isinstance(x, (list, tuple)) else x.to(device)
        y = y.to(device)
        with torch.no_grad():
            outputs = model.patch_predictions(x) if f_only else model(x)
            if outputs.is_cuda:
                t, p, s = fg_predictions(outputs, y)
                true_vals.append(t.cpu().numpy())
                pred_vals.append(p.cpu().numpy())
                pred_smax.append(s.cpu().numpy())
                continue
            outputs = outputs.permute((0, 2, 3, 1))
            outputs = torch.reshape(outputs, (-1, outputs.shape[-1]))
            labels = torch.reshape(y, (-1,))
            outputs = outputs[labels > 0]
            labels = labels[labels > 0] - 1
            smax = F.softmax(outputs, dim=1)
            pred = torch.argmax(outputs, dim=1)
            true_vals.append(labels.cpu().numpy())
            pred_vals.append(pred.cpu().numpy())
            pred_smax.append(smax.cpu().numpy())
    return np.concatenate(true_vals), np.concatenate(pred_vals), np.concatenate(pred_smax)
