"""Index helpers and the evaluation loop of the reference's utils module (hot-path subset).

Reference: /root/reference/gridnext/utils.py:20-57 (all_fgd_predictions), :64-79 (index maps).
File parsing / plotting helpers of that module are out of scope.
"""
import numpy as np
import torch
import torch.nn.functional as F


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
