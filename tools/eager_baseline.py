"""The "kernel to beat" on the same box (SURVEY.md 2.1 / 8(d) last bullet): the reference's module graph under plain PyTorch
eager (cuDNN / cuBLAS) on the B200 -- DenseNet-121 f over every spot of one array + the 5-layer hexagonal corrector g + masked CE.

This is a BASELINE for bench.py's ``gpu_eager_baseline`` object, not a product path and not the oracle: it runs the product's
own module tree (whose children are the reference's nn.Conv2d / nn.BatchNorm2d / nn.Linear, /root/reference/gridnext/densenet.py:
93-159) with the reference's eager composition (torch.cat concatenation densenet.py:14,75; eval-mode f, training.py:126), and
hexagdly.Conv2d as the dense-kernel equivalent (two 3x3 convolutions, one per row parity).  Three channels: fp32 with TF32 off,
fp32 with TF32 on (PyTorch's cuDNN default), bf16 autocast + channels_last.  f is run in equal chunks of 192 spots (4,992 = 26 x 192: one set of cuDNN-autotuned shapes; fwd + bwd per chunk, no
recompute -- the reference itself needs ``atonce_patch_limit`` + checkpointing, i.e. an extra forward, to fit a whole array)."""
import time
import torch
import torch.nn.functional as F


def densenet_eager(net, x):
    f = net.features
    h = f.pool0(f.relu0(f.norm0(f.conv0(x)))) if hasattr(f, 'norm0') else f.conv0(x)
    for name, m in f.named_children():
        if name.startswith('denseblock'):
            feats = [h]
            for layer in m.children():
                t = torch.cat(feats, 1)
                t = layer.conv1(F.relu(layer.norm1(t)))
                t = layer.conv2(F.relu(layer.norm2(t)))
                feats.append(t)
            h = torch.cat(feats, 1)
        elif name.startswith('transition'):
            h = m.pool(m.conv(F.relu(m.norm(h))))
    h = F.relu(f.norm_final(h))
    h = F.adaptive_avg_pool2d(h, (1, 1)).flatten(1)
    return net.classifier(h) if net.classify else h


def hex_dense_kernels(conv):
    """hexagdly.Conv2d (kernel_size 1) in the Visium layout as two dense 3x3 kernels: even rows reach (y-1|y+1, x-1..x),
    odd rows (y-1|y+1, x..x+1); the same-row taps x-1, x, x+1 come from kernel0."""
    k0, k1 = conv.kernel0, conv.kernel1
    co, ci = k0.shape[:2]
    ke = k0.new_zeros(co, ci, 3, 3)
    ko = k0.new_zeros(co, ci, 3, 3)
    ke[:, :, 1, :] = k0[:, :, :, 0]
    ko[:, :, 1, :] = k0[:, :, :, 0]
    for a in (0, 1):
        ke[:, :, 0, a] = k1[:, :, a, 0]
        ke[:, :, 2, a] = k1[:, :, a, 1]
        ko[:, :, 0, a + 1] = k1[:, :, a, 0]
        ko[:, :, 2, a + 1] = k1[:, :, a, 1]
    return ke, ko


def corrector_eager(corrector, x):
    import torch.nn as nn
    odd = (torch.arange(x.shape[2], device=x.device) % 2 == 1).view(1, 1, -1, 1)
    h = x
    for m in corrector:
        if isinstance(m, (nn.BatchNorm2d, nn.ReLU)):
            h = m(h)
        else:
            ke, ko = hex_dense_kernels(m)
            h = torch.where(odd, F.conv2d(h, ko, m.bias_tensor, padding=1), F.conv2d(h, ke, m.bias_tensor, padding=1))
    return h


def masked_ce_eager(out, labels):
    o = out.permute(0, 2, 3, 1).reshape(-1, out.shape[1])
    l = labels.reshape(-1)
    return F.cross_entropy(o[l > 0], l[l > 0] - 1)


def run(model, patches, labels, chunk=192, channels=('fp32', 'tf32', 'bf16_autocast_channels_last')):
    """model: GridNetHexOddr with a gridnext_b200 DenseNet f (eval) on the device; patches (N, 3, P, P) bf16|fp32 device tensor,
    labels (1, H, W).  -> {channel: dict(ms_per_array, spots_per_s)}."""
    f, g = model.patch_classifier, model.corrector
    N = patches.shape[0]
    H, W = model.grid_shape
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for ch in channels:
            tf32 = ch != 'fp32'
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            amp = ch.startswith('bf16')

            def f_chunk(xc, dlog):
                xc = xc.float()
                if amp:
                    xc = xc.contiguous(memory_format=torch.channels_last)
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
                    o = densenet_eager(f, xc)
                o.float().backward(dlog)
                return o

            def one_array():
                dlog = torch.randn(chunk, f.classifier.out_features, device=patches.device)
                logits = []
                for i in range(0, N, chunk):
                    xc = patches[i:i + chunk]
                    logits.append(f_chunk(xc, dlog[:xc.shape[0]]).detach().float())
                fg = torch.cat(logits).reshape(1, H, W, -1).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
                loss = masked_ce_eager(corrector_eager(g, fg), labels)
                loss.backward()
                for p in model.parameters():
                    p.grad = None
            if amp:
                f.to(memory_format=torch.channels_last)
            f_chunk(patches[:chunk], torch.randn(min(chunk, N), f.classifier.out_features, device=patches.device))     # warm-up (cuDNN autotune)
            f_chunk(patches[:chunk], torch.randn(min(chunk, N), f.classifier.out_features, device=patches.device))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_array()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            res[ch] = dict(ms_per_array=ms, spots_per_s=N / (ms * 1e-3))
            if amp:
                f.to(memory_format=torch.contiguous_format)
            for p in model.parameters():
                p.grad = None
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return res
