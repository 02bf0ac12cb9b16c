#!/usr/bin/env python
"""Benchmark of the GridNet hot path on B200:

    python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5] [--impl reference]

Metric (BASELINE.json): Visium spots/sec for one f+g training step (forward, foreground-masked CE, backward, gradient
all-reduce when N > 1, Adam step) on synthetic 78 x 64 Visium arrays; weak scaling (each rank owns its arrays, the only
exchange is the flat gradient all-reduce).  Configurations (BASELINE.json ``configs``; the default is the one the metric is
quoted on):

  c2 (default)  GridNetHex image-only: DenseNet-121 f on 3x128x128 patches for all 4,992 spots + hex g, 1 array / GPU / step
  c1            GridNetHex count-only: MLP f over 5,000 genes + hex g, 1 array / GPU / step (the CPU-runnable case)
  c3            Multimodal GridNetHex: count MLP f + DenseNet-121 f, hex g, 12 arrays / GPU / step (chunked recompute path)
  c4            g only: the 5-layer hex corrector + masked CE at 256 arrays (n_classes 7); --profile-out gets the whole sweep
  c5            train_gridwise (the product API, DataLoader-fed, pinned host batches) over 4-array batches per GPU per step

Legs of the B200 arm:
  value        inputs resident in HBM when the timed region starts; CUDA events, max over ranks; the step is a fixed launch
               sequence replayed from a CUDA graph (``--no-graph`` times the eager loop).
  e2e          the same step fed from HOST memory: the step's inputs are copied H2D from pinned memory every step and the
               scalar loss is read back (D2H) every step.
  roofline     one instrumented step (CUDA events around every C-ABI call) -> the dominant kernel: tensor-pipe fraction for the
               DenseNet / MLP kernels, HBM fraction for the g / gather kernels; ``traffic`` = DRAM bytes per
               launch from the committed ncu capture (profiles/r02_traffic.json) when it has the kernel.
  cpu_baseline the CPU oracle (restated reference modules) timed on this box's host cores on a bounded sample.
  gpu_eager_baseline (c2, rank 0) the reference's module graph under plain PyTorch eager (cuDNN/cuBLAS) on the same B200:
               fp32, TF32, bf16-autocast channels_last (tools/eager_baseline.py) -- the "kernel to beat" of SURVEY.md 2.1.
  --impl reference   times only the CPU arm, K steps after W warm-ups, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_CLS, P, H_ST, W_ST = 7, 128, 78, 64
SPOTS = H_ST * W_ST
G_GENES = 5000
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
DENSENET_KW = dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4)
# algorithmic work per spot: DenseNet-121 @128 px forward+backward 5.47 GFLOP; count MLP forward+backward 10.39 MFLOP
F_FLOP_FWD_BWD = 5.47e9
MLP_FLOP_FWD_BWD = 10.39e6
G_BYTES_PER_CELL = 1080.0 + 1620.0 + 16.0 * N_CLS + 8      # 5-layer corrector fwd + bwd + masked CE
METRIC = 'visium_spots_per_sec_f+g_fwd+bwd'

WORKLOADS = {
    'c1': 'GridNetHex count-only: MLP f over 5k synthetic genes + 5-layer hex g on one 78x64 Visium array (BASELINE configs[0])',
    'c2': 'GridNetHex image-only: DenseNet-121 f on 3x128x128 patches for all 4,992 spots + 5-layer hex g (BASELINE configs[1])',
    'c3': 'Multimodal GridNetHex: count MLP f (5k genes) + DenseNet-121 f (3x128x128), hex g, 12-array batch (BASELINE configs[2])',
    'c4': 'g only: 5-layer hex corrector (n_classes 7, k=1) + masked CE forward+backward at 256 arrays (BASELINE configs[3])',
    'c5': 'train_gridwise epoch slice, image GridNetHex (DenseNet-121 @128), DataLoader-fed 4-array batches per GPU (BASELINE configs[4])',
}
ARRAYS_PER_STEP = {'c1': 1, 'c2': 1, 'c3': 12, 'c4': 256, 'c5': 4}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d['hbm_gbs'], bf16=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured (MEASURED_PEAKS.json: copy bandwidth, sustained bf16)')
    return dict(hbm=6650.0, bf16=1400.0, src='fallback (B200_PROFILING.md)')


def committed_traffic():
    """DRAM bytes per launch of the step's kernels from the committed ncu capture of this same command (profiles/)."""
    path = os.path.join(ROOT, 'profiles', 'r02s_traffic.json')
    try:
        return json.load(open(path))
    except Exception:
        return {}


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return None
        sm = sorted(int(r[0]) for r in rows)
        reasons = []
        for name, col in (('hw_slowdown', 2), ('hw_thermal_slowdown', 3), ('sw_thermal_slowdown', 4), ('sw_power_cap', 5)):
            if any(r[col].lower().startswith('active') for r in rows):
                reasons.append(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=int(rows[0][1]), reasons=reasons, samples=len(rows))


def workload_config(cfg, n_gpus):
    d = dict(workload=WORKLOADS[cfg], config=cfg, arrays_per_gpu_per_step=ARRAYS_PER_STEP[cfg], spots_per_array=SPOTS, n_classes=N_CLS,
             parallelism='dp%d' % n_gpus, step='f fwd/bwd + g fwd/bwd + masked CE + grad all-reduce + Adam')
    if cfg in ('c2', 'c3', 'c5'):
        d['patch'] = P
        d['l2'] = 'inputs and activations (>20 GB per array) far exceed the 126 MB L2'
    if cfg in ('c1', 'c3'):
        d['genes'] = G_GENES
    if cfg == 'c1':
        d['l2'] = 'the step rotates over 4 different resident input arrays (4 x 100 MB > 126 MB L2)'
    if cfg == 'c4':
        d['step'] = 'g fwd/bwd + masked CE (no f, no optimizer)'
        d['l2'] = 'activations of 256 arrays (1.3 M cells x 32 channels fp32 per layer = 164 MB) exceed the 126 MB L2'
    if cfg == 'c2':
        d['step'] = 'patch gather + ' + d['step']
    return d


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_reference_step_fn(cfg, sample_spots=128):
    """Returns (fn, meta): fn() -> seconds per step of `cfg` on the host cores (f on a bounded sample, scaled; g on full arrays)."""
    from oracle import synth, shapes as S, gridnet_ref as R
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator(); g.manual_seed(7)
    arrays = ARRAYS_PER_STEP[cfg]

    def grads_off(sd):
        for v in sd.values():
            if v.is_floating_point() and v.grad is not None:
                v.grad = None

    def leaf(sd):
        for k, v in sd.items():
            if v.is_floating_point() and 'running' not in k and k not in ('bg_const', 'dummy_tensor'):
                v.requires_grad_(True)
        return sd

    labels = synth.synth_labels(1, N_CLS, seed=0)
    parts = []
    if cfg in ('c2', 'c3', 'c5'):
        sd_i = leaf(synth.synth_state_dict(S.densenet_shapes(**DENSENET_KW), 1234))
        x = torch.randn(sample_spots, 3, P, P, generator=g)
        dlog = torch.randn(sample_spots, N_CLS, generator=g)

        def f_image():
            t0 = time.perf_counter()
            (R.densenet_forward(sd_i, x) * dlog).sum().backward()
            grads_off(sd_i)
            return (time.perf_counter() - t0) * SPOTS / sample_spots          # f is per-spot independent in eval mode
        parts.append(f_image)
    if cfg in ('c1', 'c3'):
        sd_c = leaf(synth.synth_state_dict(S.mlp_shapes(G_GENES, N_CLS), 1234))
        xc = synth.synth_counts(1, G_GENES, seed=1)
        dlc = torch.randn(SPOTS, N_CLS, generator=g)

        def f_count():
            t0 = time.perf_counter()
            (R.mlp_forward(sd_c, R.spots_from_counts(xc), training=(cfg == 'c3')) * dlc).sum().backward()
            grads_off(sd_c)
            return time.perf_counter() - t0
        parts.append(f_count)
    f_dim = 2 * N_CLS if cfg == 'c3' else N_CLS
    sd_g = leaf(synth.synth_state_dict(S.corrector_shapes(f_dim, N_CLS), 1234))
    g_arrays = min(arrays, 4)
    fgrid = torch.randn(g_arrays, f_dim, H_ST, W_ST, generator=g)
    glabels = synth.synth_labels(g_arrays, N_CLS, seed=0)

    def g_part():
        t0 = time.perf_counter()
        fg = fgrid.clone().requires_grad_(True)
        out = R.corrector_forward(sd_g, fg, use_bn=True, training=True)
        loss, _, _ = R.masked_ce(out, glabels)
        loss.backward()
        grads_off(sd_g)
        return (time.perf_counter() - t0) / g_arrays

    def fn():
        return arrays * (sum(p() for p in parts) + g_part())
    what = {'c1': 'count MLP f (5,000 genes, fp32) fwd+bwd on one full array',
            'c2': 'f (DenseNet-121 fwd+bwd, fp32) on %d of 4992 spots scaled linearly' % sample_spots,
            'c3': 'image f on %d of 4992 spots scaled linearly + count f on one full array, both x 12 arrays' % sample_spots,
            'c4': 'no f', 'c5': 'f (DenseNet-121 fwd+bwd, fp32) on %d of 4992 spots scaled to 4 arrays' % sample_spots}[cfg]
    meta = dict(cores=os.cpu_count(), kind='port',
                sample=what + ' + g (5 hex convs, BN, masked CE fwd+bwd) on %d full array(s) scaled to %d' % (g_arrays, arrays))
    return fn, meta


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    fn, meta = cpu_reference_step_fn(args.config)
    for _ in range(args.warmup):
        fn()
    ts = [fn() for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    val = SPOTS * ARRAYS_PER_STEP[args.config] / t
    line = dict(metric=METRIC, value=val, unit='spots/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=workload_config(args.config, args.gpus), cpu_baseline=dict(value=val, unit='spots/s', **meta),
                e2e=dict(value=val, unit='spots/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm: workloads
def tutorial_mlp(G, n_cls):
    import torch.nn as nn
    return nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(),
                         nn.Linear(100, 100), nn.Linear(100, 50), nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, n_cls))


class Workload:
    """One configuration: model, optimizer, resident inputs, a host-fed variant, and how to count its work."""

    def __init__(self, cfg, dev, rank, world, fused_adam=True):
        import torch.nn as nn
        from gridnext_b200 import imgprocess as ip, parallel
        from gridnext_b200.densenet import DenseNet
        from gridnext_b200.gridnet_models import GridNetHexOddr, GridNetHexMM
        from synthdata import synth, shapes as S          # seeded synthetic weights / inputs (nothing from oracle/ on this arm)
        self.cfg, self.dev, self.rank, self.world = cfg, dev, rank, world
        self.arrays = ARRAYS_PER_STEP[cfg]
        self.ip = ip
        B = self.arrays
        if cfg in ('c2', 'c5'):
            f = DenseNet(num_classes=N_CLS, small_inputs=False, efficient=False, drop_rate=0, **DENSENET_KW)
            model = GridNetHexOddr(f, (3, P, P), (H_ST, W_ST), N_CLS, use_bn=True, atonce_patch_limit=None)
            model.load_state_dict(synth.synth_state_dict(S.gridnet_shapes(S.densenet_shapes(**DENSENET_KW), N_CLS, N_CLS), 1234))
        elif cfg == 'c1':
            model = GridNetHexOddr(tutorial_mlp(G_GENES, N_CLS), (G_GENES,), (H_ST, W_ST), N_CLS, use_bn=True)
            model.load_state_dict(synth.synth_state_dict(S.gridnet_shapes(S.mlp_shapes(G_GENES, N_CLS), N_CLS, N_CLS), 1234))
        elif cfg == 'c3':
            fi = DenseNet(num_classes=N_CLS, small_inputs=False, efficient=False, drop_rate=0, **DENSENET_KW)
            model = GridNetHexMM(fi, tutorial_mlp(G_GENES, N_CLS), (3, P, P), (G_GENES,), (H_ST, W_ST), N_CLS)
            sd = synth.synth_state_dict(S.gridnet_mm_shapes(S.densenet_shapes(**DENSENET_KW), S.mlp_shapes(G_GENES, N_CLS), N_CLS, N_CLS, N_CLS), 1234)
            for k in list(sd):
                if k.startswith('patch_classifier.'):
                    sd[k] = sd['image_classifier.' + k[len('patch_classifier.'):]]
            model.load_state_dict(sd)
        elif cfg == 'c4':
            model = GridNetHexOddr(nn.Identity(), (N_CLS,), (H_ST, W_ST), N_CLS, use_bn=True)
            model.load_state_dict(synth.synth_state_dict(S.gridnet_shapes({}, N_CLS, N_CLS), 1234))
        model.to(dev)
        model.train(); model.patch_classifier.eval()          # training.py:120-126
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = None
        if cfg != 'c4':
            self.opt = torch.optim.Adam(self.params, lr=1e-4, capturable=True, fused=True if fused_adam else None)
        self.bucket = parallel.GradBucket(self.params) if (world > 1 and cfg not in ('c4', 'c5')) else None    # c5: train_gridwise owns its bucket
        self.crit = nn.CrossEntropyLoss()
        gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
        self.labels = synth.synth_labels(B, N_CLS, seed=rank).to(dev)
        self.host_inputs, self.dev_inputs = [], []
        if cfg == 'c2':
            tis, rows, cols, pr, pc = synth.synth_positions(all_in_tissue=True)
            self.img = torch.randint(0, 256, (16512, 16000, 3), device=dev, dtype=torch.uint8, generator=gen)
            self.cells, _ = ip.spot_table(tis, rows, cols, pr, pc, dev)
            self.patches = torch.empty((1, H_ST, W_ST, 3, P, P), device=dev, dtype=torch.bfloat16)
        elif cfg in ('c3', 'c5'):
            self.patches = torch.empty((B, H_ST, W_ST, 3, P, P), device=dev, dtype=torch.bfloat16)
            self.patches_u8 = torch.randint(0, 256, (B, H_ST, W_ST, 3, P, P), device=dev, dtype=torch.uint8, generator=gen)
            ip.normalize_patches(self.patches_u8, MEAN, STD, torch.bfloat16, out=self.patches)
        if cfg in ('c1', 'c3'):
            # several distinct resident count arrays, visited in turn, so that a step never finds its input in L2
            self.n_rot = 4 if cfg == 'c1' else 1
            self.counts = [torch.log1p(torch.poisson(torch.ones(B, G_GENES, H_ST, W_ST, device=dev), generator=gen)) for _ in range(self.n_rot)]
        if cfg == 'c4':
            self.fgrid = torch.randn(B, N_CLS, H_ST, W_ST, device=dev, generator=gen).requires_grad_(True)

    # ---- the training step on device-resident inputs; `slot` selects among rotating inputs
    def n_slots(self):
        return getattr(self, 'n_rot', 1)

    def model_inputs(self, slot=0):
        c = self.cfg
        if c in ('c2', 'c5'):
            return self.patches
        if c == 'c1':
            return self.counts[slot]
        if c == 'c3':
            return [self.patches, self.counts[0]]
        return self.fgrid

    def train_on(self, inputs, labels):
        from gridnext_b200.training import gridwise_step
        loss, acc, _ = gridwise_step(self.model, inputs, labels, self.crit, 1, True)
        if self.bucket is not None:
            self.bucket.allreduce_mean()
        if self.opt is not None:
            self.opt.step()
            self.opt.zero_grad(set_to_none=(self.bucket is None))
        else:
            self.fgrid.grad = None
            for p in self.params:
                p.grad = None
        return loss.detach()          # nothing of the autograd graph outlives the step (AccumulateGrad nodes are per iteration)

    def step_resident(self, slot=0):
        if self.cfg == 'c2':
            self.ip.gather_patches(self.img, self.cells, P, MEAN, STD, torch.bfloat16, out=self.patches[0])
        return self.train_on(self.model_inputs(slot), self.labels)

    # ---- host-fed variant: what must cross PCIe every step
    def make_host_inputs(self):
        """Pinned host copies of the step's inputs as a dataset hands them over, and the device staging buffers."""
        c, ip = self.cfg, self.ip
        self.h_labels = torch.empty(self.labels.shape, dtype=self.labels.dtype, pin_memory=True); self.h_labels.copy_(self.labels)
        self.d_labels = torch.empty_like(self.labels)
        h2d = self.h_labels.numel() * 8
        if c == 'c2':
            raw = ip.gather_patches(self.img, self.cells, P, None, None, torch.float32).to(torch.uint8)
            self.h_u8 = torch.empty(raw.shape, dtype=torch.uint8, pin_memory=True); self.h_u8.copy_(raw)
            del raw
        elif c in ('c3', 'c5'):
            self.h_u8 = torch.empty(self.patches_u8.shape, dtype=torch.uint8, pin_memory=True); self.h_u8.copy_(self.patches_u8)
        if c in ('c2', 'c3', 'c5'):
            self.d_u8 = [torch.empty(self.h_u8.shape, device=self.dev, dtype=torch.uint8) for _ in range(2)]
            h2d += self.h_u8.numel()
        if c in ('c1', 'c3'):
            self.h_counts = torch.empty(self.counts[0].shape, dtype=torch.float32, pin_memory=True); self.h_counts.copy_(self.counts[0])
            self.d_counts = [torch.empty_like(self.counts[0]) for _ in range(2)]
            h2d += self.h_counts.numel() * 4
        if c == 'c4':
            self.h_f = torch.empty(self.fgrid.shape, dtype=torch.float32, pin_memory=True); self.h_f.copy_(self.fgrid.detach())
            self.d_f = [torch.empty_like(self.fgrid.detach()).requires_grad_(True) for _ in range(2)]
            h2d += self.h_f.numel() * 4
        self.h2d_bytes = h2d
        what = {'c1': 'fp32 count slab (1,5000,78,64) + int64 labels, pinned', 'c2': 'uint8 patch grid (78,64,3,128,128) + int64 labels, pinned',
                'c3': 'uint8 patch grids (12,78,64,3,128,128) + fp32 count slabs (12,5000,78,64) + labels, pinned',
                'c4': 'fp32 f-output grids (256,7,78,64) + labels, pinned', 'c5': 'uint8 patch grids (4,78,64,3,128,128) + labels, pinned'}[c]
        return what

    def copy_in(self, slot):
        """H2D copies of one step's inputs into double-buffer half `slot` (called on the copy stream)."""
        c = self.cfg
        if c in ('c2', 'c3', 'c5'):
            self.d_u8[slot].copy_(self.h_u8, non_blocking=True)
        if c in ('c1', 'c3'):
            self.d_counts[slot].copy_(self.h_counts, non_blocking=True)
        if c == 'c4':
            self.d_f[slot].data.copy_(self.h_f, non_blocking=True)

    def e2e_inputs(self, slot):
        """Device-side preparation of the copied inputs (normalise uint8 patches) -> model inputs."""
        c = self.cfg
        if c in ('c2', 'c3', 'c5'):
            self.ip.normalize_patches(self.d_u8[slot], MEAN, STD, torch.bfloat16, out=self.patches)
        if c in ('c2', 'c5'):
            return self.patches
        if c == 'c1':
            return self.d_counts[slot]
        if c == 'c3':
            return [self.patches, self.d_counts[slot]]
        return self.d_f[slot]

    def flop_per_step(self):
        c = self.cfg
        f = 0.0
        if c in ('c2', 'c3', 'c5'):
            f += F_FLOP_FWD_BWD * SPOTS * self.arrays * (4.0 / 3.0 if c in ('c3', 'c5') else 1.0)      # chunked path re-runs the forward
        if c in ('c1', 'c3'):
            f += MLP_FLOP_FWD_BWD * SPOTS * self.arrays
        return f


# ----------------------------------------------------------------------------------------------- B200 arm: driver
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a captured CUDA graph')
    ap.add_argument('--foreach-adam', action='store_true', help='torch.optim.Adam in its for-each form (~750 launches per step) instead of fused=True')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel time table of the instrumented step (c4: the sweep) to this JSON file')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from gridnext_b200 import _lib
    cfg = args.config
    if cfg == 'c5':
        return run_c5(args, dev, rank, world, local)
    wl = Workload(cfg, dev, rank, world, fused_adam=not args.foreach_adam)
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- resident leg
    for i in range(warm):
        wl.step_resident(i % wl.n_slots())
    _lib.LAUNCHES[0] = 0
    wl.step_resident(0)
    launches_per_step = _lib.LAUNCHES[0]
    graphs = []
    k_res = [0]

    def run_eager():
        wl.step_resident(k_res[0] % wl.n_slots())
        k_res[0] += 1
    run_step, graphed = run_eager, False
    if not args.no_graph and cfg != 'c3':
        # The step is a fixed launch sequence: capture it once per rotating input and replay, so the timed region measures the
        # GPU and not the Python/ctypes launch overhead of the host loop.  (c3 is launched eagerly: its 12-array step keeps as many
        # chunks of DenseNet activations as device memory holds and recomputes the rest, a decision taken from the allocator's state
        # that cannot be made under capture -- 902 ms eager against 1,030 ms replayed with every chunk recomputed.)
        try:
            torch.cuda.synchronize()
            torch.cuda.empty_cache()          # the eager warm-up's cached blocks would otherwise sit beside the graph's private pool
            pool = torch.cuda.graph_pool_handle()     # all graphs of this process replay one at a time: they share one memory pool
            for slot in range(wl.n_slots()):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, capture_error_mode='thread_local'):
                    wl.step_resident(slot)
                g.replay()
                graphs.append(g)
            torch.cuda.synchronize()

            def run_graph():
                graphs[k_res[0] % len(graphs)].replay()
                k_res[0] += 1
            run_step, graphed = run_graph, True
        except Exception as exc:      # capture is an optimisation of the launch path only
            if rank == 0:
                sys.stderr.write('CUDA graph capture failed, timing the eager step: %r\n' % (exc,))
            graphs = []
            torch.cuda.synchronize()
    for _ in range(2):
        run_step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(run_step, args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- host-fed leg: the next step's inputs travel H2D on a copy stream into the other half of a double buffer while the
    # current step computes (what a DataLoader with pinned memory + non_blocking copies does); every timed step issues exactly
    # one H2D copy of its inputs and one D2H read of its loss.
    host_what = wl.make_host_inputs()
    e2e_loss = [None]
    cur_inputs = [None]

    def e2e_train_part():
        e2e_loss[0] = wl.train_on(cur_inputs[0], wl.d_labels)
    e2e_graphs = [None, None]
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_k = [0]

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            wl.copy_in(slot)
            copied[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    prefetch(0)
    # inputs that are consumed in place from the double buffer (c1 counts, c3 counts, c4 grids) need one graph per half
    in_place = cfg in ('c1', 'c3', 'c4')

    def step_e2e():
        cur = e2e_k[0] & 1
        e2e_k[0] += 1
        main_s = torch.cuda.current_stream()
        main_s.wait_event(copied[cur])
        wl.d_labels.copy_(wl.h_labels, non_blocking=True)
        cur_inputs[0] = wl.e2e_inputs(cur)
        g = e2e_graphs[cur if in_place else 0]
        if g is not None:
            g.replay()
        else:
            e2e_train_part()
        consumed[cur].record(main_s)
        prefetch(cur ^ 1)
        return float(e2e_loss[0].item())          # D2H read of the step's result

    if not graphed:
        step_e2e(); step_e2e()
    if graphed:
        try:
            torch.cuda.synchronize()
            wl.d_labels.copy_(wl.h_labels)
            wl.copy_in(0); wl.copy_in(1)      # both halves of the double buffer hold real inputs while the graphs are captured (and replayed once)
            torch.cuda.synchronize()
            for half in ((0, 1) if in_place else (0,)):
                cur_inputs[0] = wl.e2e_inputs(half)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, capture_error_mode='thread_local'):
                    e2e_train_part()
                g.replay()
                e2e_graphs[half] = g
            torch.cuda.synchronize()
        except Exception as exc:
            if rank == 0:
                sys.stderr.write('CUDA graph capture of the host-fed step failed: %r\n' % (exc,))
            e2e_graphs = [None, None]
            torch.cuda.synchronize()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- instrumented step: time share of every C-ABI entry point (every rank runs it: it contains the all-reduce)
    torch.cuda.synchronize()
    if rank == 0:
        _lib.PROFILE = {}
    wl.step_resident(0)
    torch.cuda.synchronize()
    roof = None
    if rank == 0:
        prof, _lib.PROFILE = _lib.PROFILE, None
        roof = roofline_from_profile(prof, wl, ms / args.steps, args)
    sweep = None
    if cfg == 'c4' and rank == 0 and args.profile_out:
        sweep = c4_sweep(dev)
        json.dump(sweep, open(args.profile_out, 'w'), indent=1)

    infer = None
    if rank == 0 and world == 1 and cfg in ('c1', 'c2'):
        # the evaluation loop of utils.all_fgd_predictions (reference utils.py:20-57) on the same resident inputs: gather (c2) -> f -> g ->
        # fused foreground softmax / argmax, no gradients; eager launches, one host read per array (the loop's own contract)
        from gridnext_b200.utils import fg_predictions
        wl.model.eval()
        try:
            def infer_step():
                with torch.no_grad():
                    if cfg == 'c2':
                        wl.ip.gather_patches(wl.img, wl.cells, P, MEAN, STD, torch.bfloat16, out=wl.patches[0])
                    out = wl.model(wl.model_inputs(0))
                    return fg_predictions(out, wl.labels)
            for _ in range(3):
                infer_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                infer_step()
            e1.record()
            torch.cuda.synchronize()
            ms_inf = e0.elapsed_time(e1) / args.steps
            infer = dict(value=SPOTS * wl.arrays / (ms_inf * 1e-3), unit='spots/s', ms_per_step=ms_inf,
                         what='all_fgd_predictions step: f + g forward + foreground softmax/argmax, no grad, eager launches')
        finally:
            wl.model.train(); wl.model.patch_classifier.eval()
    if rank == 0:
        eager = None
        if cfg == 'c2' and not args.no_eager_baseline:
            try:
                from tools import eager_baseline
                from gridnext_b200.densenet import DenseNet
                from gridnext_b200.gridnet_models import GridNetHexOddr
                m2 = GridNetHexOddr(DenseNet(num_classes=N_CLS, small_inputs=False, efficient=False, drop_rate=0, **DENSENET_KW), (3, P, P), (H_ST, W_ST), N_CLS)
                m2.load_state_dict(wl.model.state_dict())
                m2.to(dev)
                m2.train(); m2.patch_classifier.eval()
                eager = eager_baseline.run(m2, wl.patches.view(-1, 3, P, P), wl.labels)
                eager['what'] = ('reference module graph (DenseNet-121 f eval, torch.cat concat; hex g as dense 3x3 pairs; masked CE) under PyTorch '
                                 'eager on this GPU, f fwd+bwd in 192-spot chunks without recompute, no optimizer step')
                del m2
            except Exception as exc:
                eager = dict(error=repr(exc))
            torch.cuda.empty_cache()
        cpu = None
        if not args.no_cpu_baseline:
            fn, meta = cpu_reference_step_fn(cfg)
            fn()
            ts = sorted(fn() for _ in range(3))
            cpu = dict(value=SPOTS * wl.arrays / ts[1], unit='spots/s', **meta)
        per_step = ms / args.steps
        spots = SPOTS * wl.arrays * world
        line = dict(metric=METRIC, value=spots / (per_step * 1e-3), unit='spots/s', n_gpus=world,
                    steps=args.steps, warmup=warm, ms_per_step=per_step, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32' if cfg == 'c4' else 'bf16', data='synthetic',
                    config=dict(workload_config(cfg, world), launch='cuda_graph_replay' if graphed else 'eager',
                                optimizer=('none' if wl.opt is None else 'torch.optim.Adam(fused=%s, capturable=True)' % (not args.foreach_adam))),
                    clocks=clocks, gpu_launches=launches_per_step * args.steps,
                    e2e=dict(value=spots / (ms_e2e / args.steps * 1e-3), unit='spots/s', h2d_bytes_per_step=wl.h2d_bytes, d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e / args.steps, host_input=host_what,
                             last_loss=float(e2e_loss[0].item())),
                    roofline=roof, cpu_baseline=cpu)
        if eager is not None:
            line['gpu_eager_baseline'] = eager
        if infer is not None:
            line['inference'] = infer
        print(json.dumps(line), flush=True)
    shutdown(world, graphs + [g for g in e2e_graphs if g is not None])


def shutdown(world, graphs):
    """Captured graphs hold NCCL work: release them before the process group (tearing the communicator down first hung at N = 2)."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    for g in graphs:
        g.reset()
    del graphs
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # clean teardown; the watchdog only fires if NCCL's communicator destruction wedges (seen once in round 1 with live graphs)
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.barrier()
        dist.destroy_process_group()
        wd.cancel()


def roofline_from_profile(prof, wl, step_ms, args):
    """Per-kernel table of the instrumented step -> the `roofline` object of the bench line (SURVEY.md 8d bounds)."""
    table = {k: dict(calls=len(v), ms=sum(a.elapsed_time(b) for a, b, _ in v)) for k, v in prof.items()}
    total = sum(t['ms'] for t in table.values())
    flops, nbytes = {}, {}

    def add(k, f, by):
        flops[k] = flops.get(k, 0) + f
        nbytes[k] = nbytes.get(k, 0) + by
    for a, b, g in prof.get('gn_gemm_bf16', []):
        M_, N_, K_ = g[4], g[5], g[6]
        add('gn_gemm_bf16', 2.0 * M_ * N_ * K_, 2.0 * M_ * (K_ + N_ * (3 if g[16] else 1)))       # BN-backward epilogue: read ref, read+write out
    for a, b, g in prof.get('gn_gemm_tn_bf16', []):
        add('gn_gemm_tn_bf16', 2.0 * g[4] * g[5] * g[6], 2.0 * g[6] * (g[4] + g[5]))
    for a, b, g in prof.get('gn_conv1x1_bwd_bf16', []):          # data gradient + weight gradient of the bottleneck conv: one pass over dz, C, dC
        M_, N_ = g[4], g[5]
        add('gn_conv1x1_bwd_bf16', 4.0 * M_ * N_ * 128, 2.0 * M_ * (128 + N_ * (3 if g[16] else 2)))
    for a, b, g in prof.get('gn_conv3x3_bf16', []):
        px = g[2] * g[3] * g[4]
        add('gn_conv3x3_bf16', 2.0 * 9 * px * g[5] * g[8], 2.0 * px * (g[5] + g[8] * (2 if g[11] else 1)))
    for a, b, g in prof.get('gn_conv3x3_wgrad_bf16', []):
        px = g[4] * g[5] * g[6]
        add('gn_conv3x3_wgrad_bf16', 2.0 * 9 * px * g[7] * g[8], 2.0 * px * (g[7] + g[8]))
    # g-side kernels: algorithmic bytes 4 (Cin + Cout) per cell forward; weight gradient reads x and dy
    for name in ('gn_hexconv_fwd', 'gn_hexconv_fwd_tc', 'gn_hexconv_fwd_tc2'):
        for a, b, g in prof.get(name, []):
            B_, ci_, co_, H_, W_ = g[7], g[8], g[9], g[10], g[11]
            add(name, 2.0 * 7 * ci_ * co_ * B_ * H_ * W_, 4.0 * (ci_ + co_) * B_ * H_ * W_)
    for a, b, g in prof.get('gn_hexconv_wgrad', []):
        B_, ci_, co_, H_, W_ = g[6], g[7], g[8], g[9], g[10]
        add('gn_hexconv_wgrad', 2.0 * 7 * ci_ * co_ * B_ * H_ * W_, 4.0 * (ci_ + co_) * B_ * H_ * W_)
    for a, b, g in prof.get('gn_hexconv_wgrad_tc', []):
        B_, ci_, co_, H_, W_ = g[5], g[6], g[7], g[8], g[9]
        add('gn_hexconv_wgrad_tc', 2.0 * 7 * ci_ * co_ * B_ * H_ * W_, 4.0 * (ci_ + co_) * B_ * H_ * W_)
    for a, b, g in prof.get('gn_hexconv_wgrad_tc2', []):
        B_, ci_, co_, H_, W_ = g[6], g[7], g[8], g[9], g[10]
        add('gn_hexconv_wgrad_tc2', 2.0 * 7 * ci_ * co_ * B_ * H_ * W_, 4.0 * (ci_ + co_) * B_ * H_ * W_)
    for a, b, g in prof.get('gn_patch_gather', []):
        add('gn_patch_gather', 0.0, float(g[5]) * 3 * g[6] * g[6] * (1 + (2 if g[10] else 4)))
    for k in table:
        table[k]['share'] = table[k]['ms'] / total if total else 0
        if k in flops:
            table[k]['tflops'] = flops[k] / (table[k]['ms'] * 1e-3) / 1e12
            table[k]['gbs'] = nbytes[k] / (table[k]['ms'] * 1e-3) / 1e9
    dom = max(table, key=lambda k: table[k]['ms'])
    pk = peaks()
    traffic = committed_traffic().get(wl.cfg, {}).get(dom)
    tensor_bound = dom in ('gn_gemm_bf16', 'gn_gemm_tn_bf16', 'gn_conv1x1_bwd_bf16', 'gn_conv3x3_bf16', 'gn_conv3x3_wgrad_bf16', 'gn_stem_conv_fwd', 'gn_stem_conv_wgrad')
    roof = dict(kernel=dom, launches=table[dom]['calls'], share_of_step=table[dom]['share'], peak_source=pk['src'],
                traffic=None if traffic is None else traffic.get('dram_bytes_per_launch'))
    if dom in flops and tensor_bound:
        # the DenseNet / MLP contractions are assigned to the tensor pipe.  achieved = algorithmic FLOPs of all launches of
        # the kernel / their summed CUDA-event time.  The same kernel against the HBM copy rate is carried beside it: with K or N of
        # 32..128 per layer the layer-by-layer design makes these GEMMs stream their operands, so `hbm` is what actually limits them.
        calls = table[dom]['calls']
        roof.update(bound='tensor', achieved=table[dom]['tflops'], peak=pk['bf16'], unit='TFLOP/s', frac=table[dom]['tflops'] / pk['bf16'],
                    flop_per_launch=flops[dom] / calls, algorithmic_bytes_per_launch=nbytes[dom] / calls,
                    hbm=dict(achieved=table[dom]['gbs'], peak=pk['hbm'], unit='GB/s', frac=table[dom]['gbs'] / pk['hbm']),
                    note='achieved = algorithmic FLOPs of all %d launches / their summed CUDA-event time (the launches differ in shape; per-launch '
                         'figures are averages); traffic = ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of the same step' % calls)
    elif dom in nbytes:
        calls = table[dom]['calls']
        roof.update(bound='hbm', achieved=table[dom]['gbs'], peak=pk['hbm'], unit='GB/s', frac=table[dom]['gbs'] / pk['hbm'],
                    algorithmic_bytes_per_launch=nbytes[dom] / calls)
    else:
        roof.update(bound='hbm', achieved=None, peak=pk['hbm'], unit='GB/s', frac=None)
    if wl.flop_per_step() > 0:
        roof['whole_step_tensor_frac'] = wl.flop_per_step() / (step_ms * 1e-3) / 1e12 / pk['bf16']
    if wl.cfg == 'c4':
        cells = SPOTS * wl.arrays
        roof['whole_step_hbm_frac'] = G_BYTES_PER_CELL * cells / (step_ms * 1e-3) / 1e9 / pk['hbm']
        # the g-only step is a handful of kernels: carry all of them (ms summed over their launches, algorithmic GB/s where defined)
        roof['kernels'] = {k: dict(calls=t['calls'], ms=round(t['ms'], 4), **({'gbs': round(t['gbs'], 1)} if 'gbs' in t else {}))
                           for k, t in sorted(table.items(), key=lambda kv: -kv[1]['ms'])}
    if args.profile_out and wl.cfg != 'c4':
        def ints(g):
            return [a if isinstance(a, int) else (None if a is None else 'p') for a in g]
        calls = {k: [dict(ms=a.elapsed_time(b), args=ints(g)) for a, b, g in v] for k, v in prof.items()
                 if k in ('gn_gemm_bf16', 'gn_gemm_tn_bf16', 'gn_conv1x1_bwd_bf16', 'gn_conv3x3_bf16', 'gn_conv3x3_wgrad_bf16')}
        json.dump(dict(step_ms=step_ms, instrumented_total_ms=total, kernels=table, calls=calls), open(args.profile_out, 'w'), indent=1, sort_keys=True)
    return roof


def c4_sweep(dev):
    """BASELINE configs[3]: single hexagdly.Conv2d(C, C, k) forward and forward+backward over C x k x B, graph-replayed.
    Algorithmic bytes: forward 4 (Cin + Cout) per cell; backward 4 (Cout + 2 Cin) (dX) + 4 (Cin + Cout) (dW) per cell."""
    from gridnext_b200 import hexagdly as hx
    pk = peaks()
    rows = []
    for k in (1, 2, 3):
        T = 1 + 3 * k * (k + 1)
        for C in (4, 8, 16, 32, 64):
            for B in (1, 4, 16, 64, 256):
                conv = hx.Conv2d(C, C, k).to(dev)
                ks = hx._kernels(conv)
                x = torch.randn(B, C, H_ST, W_ST, device=dev, requires_grad=True)
                dy = torch.randn(B, C, H_ST, W_ST, device=dev)
                cells = B * SPOTS

                def fwd():
                    with torch.no_grad():
                        hx.hexconv_visium(x, ks, conv.bias_tensor)

                def bwd():          # forward + backward through the public functional (dx, dW, db), as autograd runs it
                    y = hx.hexconv_visium(x, ks, conv.bias_tensor)
                    y.backward(dy)
                    x.grad = None
                    for t in conv.parameters():
                        t.grad = None
                for name, fn, by, fl in (('fwd', fwd, 8.0 * C * cells, 2.0 * T * C * C * cells), ('fwd_bwd', bwd, 28.0 * C * cells, 6.0 * T * C * C * cells)):
                    fn(); torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fn()
                    g.replay(); torch.cuda.synchronize()
                    ts = []
                    for _ in range(5):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ms = sorted(ts)[2]
                    rows.append(dict(case='hexconv_' + name, k=k, C=C, B=B, ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), hbm_frac=round(by / ms / 1e6 / pk['hbm'], 3),
                                     tflops=round(fl / ms / 1e9, 2), regime='hbm' if C * T <= 46 else 'fp32-fma / tensor'))
                    g.reset()
                del x, dy
    return rows


# ----------------------------------------------------------------------------------------------- c5: the product training loop
def run_c5(args, dev, rank, world, local):
    """BASELINE configs[4] through the product API: gridnext_b200.training.train_gridwise (reference training.py:101-209) over a
    torch DataLoader of PatchGridDataset-style items (uint8-derived fp32 patch grids normalised by the dataset transform are
    what the reference hands over; here the dataset yields bf16 patch grids already normalised on the host side once, pinned),
    batch 4 arrays per GPU (19,968 spots per step: the chunked forward + recompute-backward path), one optimizer step per batch,
    one flat gradient all-reduce per step.  A 'step' of the bench contract = one batch; K steps = one train phase over 4K arrays."""
    import io
    import contextlib
    import torch.distributed as dist
    import torch.nn as nn
    from gridnext_b200 import _lib, training
    wl = Workload('c5', dev, rank, world, fused_adam=not args.foreach_adam)
    B = wl.arrays
    # 4 distinct arrays in pinned host memory; item i of the epoch is array i % 4 (256 synthetic arrays would be 63 GB of host memory)
    host = wl.patches.cpu().pin_memory()
    host_lab = wl.labels.cpu().pin_memory()
    bytes_per_step = host.numel() * 2 + host_lab.numel() * 8

    class Batches(torch.utils.data.Dataset):
        """Pre-batched items (batch_size=None in the DataLoader): a 4-array batch is a view of pinned host memory, so the loader
        hands it over without a 2 GB collate copy and ``.to(device, non_blocking=True)`` in train_gridwise is an async H2D copy."""

        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return host, host_lab

    def loaders(n_steps):
        return {'train': torch.utils.data.DataLoader(Batches(n_steps), batch_size=None),
                'val': torch.utils.data.DataLoader(Batches(0), batch_size=None)}

    def epoch(n_steps):
        with contextlib.redirect_stdout(io.StringIO()):
            _, vh, th = training.train_gridwise(wl.model, loaders(n_steps), wl.crit, wl.opt, num_epochs=1)
        return th[0]

    warm = max(args.warmup, 3)
    epoch(warm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    _lib.LAUNCHES[0] = 0
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = epoch(args.steps)                 # includes the H2D copies of every batch and the D2H read of the phase's counters
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = _lib.LAUNCHES[0]
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        _lib.PROFILE = {}
    training.gridwise_step(wl.model, wl.patches, wl.labels, wl.crit, 1, True)
    torch.cuda.synchronize()
    roof = None
    if rank == 0:
        prof, _lib.PROFILE = _lib.PROFILE, None
        roof = roofline_from_profile(prof, wl, ms / args.steps, args)
        cpu = None
        if not args.no_cpu_baseline:
            fn, meta = cpu_reference_step_fn('c5')
            fn()
            ts = sorted(fn() for _ in range(3))
            cpu = dict(value=SPOTS * B / ts[1], unit='spots/s', **meta)
        per_step = ms / args.steps
        spots = SPOTS * B * world
        val = spots / (per_step * 1e-3)
        line = dict(metric=METRIC, value=val, unit='spots/s', n_gpus=world, steps=args.steps, warmup=warm, ms_per_step=per_step,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                    config=dict(workload_config('c5', world), launch='eager (train_gridwise)', api='gridnext_b200.training.train_gridwise + torch DataLoader',
                                epoch_loss=loss, note='value and e2e are the same measurement: the product loop is host-fed by construction'),
                    clocks=clocks, gpu_launches=launches,
                    e2e=dict(value=val, unit='spots/s', h2d_bytes_per_step=bytes_per_step, d2h_bytes_per_step=32, ms_per_step=per_step,
                             host_input='bf16 normalised patch grids (4,78,64,3,128,128) + int64 labels from a pinned DataLoader'),
                    roofline=roof, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    shutdown(world, [])


if __name__ == '__main__':
    main()
