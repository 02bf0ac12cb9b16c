#!/usr/bin/env python
"""Benchmark of the GridNet hot path on B200:  python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): Visium spots/sec for one f+g training step (forward, foreground-masked CE, backward, gradient
all-reduce when N > 1, Adam step) on synthetic 78 x 64 Visium arrays.  Workload at N = 1: BASELINE.json configs[1]
"GridNetHex image-only: DenseNet-121 f on 3x128x128 patches for all 4,992 spots + hex g" -- one array per GPU per
step (weak scaling: each rank owns its arrays, the only exchange is the flat gradient all-reduce).

Legs:
  value        inputs (the full-resolution uint8 image, spot table, labels) resident in HBM; a step = patch gather ->
               DenseNet f -> hex corrector g -> masked CE -> backward -> [all-reduce] -> Adam.  CUDA events, max over ranks.
  e2e          the same step fed from HOST memory the way train_gridwise's DataLoader feeds it: the uint8 patch grid
               (78, 64, 3, 128, 128) and labels in pinned memory are copied H2D every step, normalised on the device,
               and the scalar loss is read back (D2H) every step.
  roofline     one extra instrumented step (CUDA events around every C-ABI call) gives per-kernel time shares; the
               dominant kernel's algorithmic FLOPs / its measured time is compared with MEASURED_PEAKS.json.
  cpu_baseline the CPU oracle (oracle/gridnet_ref.py, the restated reference modules) timed on this box's cores on a
               bounded sample (f on 128 spots, g on the whole array) and scaled to one array.
  --impl reference   times only that CPU arm, K steps after W warm-ups, same metric/config.
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
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
DENSENET_KW = dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4)
# algorithmic work per spot: DenseNet-121 @128 px forward 1.850 GFLOP, forward+backward 5.47 GFLOP
F_FLOP_FWD_BWD = 5.47e9


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d['hbm_gbs'], bf16=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured (MEASURED_PEAKS.json, sustained bf16)')
    return dict(hbm=6650.0, bf16=1400.0, src='fallback (B200_PROFILING.md)')


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


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_reference_step_fn(sample_spots=128):
    """Returns (fn, meta): fn() runs f fwd+bwd on `sample_spots` spots + g fwd+bwd on a full array on the host cores."""
    from oracle import synth, shapes as S, gridnet_ref as R
    torch.set_num_threads(os.cpu_count())
    sd = synth.synth_state_dict(S.gridnet_shapes(S.densenet_shapes(**DENSENET_KW), N_CLS, N_CLS), 1234)
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k and k not in ('bg_const', 'dummy_tensor'):
            v.requires_grad_(True)
    g = torch.Generator(); g.manual_seed(7)
    x = torch.randn(sample_spots, 3, P, P, generator=g)
    labels = synth.synth_labels(1, N_CLS, seed=0)
    fgrid = torch.randn(1, N_CLS, H_ST, W_ST, generator=g)
    dlog = torch.randn(sample_spots, N_CLS, generator=g)

    def fn():
        t0 = time.perf_counter()
        logits = R.densenet_forward(R.sub(sd, 'patch_classifier.'), x)
        (logits * dlog).sum().backward()
        t1 = time.perf_counter()
        fg = fgrid.clone().requires_grad_(True)
        out = R.corrector_forward(R.sub(sd, 'corrector.'), fg, use_bn=True, training=True)
        loss, _, _ = R.masked_ce(out, labels)
        loss.backward()
        t2 = time.perf_counter()
        for v in sd.values():
            if v.is_floating_point() and v.grad is not None:
                v.grad = None
        t_array = (t1 - t0) * SPOTS / sample_spots + (t2 - t1)      # f is per-spot independent in eval mode
        return t_array
    meta = dict(cores=os.cpu_count(), kind='port',
                sample='f (DenseNet-121 fwd+bwd, fp32) on %d of 4992 spots scaled linearly + g (5 hex convs, BN, masked CE fwd+bwd) on the full array' % sample_spots)
    return fn, meta


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    fn, meta = cpu_reference_step_fn()
    for _ in range(args.warmup):
        fn()
    ts = [fn() for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    val = SPOTS / t
    line = dict(metric='visium_spots_per_sec_f+g_fwd+bwd', value=val, unit='spots/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=workload_config(args.gpus), cpu_baseline=dict(value=val, unit='spots/s', **meta),
                e2e=dict(value=val, unit='spots/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return dict(workload='GridNetHex image-only: DenseNet-121 f on 3x128x128 patches for all 4,992 spots + 5-layer hex g (BASELINE configs[1])',
                arrays_per_gpu_per_step=1, spots_per_array=SPOTS, patch=P, n_classes=N_CLS, parallelism='dp%d' % n_gpus,
                step='patch gather + f fwd/bwd + g fwd/bwd + masked CE + grad all-reduce + Adam', l2='inputs and activations (>20 GB) far exceed the 126 MB L2')


# ----------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a captured CUDA graph')
    ap.add_argument('--fused-adam', action='store_true',
                    help='torch.optim.Adam(fused=True): a few fused launches instead of ~750 for-each / single-tensor kernels per step '
                         '(DESIGN.md section 8 item 1; off by default until it has been run on a GPU)')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel time table of the instrumented step to this JSON file')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    import torch.nn as nn
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from gridnext_b200 import _lib, imgprocess as ip, parallel
    from gridnext_b200.densenet import DenseNet
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.training import gridwise_step
    from synthdata import synth, shapes as S   # seeded synthetic weights / inputs (data generators; nothing from oracle/ on this arm)

    # ---- model: reference constructor surface, synthetic weights
    f = DenseNet(num_classes=N_CLS, small_inputs=False, efficient=False, drop_rate=0, **DENSENET_KW)
    model = GridNetHexOddr(f, (3, P, P), (H_ST, W_ST), N_CLS, use_bn=True, atonce_patch_limit=None)
    model.load_state_dict(synth.synth_state_dict(S.gridnet_shapes(S.densenet_shapes(**DENSENET_KW), N_CLS, N_CLS), 1234))
    model.to(dev)
    model.train(); model.patch_classifier.eval()          # training.py:120-126
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, capturable=True, fused=True if args.fused_adam else None)
    bucket = parallel.GradBucket(params) if world > 1 else None
    crit = nn.CrossEntropyLoss()

    # ---- synthetic array: full-resolution uint8 image + Visium-style positions (all 4,992 spots in tissue), labels
    tis, rows, cols, pr, pc = synth.synth_positions(all_in_tissue=True)
    Himg, Wimg = 16512, 16000
    gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
    img = torch.randint(0, 256, (Himg, Wimg, 3), device=dev, dtype=torch.uint8, generator=gen)
    cells, _ = ip.spot_table(tis, rows, cols, pr, pc, dev)
    labels = synth.synth_labels(1, N_CLS, seed=rank).to(dev)
    patches_buf = torch.empty((H_ST, W_ST, 3, P, P), device=dev, dtype=torch.bfloat16)

    def train_on(patches):
        loss, acc, _ = gridwise_step(model, patches.view(1, H_ST, W_ST, 3, P, P), labels, crit, 1, True)
        if bucket is not None:
            bucket.allreduce_mean()
        opt.step()
        opt.zero_grad(set_to_none=(bucket is None))
        return loss

    def step_resident():
        ip.gather_patches(img, cells, P, MEAN, STD, torch.bfloat16, out=patches_buf)
        return train_on(patches_buf)

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

    # host-fed leg: uint8 patch grid + labels in pinned memory (what a PatchGridDataset batch is before ToTensor)
    ip.gather_patches(img, cells, P, None, None, torch.float32, out=None)     # warm the raw path once
    raw = ip.gather_patches(img, cells, P, None, None, torch.float32).to(torch.uint8)
    host_patches = torch.empty(raw.shape, dtype=torch.uint8, pin_memory=True); host_patches.copy_(raw)
    host_labels = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True); host_labels.copy_(labels)
    del raw
    dev_u8 = torch.empty(host_patches.shape, device=dev, dtype=torch.uint8)
    dev_lab = torch.empty_like(labels)

    # host-fed leg: H2D copies and the normalise kernel are launched eagerly into static buffers; the training step that
    # consumes them is the same captured launch sequence as above (minus the gather)
    e2e_loss = [None]

    def e2e_train_part():
        loss, acc, _ = gridwise_step(model, patches_buf.view(1, H_ST, W_ST, 3, P, P), dev_lab, crit, 1, True)
        if bucket is not None:
            bucket.allreduce_mean()
        opt.step()
        opt.zero_grad(set_to_none=(bucket is None))
        e2e_loss[0] = loss

    e2e_run = [e2e_train_part]

    # Input pipeline of the host-fed leg: the next step's uint8 patches travel H2D on a copy stream into the other half of
    # a double buffer while the current step computes (what a DataLoader with pinned memory + non_blocking copies does);
    # every timed step issues exactly one 245 MB H2D copy and one D2H read of its loss.
    copy_stream = torch.cuda.Stream()
    dev_u8_pair = [dev_u8, torch.empty_like(dev_u8)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_k = [0]

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            dev_u8_pair[slot].copy_(host_patches, non_blocking=True)
            copied[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    prefetch(0)

    def step_e2e():
        cur = e2e_k[0] & 1
        e2e_k[0] += 1
        main = torch.cuda.current_stream()
        main.wait_event(copied[cur])
        dev_lab.copy_(host_labels, non_blocking=True)
        ip.normalize_patches(dev_u8_pair[cur], MEAN, STD, torch.bfloat16, out=patches_buf)
        consumed[cur].record(main)
        prefetch(cur ^ 1)
        e2e_run[0]()
        return float(e2e_loss[0].item())          # D2H read of the step's result

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # The step is a fixed launch sequence (~600 kernels + the all-reduce): capture it once in a CUDA graph and replay it,
    # so the timed region measures the GPU and not the Python/ctypes launch overhead of the host loop.
    _lib.LAUNCHES[0] = 0
    step_resident()
    launches_per_step = _lib.LAUNCHES[0]
    run_step, graphed = step_resident, False
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                step_resident()
            graph.replay()
            torch.cuda.synchronize()
            run_step, graphed = graph.replay, True
            # the host-fed leg replays the same launch sequence minus the gather (its copies + normalise stay eager)
            dev_lab.copy_(labels)
            graph_e2e = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_e2e, capture_error_mode='thread_local'):
                e2e_train_part()
            graph_e2e.replay()
            torch.cuda.synchronize()
            e2e_run[0] = graph_e2e.replay
        except Exception as exc:      # capture is an optimisation of the launch path only
            if rank == 0:
                sys.stderr.write('CUDA graph capture failed, timing the eager step: %r\n' % (exc,))
            torch.cuda.synchronize()
    for _ in range(2):
        run_step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(run_step, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- instrumented step: time share of every C-ABI entry point
    roof = None
    # every rank runs the instrumented step (it contains the gradient all-reduce); only rank 0 records it
    torch.cuda.synchronize()
    if rank == 0:
        _lib.PROFILE = {}
    step_resident()
    torch.cuda.synchronize()
    if rank == 0:
        prof, _lib.PROFILE = _lib.PROFILE, None
        table = {k: dict(calls=len(v), ms=sum(a.elapsed_time(b) for a, b, _ in v)) for k, v in prof.items()}
        total = sum(t['ms'] for t in table.values())
        # algorithmic FLOPs and HBM bytes of the tensor-core entry points from their arguments
        flops, nbytes = {}, {}

        def add(k, f, by):
            flops[k] = flops.get(k, 0) + f
            nbytes[k] = nbytes.get(k, 0) + by
        for a, b, g in prof.get('gn_gemm_bf16', []):
            M_, N_, K_ = g[4], g[5], g[6]
            add('gn_gemm_bf16', 2.0 * M_ * N_ * K_, 2.0 * M_ * (K_ + N_ * (3 if g[16] else 1)))       # BN-backward epilogue: read ref, read+write out
        for a, b, g in prof.get('gn_gemm_tn_bf16', []):
            add('gn_gemm_tn_bf16', 2.0 * g[4] * g[5] * g[6], 2.0 * g[6] * (g[4] + g[5]))
        for a, b, g in prof.get('gn_conv3x3_bf16', []):
            px = g[2] * g[3] * g[4]
            add('gn_conv3x3_bf16', 2.0 * 9 * px * g[5] * g[8], 2.0 * px * (g[5] + g[8] * (2 if g[11] else 1)))
        for a, b, g in prof.get('gn_conv3x3_wgrad_bf16', []):
            px = g[4] * g[5] * g[6]
            add('gn_conv3x3_wgrad_bf16', 2.0 * 9 * px * g[7] * g[8], 2.0 * px * (g[7] + g[8]))
        for k in table:
            table[k]['share'] = table[k]['ms'] / total if total else 0
            if k in flops:
                table[k]['tflops'] = flops[k] / (table[k]['ms'] * 1e-3) / 1e12
                table[k]['gbs'] = nbytes[k] / (table[k]['ms'] * 1e-3) / 1e9
        dom = max(table, key=lambda k: table[k]['ms'])
        pk = peaks()
        if dom in flops:
            # the DenseNet GEMM/conv kernels are HBM-bound (K or N is 32..128 against multi-GB operands): report them against the
            # copy bandwidth; the tensor-pipe figure of the same kernel is carried beside it
            ach = table[dom]['gbs']
            roof = dict(bound='hbm', kernel=dom, achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'], traffic=None,
                        launches=table[dom]['calls'], share_of_step=table[dom]['share'], peak_source=pk['src'].replace('sustained bf16', 'copy bandwidth'),
                        tensor_tflops=table[dom]['tflops'], tensor_frac=table[dom]['tflops'] / pk['bf16'],
                        note='achieved = algorithmic bytes of all %d launches / their summed CUDA-event time; ncu dram bytes of single launches are in profiles/' % table[dom]['calls'])
        else:
            roof = dict(bound='hbm', kernel=dom, achieved=None, peak=pk['hbm'], unit='GB/s', frac=None, traffic=None,
                        share_of_step=table[dom]['share'], peak_source=pk['src'])
        roof['whole_step_tensor_frac'] = (F_FLOP_FWD_BWD * SPOTS / (ms / args.steps * 1e-3)) / 1e12 / pk['bf16']
        if args.profile_out:
            def ints(g):
                return [a if isinstance(a, int) else (None if a is None else 'p') for a in g]
            calls = {k: [dict(ms=a.elapsed_time(b), args=ints(g)) for a, b, g in v] for k, v in prof.items()
                     if k in ('gn_gemm_bf16', 'gn_gemm_tn_bf16', 'gn_conv3x3_bf16', 'gn_conv3x3_wgrad_bf16')}
            json.dump(dict(step_ms=ms / args.steps, instrumented_total_ms=total, kernels=table, calls=calls), open(args.profile_out, 'w'), indent=1, sort_keys=True)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            fn, meta = cpu_reference_step_fn()
            fn()
            t = min(fn() for _ in range(2))
            cpu = dict(value=SPOTS / t, unit='spots/s', **meta)
        per_step = ms / args.steps
        h2d = host_patches.numel() + host_labels.numel() * 8
        line = dict(metric='visium_spots_per_sec_f+g_fwd+bwd', value=SPOTS * world / (per_step * 1e-3), unit='spots/s', n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=per_step, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='bf16', data='synthetic', config=dict(workload_config(world), launch='cuda_graph_replay' if graphed else 'eager'), clocks=clocks, gpu_launches=launches,
                    e2e=dict(value=SPOTS * world / (ms_e2e / args.steps * 1e-3), unit='spots/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e / args.steps, host_input='uint8 patch grid (78,64,3,128,128) + int64 labels, pinned'),
                    roofline=roof, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    # Leave without tearing the process group down: destroying an NCCL communicator whose collectives were captured into
    # CUDA graphs that are still alive hung at exit (observed at N = 2); process exit releases everything.
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


if __name__ == '__main__':
    main()
