#!/usr/bin/env python
"""bench.py — images/sec/GPU, forward+backward(+optimizer step), ResNet-50 224x224 bs256 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic batch: Task.forward_with_gt -> JointLoss -> backward ->
(gradient exchange over NVLink) -> optimizer step, the step the reference's Lightning loop runs
(torchok/tasks/base.py:125-133).  Synthetic N(0,1) images, random labels, random-init weights (seed 42).

Workloads (`--workload`, default resnet50 = the BASELINE.json metric; the others are BASELINE.json's other configs):
    resnet50        C2  ResNet-50 ClassificationTask 3x224x224 bs256, SGD momentum   (classification_imagenet.yaml)
    resnet18_cifar  C1  ResNet-18 ClassificationTask 3x32x32 bs128, Adam             (classification_cifar10.yaml)
    swin_t          C3  Swin-T (V2, swinv2_custom img 224 window 7) bs256, AdamW
    hrnet_seg       C4  HRNet-W18 + HRNetSegmentationNeck + SegmentationHead 3x512x512 bs32, SGD
    retrieval       C5  IndexBasedMeter search: N x N cosine top-2, N = 1 M x 512 (query rows sharded over the ranks)

Output: ONE JSON line on rank 0 (see the driver contract).  `value` = device-resident inputs, whole-job img/s;
`e2e` = same loop fed from pinned HOST buffers through the public API (H2D of every batch and D2H of every loss inside
the timed region); `roofline` = the kernel FAMILY that takes the largest share of the step, every launch of one step
timed live with CUDA events on the launching stream and set against max(tensor floor, HBM floor) of its algorithmic
FLOPs / bytes (time-weighted fraction), with the per-family table beside it; `roofline_step` = whole-step FLOPs / time;
`cpu_baseline` = the CPU oracle (oracle/models.py, the torch.nn graph the reference dispatches) on the host cores;
`gpu_torch_baseline` = the same torch.nn graph on this GPU through cuDNN / cuBLAS (bf16 autocast, channels_last) — the
library path SURVEY 2.3 names as the one to beat.  `--impl reference` times the CPU path alone on the same config/metric.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = 'images/sec/GPU fwd+bwd ResNet-50 224² bs256; 1/2/4/8 GPU scaling'
UNIT = 'img/s'
CE = {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]}


def _cls(backbone, classes, opt, **bp):
    params = {'pretrained': False}
    params.update(bp)
    return {'task': {'name': 'ClassificationTask', 'params': {
        'backbone_name': backbone, 'backbone_params': params, 'pooling_name': 'Pooling',
        'head_name': 'ClassificationHead', 'head_params': {'num_classes': classes}}},
        'joint_loss': CE, 'optimization': [{'optimizer': opt}]}


# flop_per_img: SURVEY 8(d) (FlopCounterMode on the oracle, fwd + bwd without the dgrad of the first conv, 2*MAC)
WORKLOADS = {
    'resnet50': dict(
        cfg=_cls('resnet50', 1000, {'name': 'SGD', 'params': {'lr': 0.1, 'weight_decay': 0.0001, 'momentum': 0.9}},
                 in_channels=3),
        batch=256, size=224, classes=1000, flop_per_img=24.30e9, metric=METRIC,
        name='ResNet-50 ClassificationTask synthetic 3x224x224 bs256/GPU bf16 (fp32 master weights, fp32 BN '
             'statistics), SGD momentum step included', oracle=('resnet50', 2048)),
    'resnet18_cifar': dict(
        cfg=_cls('resnet18', 10, {'name': 'Adam', 'params': {'lr': 0.0001}}, in_channels=3),
        batch=128, size=32, classes=10, flop_per_img=0.2173e9,
        metric='images/sec fwd+bwd ResNet-18 32² bs128 (classification_cifar10.yaml); 1/2/4/8 GPU scaling',
        name='ResNet-18 ClassificationTask synthetic 3x32x32 bs128/GPU bf16, Adam step included (C1)',
        oracle=('resnet18', 512)),
    'swin_t': dict(
        cfg=_cls('swinv2_custom', 1000, {'name': 'AdamW', 'params': {'lr': 1e-4, 'weight_decay': 0.05}},
                 img_size=224, window_size=7),
        batch=256, size=224, classes=1000, flop_per_img=26.95e9,
        metric='images/sec fwd+bwd Swin-T(V2) 224² window 7 bs256; 1/2/4/8 GPU scaling',
        name='Swin-T (V2) ClassificationTask synthetic 3x224x224 bs256/GPU bf16, AdamW step included (C3)', oracle=None),
    'hrnet_seg': dict(
        cfg={'task': {'name': 'SegmentationTask', 'params': {
            'backbone_name': 'hrnet_w18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
            'neck_name': 'HRNetSegmentationNeck', 'head_name': 'SegmentationHead', 'head_params': {'num_classes': 19}}},
            'joint_loss': CE, 'optimization': [{'optimizer': {'name': 'SGD', 'params': {'lr': 0.01, 'momentum': 0.9}}}]},
        batch=32, size=512, classes=19, seg=True, flop_per_img=110.32e9,   # scripts/count_flops_hrnet.py
        metric='images/sec fwd+bwd HRNet-W18 + seg neck/head 512² bs32; 1/2/4/8 GPU scaling',
        name='HRNet-W18 + HRNetSegmentationNeck + SegmentationHead synthetic 3x512x512 bs32/GPU bf16, SGD step (C4)',
        oracle=None),
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get('bf16_tflops_sustained', 1400.0), p.get('bf16_tflops', 1590.0), p.get('hbm_gbs', 6650.0), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, reasons, mx = [], set(), None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def _oracle_net(wl):
    from oracle import models as om
    name, width = wl['oracle']
    return om.ClassificationTask(om.resnet(name), om.Pooling(width), om.ClassificationHead(width, wl['classes']))


def _torch_optimizer(wl, params, **kw):
    o = wl['cfg']['optimization'][0]['optimizer']
    return getattr(torch.optim, o['name'])(params, **o['params'], **kw)


def cpu_reference(wl, batch, steps, warmup):
    """The reference's CPU path restated (oracle): same graph, fp32, all host threads."""
    torch.manual_seed(42)
    torch.set_float32_matmul_precision('highest')  # torchok/__main__.py:36
    net = _oracle_net(wl)
    opt = _torch_optimizer(wl, net.parameters())
    x = torch.randn(batch, 3, wl['size'], wl['size'])
    y = torch.randint(0, wl['classes'], (batch,))
    crit = torch.nn.CrossEntropyLoss()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = net.forward_with_gt({'image': x, 'target': y})
        crit(out['prediction'], out['target']).backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3


def gpu_torch_baseline(wl, batch, steps=10, warmup=5):
    """The oracle's torch.nn graph on THIS GPU: channels_last, bf16 autocast, torch.optim (fused) — i.e. cuDNN / cuBLAS,
    the library path the product has to beat.  Not the product path; reported for comparison only."""
    import torch.nn as nn
    torch.manual_seed(42)
    dev = torch.device('cuda')
    net = _oracle_net(wl).to(dev).to(memory_format=torch.channels_last)
    opt = _torch_optimizer(wl, net.parameters(), fused=True)
    x = torch.randn(batch, 3, wl['size'], wl['size'], device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, wl['classes'], (batch,), device=dev)

    class Plain(nn.Module):        # the oracle's precision-policy wrappers are identities outside amp_bf16(): plain torch.nn
        def __init__(self, t):
            super().__init__()
            self.t = t

        def forward(self, x, y):
            return nn.functional.cross_entropy(self.t.forward_with_gt({'image': x, 'target': y})['prediction'], y)
    model = Plain(net)
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = model(x, y)
        loss.backward()
        opt.step()
        return loss
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    # CUDA-graph the torch step too, so the comparison is not about Python launch overhead
    graphed = True
    try:
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=False)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(s)

        def gstep():
            opt.zero_grad(set_to_none=False)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                loss = model(x, y)
            loss.backward()
            opt.step()
        with torch.cuda.graph(g):
            gstep()
        run = g.replay
    except Exception as e:   # capture can fail on some optimizer paths: fall back to the eager loop and say so
        graphed = False
        run = step
        print(f'[bench] torch baseline graph capture failed ({type(e).__name__}); eager loop', file=sys.stderr)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del net, opt, model
    torch.cuda.empty_cache()
    return {'value': batch / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'cuda_graph': graphed,
            'how': f'oracle torch.nn graph .cuda(), channels_last, bf16 autocast, torch.optim fused, bs{batch}, '
                   f'cudnn.benchmark, torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if wl['oracle'] is None:
        print(json.dumps({'impl': 'reference', 'unavailable': f'no CPU arm for workload {args.workload}'}), flush=True)
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch = min(args.cpu_batch, wl['batch'])
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    ips, ms = cpu_reference(wl, batch, steps, warm)
    line = {
        'impl': 'reference', 'metric': wl['metric'], 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{wl["name"]} — CPU sample at bs{batch} (img/s normalised; reference CPU path = oracle '
                               f'restatement of the torch.nn graph)'},
        'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{steps} fwd+bwd+optimizer steps at bs{batch} fp32, torch {torch.__version__} CPU'},
        'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- roofline
class CallTimer:
    """CUDA events around every kernel-launching C-ABI call of one eager step (torchok_b200._lib tracer hook)."""

    def __init__(self):
        self.calls = []

    def before(self, name, a):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        return (name, a, e0)

    def after(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream())
        self.calls.append(tok + (e1,))


def _desc(a):
    d = a._obj
    p = (d.h + 2 * d.pad - d.dil * (d.r - 1) - 1) // d.stride + 1
    q = (d.w + 2 * d.pad - d.dil * (d.s - 1) - 1) // d.stride + 1
    return d, p, q


def _call_work(name, a):
    """(family, algorithmic FLOPs, algorithmic HBM bytes) of one C-ABI call; None when the call is not modelled.
    Bytes = every operand read once and every result written once (DESIGN 4's per-unit figures x units of the launch)."""
    v = lambda x: x is not None and x != 0   # noqa: E731  (nullable pointer argument given?)
    if name in ('tok_conv_fprop', 'tok_conv_fprop_bn'):
        d, p, q = _desc(a[0])
        m = d.n * p * q
        x_elems = m * d.c if (d.r == 1 and d.s == 1) else d.n * d.h * d.w * d.c
        add = m * d.k if (name == 'tok_conv_fprop' and v(a[6])) else 0
        return 'conv fprop', 2.0 * m * d.k * d.r * d.s * d.c, 2.0 * (x_elems + m * d.k + d.k * d.r * d.s * d.c + add)
    if name in ('tok_conv_dgrad', 'tok_conv_dgrad_masked'):
        d, p, q = _desc(a[0])
        m = d.n * p * q
        dx = d.n * d.h * d.w * d.c
        return 'conv dgrad', 2.0 * m * d.k * d.r * d.s * d.c, \
            2.0 * (m * d.k + dx * (2 if v(a[4]) else 1) + d.k * d.r * d.s * d.c)
    if name == 'tok_conv_wgrad':
        d, p, q = _desc(a[0])
        m = d.n * p * q
        x_elems = m * d.c if (d.r == 1 and d.s == 1) else d.n * d.h * d.w * d.c
        return 'conv wgrad', 2.0 * m * d.k * d.r * d.s * d.c, 2.0 * (x_elems + m * d.k) + 4.0 * d.k * d.r * d.s * d.c
    if name in ('tok_linear_fwd', 'tok_linear_dgrad', 'tok_linear_dgrad_add', 'tok_linear_wgrad'):
        m, n, k = a[0], a[1], a[2]
        extra = m * k if name == 'tok_linear_dgrad_add' else 0
        fam = {'tok_linear_fwd': 'linear fwd', 'tok_linear_wgrad': 'linear wgrad'}.get(name, 'linear dgrad')
        return fam, 2.0 * m * n * k, 2.0 * (m * k + m * n + n * k + extra)
    if name == 'tok_stem_conv_fprop' or name == 'tok_stem_conv_wgrad':
        n, h, w, k = a[0], a[1], a[2], a[3]
        p, q = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        return 'stem conv', 2.0 * n * p * q * k * 147, 2.0 * (n * (p + 3) * (q + 3) * 16 + n * p * q * k)
    if name == 'tok_bn_apply':
        rows, c = a[0], a[1]
        return 'bn fwd apply', 0.0, 2.0 * rows * c * (3 if v(a[5]) else 2)
    if name == 'tok_bn_apply_bits':
        rows, c = a[0], a[1]
        return 'bn fwd apply', 0.0, 2.0 * rows * c * 3 + rows * c / 8
    if name in ('tok_bn_bwd_reduce2', 'tok_bn_bwd_reduce2_finalize'):
        rows, c = a[0], a[1]
        return 'bn bwd reduce', 0.0, 2.0 * rows * c * (3 if v(a[3]) else 2) + (rows * c / 8 if v(a[6]) else 0)
    if name == 'tok_bn_bwd_reduce2_finalize_cv':      # (rows, C, c_valid, dout, dout2, y, mask_mode, bits, ...)
        rows, c = a[0], a[1]
        return 'bn bwd reduce', 0.0, 2.0 * rows * c * (3 if v(a[4]) else 2) + (rows * c / 8 if v(a[7]) else 0)
    if name == 'tok_bn_bwd_fused_cv':   # (rows, C, c_valid, dout, dout2, y, mask_mode, bits, ..., dy [23], dres [24])
        rows, c = a[0], a[1]
        n_in = 3 if v(a[4]) else 2
        return 'bn bwd fused (reduce + apply, L2-sized)', 0.0, \
            2.0 * rows * c * (n_in + 1 + (1 if v(a[24]) else 0)) + (rows * c / 8 if v(a[7]) else 0)
    if name in ('tok_bn_apply_chain', 'tok_bn_apply_bits_chain'):
        rows, c = a[0], a[1]
        return 'bn fwd apply', 0.0, 2.0 * rows * c * (3 if v(a[18]) else 2) + (rows * c / 8 if name.endswith('bits_chain') else 0)
    if name == 'tok_bn_bwd_apply2':
        rows, c = a[0], a[1]
        return 'bn bwd apply', 0.0, 2.0 * rows * c * ((3 if v(a[3]) else 2) + 1 + (1 if v(a[13]) else 0)) + \
            (rows * c / 8 if v(a[6]) else 0)
    if name in ('tok_layernorm_fwd', 'tok_layernorm_bwd'):
        return 'layernorm', 0.0, 6.0 * a[0] * a[1]
    if name == 'tok_gelu_fwd':
        return 'gelu', 0.0, 4.0 * a[0]
    if name == 'tok_gelu_bwd':
        return 'gelu', 0.0, 6.0 * a[0]
    if name in ('tok_window_attn_fwd', 'tok_window_attn_bwd'):
        # (batch, h, w, heads, window, shift, head_dim): 4 N^2 d fwd / 10 N^2 d bwd per (window, head)
        b, h, w, heads, ws = a[0], a[1], a[2], a[4], a[5]      # (B, H, W, C, heads, window, shift, ...)
        nwin = b * (h // ws) * (w // ws)
        n2 = (ws * ws) ** 2
        f = (4.0 if name.endswith('fwd') else 10.0) * n2 * 32 * nwin * heads
        tok = b * h * w * heads * 32
        return 'window attention', f, 2.0 * tok * (4 if name.endswith('fwd') else 8)
    return None


def _call_writes(name, a, w):
    """Bytes one C-ABI call WRITES to HBM (subset of w[2]); None when unknown (then half of the bytes is assumed).
    HBM on this part writes at ~3.9 TB/s at best (`hbm_write_gbs`, measured live) while a copy moves 6.5 TB/s, so a
    launch whose traffic is mostly stores (an expansion 1x1: reads M x 64, writes M x 256) has a floor set by its writes."""
    v = lambda x: x is not None and x != 0   # noqa: E731
    if name in ('tok_conv_fprop', 'tok_conv_fprop_bn'):
        d, p, q = _desc(a[0])
        return 2.0 * d.n * p * q * d.k
    if name in ('tok_conv_dgrad', 'tok_conv_dgrad_masked'):
        d, p, q = _desc(a[0])
        return 2.0 * d.n * d.h * d.w * d.c
    if name == 'tok_conv_wgrad':
        d, p, q = _desc(a[0])
        return 4.0 * d.k * d.r * d.s * d.c
    if name == 'tok_linear_fwd':
        return 2.0 * a[0] * a[1]
    if name in ('tok_linear_dgrad', 'tok_linear_dgrad_add'):
        return 2.0 * a[0] * a[2]
    if name == 'tok_linear_wgrad':
        return 4.0 * a[1] * a[2]
    if name in ('tok_bn_apply', 'tok_bn_apply_bits', 'tok_bn_apply_chain', 'tok_bn_apply_bits_chain'):
        return 2.0 * a[0] * a[1] * (1.0625 if 'bits' in name else 1.0)
    if name in ('tok_bn_bwd_reduce2', 'tok_bn_bwd_reduce2_finalize', 'tok_bn_bwd_reduce2_finalize_cv'):
        return 0.0
    if name == 'tok_bn_bwd_apply2':
        return 2.0 * a[0] * a[1] * (2 if v(a[13]) else 1)
    if name in ('tok_gelu_fwd',):
        return 2.0 * a[0]
    if name in ('tok_gelu_bwd',):
        return 2.0 * a[0]
    if name in ('tok_layernorm_fwd', 'tok_layernorm_bwd'):
        return 2.0 * a[0] * a[1]
    return None


def measure_hbm_write_gbs():
    """Write-only HBM bandwidth (2 GiB zero fill, best of 3, CUDA events) — the store-side ceiling of the roofline."""
    buf = torch.empty(1 << 30, dtype=torch.bfloat16, device='cuda')
    best = float('inf')
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        buf.zero_()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del buf
    torch.cuda.empty_cache()
    return (1 << 31) / (best * 1e-3) / 1e9


def step_roofline(loop, batch, sustained_tf, hbm_gbs):
    """Time every C-ABI launch of ONE eager step with CUDA events on the launching stream and aggregate by kernel
    family: share of the step, achieved rate, and time-weighted fraction of max(tensor floor, HBM floor)."""
    from torchok_b200._lib import lib
    L = lib()
    from torchok_b200 import kernels as K
    wgrad_async, K._WGRAD_ASYNC = K._WGRAD_ASYNC, False    # events on the launching stream: keep every launch in line
    use_graph, loop.use_graph = loop.use_graph, False
    snap = loop._snapshot()
    for _ in range(2):
        loop.train_step(batch)
    torch.cuda.synchronize()
    # Two traced steps with the Python garbage collector off; every call keeps the smaller of its two durations.  (One
    # eager step of HRNet holds ~10^5 live tensors; a generation-2 collection between the start event and the launch
    # showed up as a single 37 ms "BatchNorm reduce" in the r2 traces.)
    import gc
    gc_was = gc.isenabled()
    gc.disable()
    traces, step_ms = [], float('inf')
    for _ in range(2):
        timer = CallTimer()
        L.tracer = timer
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        loop.train_step(batch)
        e1.record(torch.cuda.current_stream())
        L.tracer = None
        torch.cuda.synchronize()
        step_ms = min(step_ms, e0.elapsed_time(e1))
        traces.append([(name, a, s.elapsed_time(e)) for name, a, s, e in timer.calls])
    if gc_was:
        gc.enable()
    if len(traces[0]) == len(traces[1]) and all(x[0] == y[0] for x, y in zip(*traces)):
        calls = [(x[0], x[1], min(x[2], y[2])) for x, y in zip(*traces)]
    else:
        calls = traces[1]
    loop._restore(snap)
    loop.use_graph = use_graph
    K._WGRAD_ASYNC = wgrad_async
    fam = {}
    hbm_write_gbs = measure_hbm_write_gbs()
    dump = open(os.environ['TOK_BENCH_CALLS'], 'w') if os.environ.get('TOK_BENCH_CALLS') else None
    for name, a, ms in calls:
        w = _call_work(name, a)
        if dump is not None:    # per-launch table for profiles/: call, shape, measured us, max(tensor, HBM) floor us
            shape = ''
            if name.startswith('tok_conv_'):
                d_, p_, q_ = _desc(a[0])
                shape = f'n{d_.n} {d_.c}x{d_.h}x{d_.w}->{d_.k}x{p_}x{q_} k{d_.r} s{d_.stride}'
            elif w is not None and len(a) > 2 and isinstance(a[0], int) and isinstance(a[1], int):
                shape = f'{a[0]}x{a[1]}' + (f'x{a[2]}' if isinstance(a[2], int) else '')
            fl = max(w[1] / (sustained_tf * 1e12), w[2] / (hbm_gbs * 1e9)) * 1e6 if w else 0.0
            dump.write(f'{name},{shape},{ms * 1e3:.1f},{fl:.1f},{(w[1] if w else 0):.3e},{(w[2] if w else 0):.3e}\n')
        key, flops, byts = w if w else ('other (' + name.replace('tok_', '') + ')', 0.0, 0.0)
        f = fam.setdefault(key, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0, floor_ms=0.0, floor_w_ms=0.0,
                                     modelled=w is not None))
        floor = max(flops / (sustained_tf * 1e12), byts / (hbm_gbs * 1e9)) * 1e3
        wb = _call_writes(name, a, w) if w else None
        wb = byts / 2 if wb is None else wb
        f['floor_w_ms'] += max(floor, wb / (hbm_write_gbs * 1e9) * 1e3)
        f['launches'] += 1
        f['ms'] += ms
        f['flops'] += flops
        f['bytes'] += byts
        f['floor_ms'] += floor
    if dump is not None:
        dump.close()
    total = sum(f['ms'] for f in fam.values())
    groups = {'tcgen05 implicit-GEMM conv/linear (fprop+dgrad+wgrad)':
              [k for k in fam if k.startswith('conv ') or k.startswith('linear ') or k == 'stem conv'],
              'BatchNorm elementwise (apply / backward reduce / backward apply)': [k for k in fam if k.startswith('bn ')],
              'window attention (tcgen05)': [k for k in fam if k == 'window attention'],
              'LayerNorm / GELU elementwise': [k for k in fam if k in ('layernorm', 'gelu')]}
    table = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
        if f['ms'] < 0.005 * total and not f['modelled']:
            continue
        table[k] = {'launches': f['launches'], 'ms': round(f['ms'], 4), 'share': round(f['ms'] / total, 4),
                    'frac_of_floor': round(f['floor_ms'] / f['ms'], 4) if f['modelled'] and f['ms'] > 0 else None,
                    'frac_of_write_aware_floor': round(f['floor_w_ms'] / f['ms'], 4) if f['modelled'] and f['ms'] > 0 else None,
                    'tflops': round(f['flops'] / (f['ms'] * 1e-3) / 1e12, 1) if f['flops'] else None,
                    'gbs': round(f['bytes'] / (f['ms'] * 1e-3) / 1e9, 1) if f['bytes'] else None}
    best, best_ms = None, 0.0
    for g, keys in groups.items():
        ms = sum(fam[k]['ms'] for k in keys)
        if ms > best_ms:
            best, best_ms = g, ms
    keys = groups[best]
    ms = sum(fam[k]['ms'] for k in keys)
    flops = sum(fam[k]['flops'] for k in keys)
    byts = sum(fam[k]['bytes'] for k in keys)
    floor = sum(fam[k]['floor_ms'] for k in keys)
    floor_w = sum(fam[k]['floor_w_ms'] for k in keys)
    tensor_bound = sum(fam[k]['flops'] for k in keys) / (sustained_tf * 1e12) >= byts / (hbm_gbs * 1e9)
    if tensor_bound:
        ach, peak, unit = flops / (ms * 1e-3) / 1e12, sustained_tf, 'TFLOP/s'
    else:
        ach, peak, unit = byts / (ms * 1e-3) / 1e9, hbm_gbs, 'GB/s'
    return {'bound': 'tensor' if tensor_bound else 'hbm', 'achieved': ach, 'peak': peak, 'unit': unit,
            'frac': floor / ms, 'traffic': None,
            'frac_write_aware': floor_w / ms, 'hbm_write_gbs': hbm_write_gbs,
            'frac_write_aware_note': 'same time-weighted fraction with every launch floor = max(tensor, bytes / HBM copy '
                                     'peak, written bytes / write-only HBM bandwidth measured live by a 2 GiB fill): stores '
                                     'alone top out at ~3.9 TB/s on this part, so the store-heavy expansion layers sit '
                                     'closer to what the memory system allows than frac says',
            'kernel': best, 'launches': sum(fam[k]['launches'] for k in keys), 'ms_per_step': ms,
            'share_of_step': ms / total,
            'frac_note': 'time-weighted: sum over the launches of max(FLOPs / sustained bf16 peak, algorithmic bytes / '
                         'HBM peak) divided by the sum of their CUDA-event durations; achieved/peak is the raw rate of '
                         'the family in the unit of its dominant bound',
            'eager_step_ms': step_ms, 'families': table}


def _trace(msg):
    if os.environ.get('TOK_BENCH_TRACE'):
        print(f'[bench rank {os.environ.get("RANK", "0")} {time.strftime("%H:%M:%S")}] {msg}', file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def _init_dist(dev):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        os.environ.setdefault('NCCL_IB_DISABLE', '1')      # NVLink only (north_star)
        os.environ.setdefault('NCCL_P2P_LEVEL', 'NVL')
        os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')
        os.environ['NCCL_DEBUG'] = os.environ.get('TOK_NCCL_DEBUG', 'NONE')   # NCCL prints its version banner to STDOUT at VERSION and above
        dist.init_process_group('nccl', device_id=dev)
        _trace('process group up')
    return world


def run_retrieval(args):
    """C5: IndexBasedMeter's search over N x 512 vectors; with N ranks: one all-gather of the vectors, each rank
    searches its N/world query rows (torchok/metrics/index_base_metric.py:112-120,170-270)."""
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    world = _init_dist(dev)
    from torchok_b200._lib import lib
    from torchok_b200.metrics import index_base_metric as ibm
    n, d, k = args.retrieval_n, 512, 1
    per = n // world
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    mine = ibm.normalize_rows(torch.randn(per, d, device=dev, generator=g))

    def search():
        if world > 1:
            allv = torch.empty(world * per, d, device=dev)
            dist.all_gather_into_tensor(allv, mine)
        else:
            allv = mine
        return ibm.search_topk(mine, allv, k + 1)
    n0 = lib().launches
    s, i = search()
    per_call = lib().launches - n0
    torch.cuda.synchronize()
    ok = bool((i[:, 0] == torch.arange(per, device=dev) + rank * per).float().mean() > 0.999)
    steps = max(1, min(args.steps, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        search()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / steps
    if rank == 0:
        sustained, burst, hbm, src = peaks()
        flop = 2.0 * n * n * d
        tf = flop / (ms * 1e-3) / 1e12 / world
        print(json.dumps({
            'metric': f'queries/sec retrieval cosine top-{k + 1} N={n} D={d}', 'value': n / (ms * 1e-3), 'unit': 'queries/s',
            'n_gpus': world, 'steps': steps, 'warmup': 1, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'IndexBasedMeter search, {n} x {d} unit vectors, top-{k + 1}, query rows sharded over '
                                   f'{world} rank(s), one all-gather of the vectors per search (C5)',
                       'self_is_top1': ok},
            'gpu_launches': per_call * steps,
            'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': sustained, 'unit': 'TFLOP/s', 'frac': tf / sustained,
                         'traffic': None, 'kernel': 'cosine_topk_kernel (tcgen05 GEMM + fused running top-k), per GPU',
                         'peak_source': f'MEASURED_PEAKS.json bf16_tflops_sustained ({src})'}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args):
    if args.workload == 'retrieval':
        return run_retrieval(args)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    world = _init_dist(dev)
    import torchok_b200 as tb
    from torchok_b200._lib import lib
    from torchok_b200.engine import StreamLoop

    torch.manual_seed(42)
    cfg = tb.load_config(wl['cfg'])
    task = tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params).to(dev)
    loop = StreamLoop(task, use_graph=not args.no_graph)
    _trace('StreamLoop built')
    B = args.batch or wl['batch']
    size, classes, seg = wl['size'], wl['classes'], wl.get('seg', False)
    tshape = (B, size, size) if seg else (B,)
    if args.profile_step:
        loop.use_graph = False
        x = torch.randn(B, 3, size, size, device=dev)
        y = torch.randint(0, classes, tshape, device=dev)
        for _ in range(3):
            loop.train_step({'image': x, 'target': y})
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        loop.train_step({'image': x, 'target': y})
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    g = torch.Generator().manual_seed(42 + rank)
    # host batches as the reference's datasets deliver them with `input_dtype: float16`
    # (examples/configs/classification_cifar10.yaml:13,44): half-precision images, int64 targets, pinned
    host_img = [torch.randn(B, 3, size, size, generator=g).to(torch.bfloat16).pin_memory() for _ in range(2)]
    host_tgt = [torch.randint(0, classes, tshape, generator=g).pin_memory() for _ in range(2)]
    dev_batch = {'image': host_img[0].to(dev), 'target': host_tgt[0].to(dev)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches per step, counted on the eager warm-up inside the first train_step
    n0 = lib().launches
    loop.train_step(dev_batch)
    _trace('first train_step (warm-up + capture) done')
    per_step = (lib().launches - n0) // (loop.warmup + (1 if loop.use_graph else 0)) if loop.use_graph else \
        (lib().launches - n0)
    for _ in range(max(args.warmup, 3) - 1):
        loop.train_step(dev_batch)
    barrier()

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for _ in range(args.steps):
        loss = loop.train_step(dev_batch)
    e1.record(st)
    barrier()
    ms_total = e0.elapsed_time(e1)
    _trace(f'timed region done: {ms_total / args.steps:.2f} ms/step')
    clocks = sampler.stop() if sampler else None
    loss_val = float(loss)

    # ---- e2e: pinned host buffers in, loss out, every step ------------------------------------------------------
    copy_stream = torch.cuda.Stream()
    stage = [{'image': torch.empty_like(dev_batch['image']), 'target': torch.empty_like(dev_batch['target'])}
             for _ in range(2)]
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            stage[s]['image'].copy_(host_img[s], non_blocking=True)
            stage[s]['target'].copy_(host_tgt[s], non_blocking=True)
            ready[s].record(copy_stream)

    for ev in consumed:
        ev.record(st)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(st)
    upload(0)
    for i in range(args.steps):
        s = i % 2
        if i + 1 < args.steps:
            upload(i + 1)                       # H2D of the next batch overlaps this step's compute
        st.wait_event(ready[s])
        loss = loop.train_step(stage[s])        # public API: task.training_step + optimizer inside the stream loop
        consumed[s].record(st)
        loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)   # D2H of the step's result
    e3.record(st)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    _trace('e2e region done')
    assert bool(torch.isfinite(loss_host).all()), 'non-finite loss in the e2e run'

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()

    if rank == 0:
        sustained, burst, hbm, src = peaks()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        h2d = host_img[0].numel() * host_img[0].element_size() + host_tgt[0].numel() * host_tgt[0].element_size()
        roof = step_rf = cpu = torch_gpu = None
        if world == 1:   # single-GPU legs: per-launch roofline of one eager step, CPU baseline, cuDNN baseline
            roof = step_roofline(loop, dev_batch, sustained, hbm)
            roof['peak_source'] = f'MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs ({src})'
            if not args.skip_cpu and wl['oracle'] is not None:
                cores = os.cpu_count()
                torch.set_num_threads(cores)
                cb = min(args.cpu_batch, B)
                ips, ms = cpu_reference(wl, cb, 3, 1)
                cpu = {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                       'sample': f'3 fwd+bwd+optimizer steps of the oracle at bs{cb} fp32 ({ms:.0f} ms/step), img/s normalised'}
            if not args.skip_torch and wl['oracle'] is not None:
                try:
                    torch_gpu = gpu_torch_baseline(wl, B)
                except Exception as e:   # the comparison leg must never cost the bench line
                    torch_gpu = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
        if wl['flop_per_img']:
            step_tf = wl['flop_per_img'] * B / (ms_step * 1e-3) / 1e12
            step_rf = {'bound': 'tensor', 'achieved': step_tf, 'peak': sustained, 'unit': 'TFLOP/s',
                       'frac': step_tf / sustained,
                       'note': f'whole step: {wl["flop_per_img"] / 1e9:.2f} GFLOP/img x {B} img / step time'}
        line = {
            'metric': wl['metric'], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'global_batch': world * B, 'parallelism': f'dp{world}',
                       'l2': 'activations of one step far exceed the 126 MB L2; no explicit flush'
                             if args.workload != 'resnet18_cifar' else
                             'whole working set of a step fits the 126 MB L2 (launch/latency-bound workload)',
                       'cuda_graph': bool(loop.use_graph), 'grad_exchange': loop.exchange, 'final_loss': loss_val},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(per_step) * args.steps,
            'gpu_launches_note': 'C-ABI kernel-launching calls per step x steps (each call launches >= 1 kernel)',
            'clocks': clocks,
            'roofline': roof,
            'roofline_step': step_rf,
            'cpu_baseline': cpu,
            'gpu_torch_baseline': torch_gpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        loop.close()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='resnet50', choices=sorted(WORKLOADS) + ['retrieval'])
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s BASELINE.json batch)')
    ap.add_argument('--retrieval-n', type=int, default=1 << 20)
    ap.add_argument('--cpu-batch', type=int, default=32)
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-torch', action='store_true', help='skip the cuDNN (torch.nn on this GPU) comparison leg')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--profile-step', action='store_true',
                    help='ncu aid: warm up eagerly, then run ONE eager step between cudaProfilerStart/Stop and exit '
                         '(use with ncu --profile-from-start off); prints no bench line')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference)')
        run_ours(args)


if __name__ == '__main__':
    main()
