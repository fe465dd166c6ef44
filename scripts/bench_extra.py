#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 3-5), one JSON line each — NOT the driver's bench contract (that is
bench.py / ResNet-50); used to fill DESIGN.md / profiles with measured numbers for the other §8 rows.

    python scripts/bench_extra.py retrieval [N] [D] [k]     # N x N cosine top-k, IndexBasedMeter search path
    python scripts/bench_extra.py swin [batch]              # Swin-T(V2) w7 @224 ClassificationTask fwd+bwd+AdamW step
    python scripts/bench_extra.py hrnet [batch] [size]      # HRNet-W18 + seg neck/head SegmentationTask step
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import torchok_b200 as tb  # noqa: E402
from torchok_b200.engine import StreamLoop  # noqa: E402
from torchok_b200.metrics import index_base_metric as ibm  # noqa: E402


def timed(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def retrieval(n=1_000_000, d=512, k=1):
    g = torch.Generator(device='cuda').manual_seed(42)
    v = ibm.normalize_rows(torch.randn(n, d, device='cuda', generator=g))
    res = {}

    def run():
        res['s'], res['i'] = ibm.search_topk(v, v, k + 1)
    ms = timed(run, warm=1, iters=2)
    ok = bool((res['i'][:, 0] == torch.arange(n, device='cuda')).float().mean() > 0.999)
    flop = 2.0 * n * n * d
    print(json.dumps({'workload': f'retrieval cosine top-{k + 1} N={n} D={d} (IndexBasedMeter search)', 'ms': ms,
                      'queries_per_s': n / (ms * 1e-3), 'tflops': flop / (ms * 1e-3) / 1e12,
                      'self_is_top1': ok}), flush=True)


def _step_bench(cfg, batch, size, classes, seg=False, steps=10):
    torch.manual_seed(42)
    task = tb.TASKS.get(cfg['task']['name'])(tb.load_config(cfg), **cfg['task']['params']).cuda()
    loop = StreamLoop(task, use_graph=os.environ.get('TOK_EXTRA_GRAPH', '1') == '1')
    x = torch.randn(batch, 3, size, size, device='cuda')
    y = torch.randint(0, classes, (batch, size, size) if seg else (batch,), device='cuda')
    b = {'image': x, 'target': y}
    if os.environ.get('TOK_EXTRA_PROFILE'):   # ncu --profile-from-start off: one eager step inside the profiler range
        loop.use_graph = False
        for _ in range(2):
            loop.train_step(b)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        loop.train_step(b)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0.0, float(loop.loss)
    if os.environ.get('TOK_EXTRA_TORCHPROF'):  # attribute the non-tok helper kernels (aten copies / fills / adds) to shapes
        from torch.profiler import ProfilerActivity, profile
        loop.use_graph = False
        for _ in range(2):
            loop.train_step(b)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
            loop.train_step(b)
            torch.cuda.synchronize()
        print(prof.key_averages(group_by_input_shape=True).table(sort_by='device_time_total', row_limit=45,
                                                                 max_name_column_width=48, max_shapes_column_width=60))
        print(prof.key_averages(group_by_stack_n=4).table(sort_by='device_time_total', row_limit=30,
                                                          max_name_column_width=40, max_src_column_width=90))
        return 0.0, float(loop.loss)
    ms = timed(lambda: loop.train_step(b), warm=3, iters=steps)
    return ms, float(loop.loss)


def swin(batch=128):
    cfg = {'task': {'name': 'ClassificationTask', 'params': {
        'backbone_name': 'swinv2_custom', 'backbone_params': {'pretrained': False, 'img_size': 224, 'window_size': 7},
        'pooling_name': 'Pooling', 'head_name': 'ClassificationHead', 'head_params': {'num_classes': 1000}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]},
        'optimization': [{'optimizer': {'name': 'AdamW', 'params': {'lr': 1e-4, 'weight_decay': 0.05}}}]}
    ms, loss = _step_bench(cfg, batch, 224, 1000)
    print(json.dumps({'workload': f'Swin-T(V2) w7 224 ClassificationTask bs{batch} fwd+bwd+AdamW (StreamLoop)', 'ms_per_step': ms,
                      'img_per_s': batch / (ms * 1e-3), 'tflops': 26.95e9 * batch / (ms * 1e-3) / 1e12, 'loss': loss}),
          flush=True)


def hrnet(batch=32, size=512):
    cfg = {'task': {'name': 'SegmentationTask', 'params': {
        'backbone_name': 'hrnet_w18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
        'neck_name': 'HRNetSegmentationNeck', 'head_name': 'SegmentationHead', 'head_params': {'num_classes': 19}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]},
        'optimization': [{'optimizer': {'name': 'SGD', 'params': {'lr': 0.01, 'momentum': 0.9}}}]}
    ms, loss = _step_bench(cfg, batch, size, 19, seg=True, steps=5)
    print(json.dumps({'workload': f'HRNet-W18 + HRNetSegmentationNeck + SegmentationHead {size}x{size} bs{batch} step (StreamLoop)',
                      'ms_per_step': ms, 'img_per_s': batch / (ms * 1e-3), 'loss': loss}), flush=True)


if __name__ == '__main__':
    what = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    t0 = time.time()
    {'retrieval': retrieval, 'swin': swin, 'hrnet': hrnet}[what](*args)
    print(f'# {what} done in {time.time() - t0:.1f}s', file=sys.stderr)
