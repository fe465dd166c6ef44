#!/usr/bin/env python
"""bench.py — images/sec/GPU, forward+backward(+optimizer step), ResNet-50 224x224 bs256 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic batch: ClassificationTask(resnet50 -> Pooling ->
ClassificationHead(1000)).forward_with_gt -> JointLoss(CrossEntropyLoss) -> backward -> (bucketed NCCL all-reduce) ->
SGD(momentum 0.9, wd 1e-4) step, the step the reference's Lightning loop runs for
examples/configs/classification_imagenet.yaml (torchok/tasks/base.py:125-133).  Synthetic N(0,1) images, random
labels, random-init weights (seed 42).

Output: ONE JSON line on rank 0 (see the driver contract).  `value` = device-resident inputs, whole-job img/s;
`e2e` = same loop fed from pinned HOST buffers through the public API (H2D of every batch and D2H of every loss inside
the timed region); `roofline` = the dominant kernel (tcgen05 implicit-GEMM conv) timed live with CUDA events;
`cpu_baseline` = the CPU oracle (oracle/models.py, the torch.nn graph the reference dispatches) on the host cores.
`--impl reference` times that CPU path alone on the same config/metric.
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
import torch.distributed as dist  # noqa: E402

METRIC = 'images/sec/GPU fwd+bwd ResNet-50 224² bs256; 1/2/4/8 GPU scaling'
UNIT = 'img/s'
FLOP_PER_IMG = 24.30e9  # fwd 8.178 + bwd (no dgrad of the first conv), FlopCounterMode, 2*MAC
NUM_CLASSES = 1000


def task_config(model='resnet50'):
    """The task/loss/optimizer blocks of examples/configs/classification_imagenet.yaml with backbone resnet50."""
    return {
        'task': {'name': 'ClassificationTask', 'params': {
            'backbone_name': model, 'backbone_params': {'pretrained': False, 'in_channels': 3},
            'pooling_name': 'Pooling', 'head_name': 'ClassificationHead',
            'head_params': {'num_classes': NUM_CLASSES}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss',
                                   'mapping': {'input': 'prediction', 'target': 'target'}}]},
        'optimization': [{'optimizer': {'name': 'SGD',
                                        'params': {'lr': 0.1, 'weight_decay': 0.0001, 'momentum': 0.9}}}],
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
def cpu_reference(batch, steps, warmup, model='resnet50', size=224):
    """The reference's CPU path restated (oracle): same graph, fp32, all host threads."""
    from oracle import models as om
    torch.manual_seed(42)
    torch.set_float32_matmul_precision('highest')  # torchok/__main__.py:36
    net = om.ClassificationTask(om.resnet(model), om.Pooling(2048 if model == 'resnet50' else 512),
                                om.ClassificationHead(2048 if model == 'resnet50' else 512, NUM_CLASSES))
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(batch, 3, size, size)
    y = torch.randint(0, NUM_CLASSES, (batch,))
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


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch = args.cpu_batch
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    ips, ms = cpu_reference(batch, steps, warm)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'ResNet-50 ClassificationTask synthetic 3x224x224, CPU sample at bs{batch} '
                               f'(img/s normalised; reference CPU path = oracle restatement of the torch.nn graph)'},
        'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{steps} fwd+bwd+SGD steps at bs{batch} fp32, torch {torch.__version__} CPU'},
        'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def time_dominant_kernel(batch, sustained_tf):
    """roofline: the tcgen05 implicit-GEMM conv on ResNet-50's heaviest layer shape (3x3 256->256 @14x14, 5 layers,
    1.156 GFLOP/img fwd — SURVEY §8a census), timed alone with CUDA events on the launching stream."""
    from torchok_b200 import kernels as K
    dev = torch.device('cuda')
    n, h, c, k = batch, 14, 256, 256
    d, p, q = K.conv_desc(n, h, h, c, k, 3, 3, 1, 1, 1)
    x = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
    w = torch.randn(k, 3, 3, c, device=dev).to(torch.bfloat16)
    y = torch.empty(n, p, q, k, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, k, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(3):
        K.conv_fprop(d, x, w, y, stats)
    reps, total = 20, 0.0
    for _ in range(reps):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        K.conv_fprop(d, x, w, y, stats)
        e1.record(st)
        e1.synchronize()
        total += e0.elapsed_time(e1)
    ms = total / reps
    flops = 2.0 * n * p * q * k * 9 * c
    ach = flops / (ms * 1e-3) / 1e12
    return {'bound': 'tensor', 'achieved': ach, 'peak': sustained_tf, 'unit': 'TFLOP/s', 'frac': ach / sustained_tf,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from profiles/r1_roofline_kernel_full.md
            # (ncu --set full, scripts/roofline_kernel.py): 26.94 MB read + 0.11 MB written — the 25.7 MB output tile
            # stream stays in the 126 MB L2; algorithmic bytes are 25.7 (x) + 25.7 (y) + 1.2 (w) MB.
            'traffic': 27.05e6 if n == 256 else None, 'traffic_unit': 'B',
            'kernel': 'conv_fwd_persist_kernel<256,3,0,1,0> fprop 3x3 256->256 @14x14',
            'ms_per_launch': ms, 'flop_per_launch': flops}


def _trace(msg):
    if os.environ.get('TOK_BENCH_TRACE'):
        print(f'[bench rank {os.environ.get("RANK", "0")} {time.strftime("%H:%M:%S")}] {msg}', file=sys.stderr, flush=True)


def run_ours(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('TOK_GRAPH_DDP', '0') == '1':    # experimental: NCCL all-reduce captured in the step graph
            os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')
            os.environ.setdefault('NCCL_ASYNC_ERROR_HANDLING', '0')
        os.environ.setdefault('NCCL_IB_DISABLE', '1')      # NVLink only (north_star)
        os.environ.setdefault('NCCL_P2P_LEVEL', 'NVL')
        dist.init_process_group('nccl', device_id=dev)
        _trace('process group up')
    import torchok_b200 as tb
    from torchok_b200._lib import lib
    from torchok_b200.engine import StreamLoop

    torch.manual_seed(42)
    cfg = tb.load_config(task_config(args.model))
    task = tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params).to(dev)
    loop = StreamLoop(task, use_graph=not args.no_graph)
    _trace('StreamLoop built')
    B = args.batch
    if args.profile_step:
        loop.use_graph = False
        x = torch.randn(B, 3, args.size, args.size, device=dev)
        y = torch.randint(0, NUM_CLASSES, (B,), device=dev)
        for _ in range(3):
            loop.train_step({'image': x, 'target': y})
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        loop.train_step({'image': x, 'target': y})
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    g = torch.Generator().manual_seed(42 + rank)
    host_img = [torch.randn(B, 3, args.size, args.size, generator=g).pin_memory() for _ in range(2)]
    host_tgt = [torch.randint(0, NUM_CLASSES, (B,), generator=g).pin_memory() for _ in range(2)]
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
        roof = time_dominant_kernel(B, sustained)
        roof['peak_source'] = f'MEASURED_PEAKS.json bf16_tflops_sustained ({src})'
        step_tf = FLOP_PER_IMG * B / (ms_step * 1e-3) / 1e12
        h2d = host_img[0].numel() * host_img[0].element_size() + host_tgt[0].numel() * host_tgt[0].element_size()
        cpu = None
        if not args.skip_cpu and world == 1:     # the CPU baseline is an N=1 leg (rank 0 only; the other ranks would idle)
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            ips, ms = cpu_reference(args.cpu_batch, 3, 1)
            cpu = {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'3 fwd+bwd+SGD steps of the oracle at bs{args.cpu_batch} fp32 ({ms:.0f} ms/step), '
                             f'img/s normalised'}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'ResNet-50 ClassificationTask synthetic 3x{args.size}x{args.size} bs{B}/GPU bf16 '
                                   f'(fp32 master weights, fp32 BN statistics), SGD momentum step included',
                       'global_batch': world * B, 'parallelism': f'dp{world}',
                       'l2': 'inputs/activations per step (>5 GB) far exceed the 126 MB L2; no explicit flush',
                       'cuda_graph': bool(loop.use_graph), 'final_loss': loss_val},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(per_step) * args.steps,
            'gpu_launches_note': 'C-ABI kernel-launching calls per step x steps (each call launches >= 1 kernel)',
            'clocks': clocks,
            'roofline': roof,
            'roofline_step': {'bound': 'tensor', 'achieved': step_tf, 'peak': sustained, 'unit': 'TFLOP/s',
                              'frac': step_tf / sustained,
                              'note': f'whole step: {FLOP_PER_IMG / 1e9:.2f} GFLOP/img x {B} img / step time'},
            'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() was seen to block after graph-captured collectives,
        # and a benchmark process has nothing left to clean up.  All ranks meet at a barrier first.
        sys.stdout.flush()
        sys.stderr.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--size', type=int, default=224)
    ap.add_argument('--model', default='resnet50')
    ap.add_argument('--cpu-batch', type=int, default=32)
    ap.add_argument('--skip-cpu', action='store_true')
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
