#!/usr/bin/env python
"""Multi-GPU parity checks, run under torchrun by tests/test_ddp_gpu.py (one process per GPU, NCCL for the plumbing).

    torchrun --nproc-per-node 2 tests/gpu/ddp_check.py step   [peer|nccl]   # N-rank step == 1-rank large-batch step
    torchrun --nproc-per-node 2 tests/gpu/ddp_check.py retrieval            # sharded search == single-GPU search

`step`: every rank trains on its shard of one seeded global batch through StreamLoop (gradient exchange over NVLink peer
memory fused with the optimizer, or the NCCL all-reduce path) and compares the updated fp32 master weights with the
expectation computed WITHOUT any exchange: the gradients of every shard from a world-size-1 loop on this very GPU,
averaged, then torch.optim.SGD's update rule (DDP semantics with local BatchNorm statistics: the reference's
`strategy: ddp`, `sync_batchnorm: False`, torchok/constructor/config_structure.py:137-140,170).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build_task(tb, seed=0):
    torch.manual_seed(seed)
    cfg = tb.load_config({
        'task': {'name': 'ClassificationTask', 'params': {
            'backbone_name': 'resnet18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
            'pooling_name': 'Pooling', 'head_name': 'ClassificationHead', 'head_params': {'num_classes': 10}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]}})
    task = tb.TASKS.get('ClassificationTask')(cfg, **cfg.task.params)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():       # de-degenerate: zero_init_last would switch every residual branch off
        for m in task.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return task


def check_step(mode):
    os.environ['TOK_DDP'] = mode
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    import torchok_b200 as tb
    from torchok_b200.engine import StreamLoop
    opt = {'name': 'SGD', 'params': {'lr': 0.1, 'momentum': 0.9, 'weight_decay': 1e-4}}
    per = 32
    g = torch.Generator().manual_seed(123)
    x = torch.randn(world * per, 3, 64, 64, generator=g).to(torch.bfloat16).float()
    y = torch.randint(0, 10, (world * per,), generator=g)
    shard = lambda r: {'image': x[r * per:(r + 1) * per].to(dev), 'target': y[r * per:(r + 1) * per].to(dev)}  # noqa: E731

    # expectation: per-shard gradients from a world-size-1 loop (no exchange), averaged, one SGD step from w0
    ref_task = build_task(tb).to(dev)
    ref = StreamLoop(ref_task, optimizer=opt, use_graph=False, distributed=False)
    w0 = ref.arena.master.clone()
    state0 = [b.clone() for b in ref_task.buffers()]
    grads = []
    for r in range(world):
        for b, s in zip(ref_task.buffers(), state0):
            b.copy_(s)
        ref.arena.zero_grad()
        ref.arena.begin_step()
        ref_task.train()
        ref_task.training_step(shard(r))['loss'].backward()
        torch.cuda.synchronize()
        grads.append(ref.arena.grad.clone())
    gmean = sum(grads) / world
    # noise floor of the comparison: the same shard twice on this GPU (BatchNorm statistics and weight gradients are
    # accumulated with fp32 atomics, so two runs differ by summation order and the occasional bf16 rounding it flips)
    for b, s in zip(ref_task.buffers(), state0):
        b.copy_(s)
    ref.arena.zero_grad()
    ref.arena.begin_step()
    ref_task.training_step(shard(world - 1))['loss'].backward()
    torch.cuda.synchronize()
    noise = float((ref.arena.grad - grads[-1]).norm() / grads[-1].norm())
    d = gmean + 1e-4 * w0
    expected1 = w0 - 0.1 * d                                  # first step: momentum buffer = d
    # the N-rank loop
    task = build_task(tb).to(dev)
    loop = StreamLoop(task, optimizer=opt, use_graph=False, bucket_mb=8.0)
    assert loop.exchange == ('peer-fused' if mode == 'peer' else 'nccl'), loop.exchange
    assert torch.equal(loop.arena.master, w0)
    loss = loop.train_step(shard(rank))
    torch.cuda.synchronize()
    got = loop.arena.master
    # compare the UPDATES (w - w0): BatchNorm's atomically accumulated statistics make two runs of the same forward differ
    # by an occasional bf16 rounding, so the bar is 1e-2 in the L2 norm of the update and 5e-2 of its maximum
    du, du_exp = got - w0, expected1 - w0
    err = float((du - du_exp).abs().max() / du_exp.abs().max())
    upd = float((du - du_exp).norm() / du_exp.norm())
    # identical replicas: every rank must hold bit-identical weights and shadows
    flat = torch.stack([got.double().sum(), got.double().abs().sum(), loop.arena.shadow.double().sum()])
    allf = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(allf, flat)
    same = all(torch.equal(allf[0], t) for t in allf)
    shadow_ok = torch.equal(loop.arena.shadow, got.to(torch.bfloat16))
    grads_cleared = float(loop.arena.grad.abs().max()) == 0.0
    print(f'[rank {rank}] {mode}: loss {float(loss):.4f} update vs expectation: max {err:.3e}, l2 {upd:.3e} '
          f'replicas identical: {same} shadow == bf16(master): {shadow_ok} grads cleared: {grads_cleared}', flush=True)
    print(f'[rank {rank}] run-to-run noise of one shard gradient (l2): {noise:.3e}', flush=True)
    assert upd < max(1e-2, 4 * noise) and err < max(5e-2, 20 * noise), (err, upd, noise)
    assert same and shadow_ok and grads_cleared
    # graph replay at world size > 1 (peer path: plain kernel launches): three more steps stay finite and identical
    if mode == 'peer':
        loop.use_graph = True
        for _ in range(3):
            loss = loop.train_step(shard(rank))
        torch.cuda.synchronize()
        assert loop.graph is not None and bool(torch.isfinite(loss))
        flat = torch.stack([loop.arena.master.double().sum(), loop.arena.master.double().abs().sum()])
        allf = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(allf, flat)
        assert all(torch.equal(allf[0], t) for t in allf)
        # ZeRO-1 state gather: the assembled momentum buffer is the same on every rank and non-trivial
        st = loop.optimizer.state_dict(task)
        tot = sum(float(v.double().abs().sum()) for v in st['state']['buf'].values())
        tt = torch.tensor([tot], device=dev, dtype=torch.float64)
        al = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(al, tt)
        assert tot > 0 and all(torch.equal(al[0], t) for t in al)
        print(f'[rank {rank}] peer: graph replay x3 ok, loss {float(loss):.4f}, gathered momentum |sum| {tot:.4e}', flush=True)
    loop.close()


def check_exchange(mode):
    """The exchange + optimizer arithmetic alone, on KNOWN gradients (no backward pass, hence no bf16 / atomics noise):
    rank r's gradient arena is filled from a generator seeded with r, every rank can rebuild all of them, and two SGD
    steps (momentum, weight decay) must reproduce torch's update rule applied to the rank-ordered mean to fp32 round-off."""
    os.environ['TOK_DDP'] = mode
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    import torchok_b200 as tb
    from torchok_b200.engine import StreamLoop
    lr, mu, wd = 0.1, 0.9, 1e-4
    task = build_task(tb).to(dev)
    loop = StreamLoop(task, optimizer={'name': 'SGD', 'params': {'lr': lr, 'momentum': mu, 'weight_decay': wd}},
                      use_graph=False, bucket_mb=4.0)
    a = loop.arena
    w = a.master.clone()
    buf = None
    for step in range(2):
        gs = []
        for r in range(world):
            g = torch.Generator(device=dev).manual_seed(1000 * step + r)
            gs.append(torch.randn(a.numel, device=dev, generator=g))
        a.grad.copy_(gs[rank])
        a.begin_step()
        if loop.peer is not None:
            loop.peer.begin_step()
        a.finish()                       # every bucket: exchange (+ fused optimizer on the peer path)
        if loop.peer is None:
            loop.optimizer.step()
        torch.cuda.synchronize()
        gsum = gs[0].clone()
        for r in range(1, world):        # rank order, like the kernel
            gsum += gs[r]
        d = gsum * (1.0 / world) + wd * w
        buf = d.clone() if step == 0 else mu * buf + d
        w = w - lr * buf
        err = float((a.master - w).abs().max() / w.abs().max())
        # padding elements between parameters are not parameters: compare parameter slices only
        worst = 0.0
        for p_, off in zip(a.params, a.offsets):
            n = p_.numel()
            worst = max(worst, float((a.master[off:off + n] - w[off:off + n]).abs().max()))
        print(f'[rank {rank}] {mode} exchange step {step}: max |w - expected| = {worst:.3e} (whole arena rel {err:.3e}), '
              f'grads cleared: {float(a.grad.abs().max()) == 0.0}', flush=True)
        assert worst < 2e-6, worst
        assert float(a.grad.abs().max()) == 0.0
        assert torch.equal(a.shadow, a.master.to(torch.bfloat16))
    loop.close()


def check_retrieval():
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    from torchok_b200.metrics import index_base_metric as ibm
    n, d, k = 32768, 512, 4
    g = torch.Generator().manual_seed(5)
    v = ibm.normalize_rows(torch.randn(n, d, generator=g).to(dev))
    per = n // world
    mine = v[rank * per:(rank + 1) * per].contiguous()
    allv = torch.empty(n, d, device=dev)
    dist.all_gather_into_tensor(allv, mine)                   # the one exchange of the metric (index_base_metric.py:112-120)
    assert torch.equal(allv, v)
    s, i = ibm.search_topk(mine, allv, k)
    s1, i1 = ibm.search_topk(v, v, k)                         # single-GPU search of everything
    same = torch.equal(i, i1[rank * per:(rank + 1) * per]) and torch.equal(s, s1[rank * per:(rank + 1) * per])
    print(f'[rank {rank}] retrieval: sharded rows == single-GPU rows: {same}', flush=True)
    assert same


if __name__ == '__main__':
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ.setdefault('NCCL_IB_DISABLE', '1')
    os.environ.setdefault('NCCL_P2P_LEVEL', 'NVL')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    what = sys.argv[1]
    if what == 'step':
        check_step(sys.argv[2] if len(sys.argv) > 2 else 'peer')
    elif what == 'exchange':
        check_exchange(sys.argv[2] if len(sys.argv) > 2 else 'peer')
    else:
        check_retrieval()
    dist.barrier()
    dist.destroy_process_group()
