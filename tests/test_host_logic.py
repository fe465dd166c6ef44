"""CPU tests of the host-side mirror of the reference's plug-in surface: registries, JointLoss known answers, config
loading, task assembly (no kernel is launched: module construction and error behaviour only), and the C ABI symbol
table of libtokb200.so."""
import ctypes
import os

import pytest
import torch
from torch.nn import Module

import torchok_b200 as tb
from torchok_b200._lib import LIB_PATH, header_symbols
from torchok_b200.constructor.registry import Registry
from torchok_b200.losses.base import JointLoss


# ---------------------------------------------------------------------------------------------- Registry
def test_registry_semantics():
    """torchok/constructor/registry.py:45-99: KeyError text, duplicate/not-callable errors, containment."""
    r = Registry('things')

    @r.register_class
    def alpha():
        return 1

    assert r.get('alpha') is alpha and r['alpha'] is alpha and 'alpha' in r
    with pytest.raises(KeyError, match='beta is not in the things registry'):
        r.get('beta')
    with pytest.raises(KeyError):
        r.register_class(alpha)
    with pytest.raises(TypeError):
        r.register_class(3)
    assert 'alpha' in r.list_models()


def test_the_fourteen_registries_exist_with_reference_names():
    for name in ('DATASETS', 'TRANSFORMS', 'OPTIMIZERS', 'SCHEDULERS', 'LOSSES', 'METRICS', 'CALLBACKS', 'TASKS',
                 'BACKBONES', 'POOLINGS', 'HEADS', 'NECKS', 'DETECTION_NECKS', 'SAMPLERS'):
        assert isinstance(getattr(tb, name), Registry), name
    for n in ('resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152'):
        assert n in tb.BACKBONES
    for n in ('Pooling', 'PoolingLinear'):
        assert n in tb.POOLINGS
    for n in ('LinearHead', 'ClassificationHead'):
        assert n in tb.HEADS
    assert 'ClassificationTask' in tb.TASKS and 'CrossEntropyLoss' in tb.LOSSES


# ---------------------------------------------------------------------------------------------- JointLoss KATs
class Loss1(Module):
    def forward(self, input, target):
        return torch.abs(input * 10. - target)


class Loss2(Module):
    def forward(self, input, target):
        return torch.abs(input * 20. - target)


def _joint(weights):
    return JointLoss(losses=[Loss1(), Loss2()], tags=['loss1', 'loss2'],
                     mappings=[{'input': 'x', 'target': 'y'}] * 2, weights=weights)


def test_joint_loss_known_answers():
    """tests/base_tests/losses/test_base_losses.py:19-77 of the reference: 8.0 / 10.0 / tagged 5 and 15."""
    x, y = torch.ones(1), torch.full((1,), 5.)
    total, tagged = _joint([0.7, 0.3]).forward(x=x, y=y)
    torch.testing.assert_close(total, torch.tensor([8.]))
    torch.testing.assert_close(tagged['loss1'], torch.tensor([5.]))
    torch.testing.assert_close(tagged['loss2'], torch.tensor([15.]))
    total, _ = _joint([None, None]).forward(x=x, y=y)
    torch.testing.assert_close(total, torch.tensor([10.]))
    with pytest.raises(ValueError):
        _joint([0.7, None])
    j = _joint([0.7, 0.3])
    assert isinstance(j['loss1'], Loss1)
    with pytest.raises(KeyError):
        j['nope']
    with pytest.raises(ValueError):
        j.forward(x=x)  # mapped output missing (losses/base.py:104-113)


# ---------------------------------------------------------------------------------------------- config + task
CFG = {
    'task': {'name': 'ClassificationTask', 'params': {
        'backbone_name': 'resnet18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
        'pooling_name': 'Pooling', 'head_name': 'ClassificationHead', 'head_params': {'num_classes': 10},
        'inputs': [{'shape': [3, 32, 32], 'dtype': 'float32'}]}},
    'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]},
    'optimization': [{'optimizer': {'name': 'Adam', 'params': {'lr': 1e-4}}}],
}


def test_task_assembly_from_config_and_state_dict_contract():
    cfg = tb.load_config(CFG)
    task = tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params)
    assert task.backbone.out_channels == 512 and task.pooling.out_channels == 512 and task.head.out_channels == 10
    assert tuple(task.backbone.out_encoder_channels) == (64, 64, 128, 256, 512)
    keys = set(task.state_dict())
    # timm / torchvision naming: what load_checkpoint / pretrained weights rely on
    for k in ('backbone.conv1.weight', 'backbone.bn1.running_var', 'backbone.layer1.0.conv1.weight',
              'backbone.layer2.0.downsample.0.weight', 'backbone.layer4.1.bn2.num_batches_tracked',
              'head.fc.weight', 'head.fc.bias', 'input_tensors_0'):
        assert k in keys, k
    import torchvision
    tv = torchvision.models.resnet18()
    sd = {k[len('backbone.'):]: v for k, v in task.state_dict().items() if k.startswith('backbone.')}
    tv_sd = {k: v for k, v in tv.state_dict().items() if not k.startswith('fc.')}
    assert set(sd) == set(tv_sd)
    assert all(tuple(sd[k].shape) == tuple(tv_sd[k].shape) for k in sd)
    assert isinstance(task.as_module(), torch.nn.Sequential)
    with pytest.raises(KeyError, match='is not in the backbones registry'):
        tb.TASKS.get('ClassificationTask')(cfg, backbone_name='resnet_nope')


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing something else."""
    cfg = tb.load_config(CFG)
    task = tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params)
    with pytest.raises(RuntimeError, match='CUDA'):
        task.forward_with_gt({'image': torch.randn(2, 3, 32, 32), 'target': torch.zeros(2, dtype=torch.long)})


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import sys, torchok_b200, torchok_b200.engine, torchok_b200.kernels; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; sys.exit(1 if bad else 0)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, '-c', code], cwd=root).returncode == 0
    for dirpath, _, files in os.walk(os.path.join(root, 'torchok_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(dirpath, f)


# ---------------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB_PATH):
        import __graft_entry__ as g
        g.build()
    dll = ctypes.CDLL(LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing
    dll.tok_version.restype = ctypes.c_int
    assert dll.tok_version() >= 1
    # no GPU here: the device probe must report an error code, not crash
    if not torch.cuda.is_available():
        dll.tok_device_ok.restype = ctypes.c_int
        assert dll.tok_device_ok() < 0
        dll.tok_last_error.restype = ctypes.c_char_p
        assert dll.tok_last_error()


# ------------------------------------------------------------------------------------------------ N > 1 gradient path
def _bucket_worker(rank, world, port, ret):
    """The data-parallel gradient path of engine.ParamArena / BucketAllReduce on CPU tensors: bucket plan, readiness
    counting in backward order, one all-reduce (sum) per bucket, leftovers reduced by finish()."""
    import torch.distributed as dist
    from torchok_b200 import engine
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    numels = [9408, 64, 64, 36864, 64, 64, 147456, 513, 1000, 7]
    arena = engine.ParamArena.__new__(engine.ParamArena)       # the bookkeeping half only: no CUDA arenas on this box
    arena.buckets = engine.plan_buckets(numels, bucket_mb=0.05)
    arena.begin_step()
    total = arena.buckets[-1][1]
    gen = torch.Generator().manual_seed(100 + rank)
    arena.grad = torch.randn(total, generator=gen)
    mine = arena.grad.clone()
    launched = []

    class GlooReducer:
        def launch(self, b):
            begin, end, _ = arena.buckets[b]
            launched.append(b)
            dist.all_reduce(arena.grad[begin:end], op=dist.ReduceOp.SUM)

        def join(self):
            pass
    arena.reducer = GlooReducer()
    owner = {i: b for b, (_, _, m) in enumerate(arena.buckets) for i in m}
    for i in reversed(range(len(numels))):      # backward produces gradients last layer first
        if i == 3:
            continue                            # a parameter whose backward never announces itself
        arena.ready(owner[i])
    arena.finish()
    other = torch.randn(total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ret[rank] = dict(ok=bool(torch.allclose(arena.grad, mine + other)), launched=launched, buckets=arena.buckets)
    dist.destroy_process_group()


def test_gradient_buckets_gloo_world2():
    import torch.multiprocessing as mp
    from torchok_b200 import engine
    numels = [9408, 64, 64, 36864, 64, 64, 147456, 513, 1000, 7]
    buckets = engine.plan_buckets(numels, bucket_mb=0.05)
    # the plan tiles the arena: contiguous, aligned, every parameter in exactly one bucket, in order
    assert buckets[0][0] == 0 and all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    assert [i for _, _, m in buckets for i in m] == list(range(len(numels)))
    assert buckets[-1][1] == sum((n + 63) // 64 * 64 for n in numels)
    assert all(e - b >= int(0.05 * (1 << 20) / 4) for b, e, _ in buckets[:-1])   # 0.05 MiB of fp32 before a bucket closes
    assert len(buckets) > 2
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        assert ret[r]['ok']                                        # every element is the sum over both ranks
        assert sorted(ret[r]['launched']) == list(range(len(buckets)))   # each bucket reduced exactly once
    assert ret[0]['launched'] == ret[1]['launched']                # same collective order on both ranks
    # buckets complete from the end of the arena backwards; the one with the silent parameter is left to finish()
    silent = [b for b, (_, _, m) in enumerate(buckets) if 3 in m][0]
    assert ret[0]['launched'][-1] == silent


# ------------------------------------------------------------------------------------------------ paramwise_cfg
def test_paramwise_cfg_rules():
    """The rules of Constructor.add_params (torchok/constructor/constructor.py:162-251): custom_keys beat everything
    (longest key first), bias_lr_mult skips norm biases, decay precedence norm > depth-wise conv > bias."""
    import torch.nn as nn
    from torchok_b200.constructor.paramwise import paramwise_multipliers
    m = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 8, 3, groups=8), nn.Linear(8, 4))
    m[3].bias.requires_grad_(False)
    cfg = dict(bias_lr_mult=2., bias_decay_mult=0., norm_decay_mult=0.5, dwconv_decay_mult=0.25,
               custom_keys={'0': dict(lr_mult=7.), '3.weight': dict(lr_mult=0.1), '3.w': dict(decay_mult=9.)})
    got = {n: paramwise_multipliers(m, cfg)[p] for n, p in m.named_parameters()}
    assert got == {'0.weight': (7., 1.), '0.bias': (7., 1.),      # custom key: other rules skipped (no bias_lr_mult)
                   '1.weight': (1., 0.5), '1.bias': (1., 0.5),    # norm: no bias_lr_mult, norm decay
                   '2.weight': (1., 0.25), '2.bias': (2., 0.25),  # depth-wise conv decay wins over bias decay
                   '3.weight': (0.1, 1.),                         # the longer key '3.weight' wins over '3.w'
                   '3.bias': (1., 1.)}                            # frozen: defaults
    assert all(v == (1., 1.) for v in paramwise_multipliers(m, None).values())
