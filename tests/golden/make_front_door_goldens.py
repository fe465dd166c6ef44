"""Golden vectors for the front-door host logic, produced by the REFERENCE's own files.

Like make_reference_goldens.py: the reference package cannot be imported here (pytorch_lightning, torchmetrics,
omegaconf, timm ... are not installed), but these files only need torch once their framework imports are stubs:

    torchok/constructor/config_structure.py   (dataclasses only: executed as is)
    torchok/constructor/load.py               (stubs: pytorch_lightning [annotation], lightning_fabric …cloud_io._load = torch.load)
    torchok/metrics/metrics_manager.py        (stubs: torchmetrics.Metric [annotation]; METRICS = a name → class table)
    torchok/callbacks/freeze_unfreeze.py      (stub: pytorch_lightning.callbacks.BaseFinetuning with Lightning 2.0's
                                               flatten_modules / filter_params / filter_on_optimizer restated below —
                                               the freeze policy itself is the reference's code)

Output: tests/golden/front_door_goldens.pt, replayed by tests/test_front_door_goldens.py through
torchok_b200.constructor.load / metrics.MetricsManager / callbacks.FreezeUnfreeze.   Needs /root/reference:

    python tests/golden/make_front_door_goldens.py
"""
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_reference_goldens import install_stub_tree, load  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'front_door_goldens.pt')


# ------------------------------------------------------------------------------------------------ shared fixtures
def tree_from_spec(spec):
    """Module tree from [(name, 'Linear'|'Conv2d'|'BatchNorm2d'|'Dropout'|'ReLU', args) | (name, [children])];
    a container named 'backbone' gets `get_stages(n)` = its first n+1 children (stage 0 = stem)."""
    kinds = {'Linear': nn.Linear, 'Conv2d': nn.Conv2d, 'BatchNorm2d': nn.BatchNorm2d, 'Dropout': nn.Dropout,
             'ReLU': nn.ReLU, 'BatchNorm1d': nn.BatchNorm1d}

    class Staged(nn.Sequential):
        def get_stages(self, stage):
            return nn.ModuleList(list(self.children())[:stage + 1])
    seq = Staged() if any(item[0] == 'stem' for item in spec) else nn.Sequential()
    for item in spec:
        if isinstance(item[1], list):
            seq.add_module(item[0], tree_from_spec(item[1]))
        else:
            seq.add_module(item[0], kinds[item[1]](*item[2]))
    return seq


TASK_SPEC = [('backbone', [('stem', [('conv', 'Conv2d', (3, 4, 3)), ('bn', 'BatchNorm2d', (4,))]),
                           ('layer1', [('conv', 'Conv2d', (4, 4, 3)), ('bn', 'BatchNorm2d', (4,)), ('act', 'ReLU', ())]),
                           ('layer2', [('conv', 'Conv2d', (4, 8, 3)), ('bn', 'BatchNorm2d', (8,)), ('drop', 'Dropout', (0.1,))])]),
             ('neck', [('fc', 'Linear', (8, 8)), ('bn', 'BatchNorm1d', (8,))]),
             ('head', [('fc', 'Linear', (8, 3))])]

FREEZE_CASES = [
    dict(rules=[dict(module_name='backbone', epoch=2), dict(module_name='backbone', stages=1),
                dict(module_name='backbone', module_class='_BatchNorm', bn_requires_grad=False,
                     bn_track_running_stats=False)], top_down=True),
    dict(rules=[dict(module_name='backbone', epoch=1), dict(module_name='backbone.layer2', epoch=3),
                dict(module_name='neck', epoch=2, bn_requires_grad=True, bn_track_running_stats=True)], top_down=True),
    dict(rules=[dict(module_name='backbone.layer2', epoch=1), dict(module_name='backbone')], top_down=False),
    dict(rules=[dict(module_name='', module_class='Dropout'), dict(module_name='', module_class='Linear', epoch=1)],
         top_down=True),
]


def flags(task):
    return ({n: p.requires_grad for n, p in task.named_parameters()},
            {n: m.track_running_stats for n, m in task.named_modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)})


class MockSum:
    """Counts updates (the reference test's MockSumMetric, test_metric_manager.py:13-24)."""

    def __init__(self, start=0, **kw):
        self.start = self.sum = start

    def update(self, predict, target):
        self.sum += 1

    def compute(self):
        return torch.tensor(self.sum)

    def reset(self):
        self.sum = self.start


class MockDict(MockSum):
    def compute(self):
        return {'a': torch.tensor(self.sum), 'b': torch.tensor(2 * self.sum), 'text': 'skipped'}


MANAGER_CASES = [
    [dict(name='MockSum', mapping=dict(predict='emb', target='y'), tag=None, phases=['TRAIN', 'VALID'])],
    [dict(name='MockSum', mapping=dict(predict='emb', target='y'), tag='first', phases=['TRAIN']),
     dict(name='MockSum', mapping=dict(predict='emb', target='y'), tag=None, params=dict(start=10))],
    [dict(name='MockDict', mapping=dict(predict='emb', target='y'), tag='d', val_dataloader_idxs=[0, 2],
          test_dataloader_idxs=[1])],
]
UPDATES = [('TRAIN', 0), ('TRAIN', 0), ('VALID', 0), ('VALID', 2), ('VALID', 2), ('VALID', 1), ('TEST', 1), ('TEST', 0),
           ('PREDICT', 0)]


REGISTRY_NAMES = [('resnet18', 'resnet'), ('resnet50', 'resnet'), ('resnet101', 'resnet'), ('resnet9', 'resnet'),
                  ('hrnet_w18', 'hrnet'), ('hrnet_w18_small_v2', 'hrnet'), ('swinv2_tiny_window8_256', 'swin'),
                  ('Resnet_B', 'resnet')]
REGISTRY_QUERIES = [dict(), dict(filter='resnet*'), dict(filter='*net*', exclude_filters='hr*'),
                    dict(filter=['hrnet*', 'swin*']), dict(module='resnet'), dict(filter='*18*', module='hrnet'),
                    dict(filter='nothing*'), dict(exclude_filters=['*small*', 'resnet1*']), dict(module='missing')]


def registry_scenario(registry_cls):
    """Everything a caller can observe from a Registry: listing under filters, lookups, error types and texts."""
    reg = registry_cls('backbones')
    for name, module in REGISTRY_NAMES:
        mod = sys.modules.setdefault(f'fake_pkg.{module}', types.ModuleType(f'fake_pkg.{module}'))
        fn = types.FunctionType((lambda: None).__code__, {}, name)
        fn.__module__ = mod.__name__
        assert reg.register_class(fn) is fn
    res = {'lists': [reg.list_models(**q) for q in REGISTRY_QUERIES],
           'contains': ['resnet18' in reg, 'nope' in reg], 'repr': repr(reg),
           'object_to_module': dict(reg.object_to_module),
           'exported': sorted(sys.modules['fake_pkg.hrnet'].__all__)}
    errors = []
    for action in (lambda: reg.get('nope'), lambda: reg['nope2'], lambda: reg.register_class(42),
                   lambda: reg.register_class(reg.get('resnet18'))):
        try:
            action()
            errors.append(None)
        except Exception as e:  # noqa: BLE001
            errors.append((type(e).__name__, str(e)))
    res['errors'] = errors
    for module in {m for _, m in REGISTRY_NAMES}:
        del sys.modules[f'fake_pkg.{module}']
    return res


def main():
    install_stub_tree()
    out = {}
    g = torch.Generator().manual_seed(7)

    # ---- config_structure + metrics_manager -------------------------------------------------------------------
    cs = load('torchok.constructor.config_structure')                  # the real dataclasses / Phase enum
    tm = types.ModuleType('torchmetrics')
    tm.Metric = object
    sys.modules['torchmetrics'] = tm

    class Table:
        def get(self, name):
            return {'MockSum': MockSum, 'MockDict': MockDict}[name]
    sys.modules['torchok.constructor'].METRICS = Table()
    import dataclasses
    out['config_structure'] = {name: [f.name for f in dataclasses.fields(obj)] for name, obj in vars(cs).items()
                               if dataclasses.is_dataclass(obj)}
    out['phases'] = [p.name for p in cs.Phase]
    mmod = load('torchok.metrics.metrics_manager')
    cases = []
    for spec in MANAGER_CASES:
        params = [cs.MetricParams(**{**p, 'phases': [cs.Phase[x] for x in p['phases']]} if 'phases' in p else p)
                  for p in spec]
        mgr = mmod.MetricsManager(params)
        for phase, idx in UPDATES:
            mgr.update(cs.Phase[phase], idx, emb=torch.zeros(1), y=torch.zeros(1))
        logs = {}
        for phase in ('TRAIN', 'VALID', 'TEST', 'PREDICT'):
            logs[phase] = {k: (int(v) if torch.is_tensor(v) else v)
                           for k, v in mgr.on_epoch_end(cs.Phase[phase]).items()}
        again = {k: int(v) for k, v in mgr.on_epoch_end(cs.Phase.VALID).items()}      # after the reset
        cases.append(dict(spec=spec, updates=UPDATES, logs=logs, valid_after_reset=again))
    out['MetricsManager'] = cases

    # ---- load.py ----------------------------------------------------------------------------------------------
    pl = types.ModuleType('pytorch_lightning')
    pl.LightningModule = nn.Module
    sys.modules['pytorch_lightning'] = pl
    for name in ('lightning_fabric', 'lightning_fabric.utilities'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    cio = types.ModuleType('lightning_fabric.utilities.cloud_io')
    cio._load = lambda path, map_location=None: torch.load(path, map_location=map_location, weights_only=False)
    sys.modules['lightning_fabric.utilities.cloud_io'] = cio
    ld = load('torchok.constructor.load')

    def rand_state(module):
        return {k: (torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone())
                for k, v in module.state_dict().items()}
    cases = []
    for overrides, excludes in [({}, []), ({'backbone': 'backbone'}, []), ({'backbone': 'backbone'}, ['head']),
                                ({'backbone': 'backbone', 'backbone.layer2': 'backbone.layer2'}, ['backbone.stem', 'neck.bn']),
                                ({'backbone.layer1': 'backbone.layer1', 'backbone': 'backbone'}, ['backbone.layer2.conv.bias'])]:
        task = tree_from_spec(TASK_SPEC)
        initial = {k: v.clone() for k, v in task.state_dict().items()}
        base = rand_state(task)
        files = {}
        with tempfile.TemporaryDirectory() as tmp:
            torch.save({'state_dict': base}, os.path.join(tmp, 'base.ckpt'))       # Lightning layout
            name2path = {}
            for name, sub in overrides.items():
                submodule = task.get_submodule(sub)
                files[name] = rand_state(submodule)                                # keys relative to the module
                torch.save(files[name], os.path.join(tmp, f'{name}.pth'))
                name2path[name] = os.path.join(tmp, f'{name}.pth')
            ld.load_checkpoint(task, os.path.join(tmp, 'base.ckpt'), name2path or None, excludes or None)
        cases.append(dict(spec=TASK_SPEC, initial=initial, base=base, overrides=files, exclude_keys=excludes,
                          loaded={k: v.clone() for k, v in task.state_dict().items()}))
    out['load_checkpoint'] = cases

    # ---- freeze_unfreeze.py -------------------------------------------------------------------------------------
    class BaseFinetuning:          # pytorch_lightning 2.0.2 callbacks/finetuning.py, the three helpers the file calls
        def __init__(self):
            pass

        @staticmethod
        def flatten_modules(modules):
            if isinstance(modules, nn.ModuleDict):
                modules = modules.values()
            if isinstance(modules, nn.Module):
                mods = modules.modules()
            else:
                mods = [x for m in modules for x in BaseFinetuning.flatten_modules(m)]
            return [m for m in mods if not list(m.children()) or m._parameters]

        @staticmethod
        def filter_params(modules, train_bn=True, requires_grad=True):
            for mod in BaseFinetuning.flatten_modules(modules):
                if isinstance(mod, nn.modules.batchnorm._BatchNorm) and not train_bn:
                    continue
                for p in mod.parameters(recurse=False):
                    if p.requires_grad == requires_grad:
                        yield p

        @staticmethod
        def filter_on_optimizer(optimizer, params):
            have = {id(p) for gr in optimizer.param_groups for p in gr['params']}
            return [p for p in params if id(p) not in have]
    cb = types.ModuleType('pytorch_lightning.callbacks')
    cb.BaseFinetuning = BaseFinetuning
    pl.__path__ = []
    sys.modules['pytorch_lightning.callbacks'] = cb
    fu = load('torchok.callbacks.freeze_unfreeze')
    cases = []
    for case in FREEZE_CASES:
        task = tree_from_spec(TASK_SPEC)
        opt = torch.optim.SGD(task.parameters(), lr=0.1)          # every parameter already has a group (add_params)
        callback = fu.FreezeUnfreeze(case['rules'], top_down_freeze_order=case['top_down'])
        callback.freeze_before_training(task)
        history = [flags(task)]
        for epoch in range(4):
            callback.finetune_function(task, epoch, opt, 0)
            history.append(flags(task))
        cases.append(dict(spec=TASK_SPEC, rules=case['rules'], top_down=case['top_down'], history=history,
                          n_groups=len(opt.param_groups)))
    out['FreezeUnfreeze'] = cases

    # ---- registry.py (the plug-in boundary itself) ---------------------------------------------------------------
    import re
    for name in ('timm', 'timm.models'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    treg = types.ModuleType('timm.models.registry')      # timm 0.6.13 models/registry.py `_natural_key`
    treg._natural_key = lambda s: [int(t) if t.isdigit() else t for t in re.split(r'(\d+)', s.lower())]
    sys.modules['timm.models.registry'] = treg
    rg = load('torchok.constructor.registry')
    out['Registry'] = registry_scenario(rg.Registry)

    torch.save(out, OUT)
    print(f'wrote {OUT}: ' + ', '.join(f'{k} x{len(v)}' for k, v in out.items()))


if __name__ == '__main__':
    main()
