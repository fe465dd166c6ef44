"""CPU tests of the YAML / CLI front door: load_checkpoint, MetricsManager + meters, lr schedulers,
FreezeUnfreeze, ModelCheckpoint / resume and the Runner's epoch logic.

Known answers come from the reference's own tests: tests/base_tests/constructor/test_load_checkpoint.py:43-140,
tests/base_tests/metrics/metric_manager/test_metric_manager.py:100-185 and test_metric_manager_ddp.py:15-24.
The Runner tests drive the real Runner with a CPU stand-in for the step engine (engine.StreamLoop needs the GPU); the
product path never takes that hook — tests/test_front_door_gpu.py runs the module entry point on the GPU.
"""
import os
import textwrap

import pytest
import torch
import torch.nn as nn

import torchok_b200 as tb
from torchok_b200.callbacks import FreezeUnfreeze, ModelCheckpoint, get_modules
from torchok_b200.constructor.load import generate_required_state_dict, load_checkpoint
from torchok_b200.metrics.metrics_manager import Metric, MetricsManager, Phase
from torchok_b200.optim import LrDriver
from torchok_b200.runner import Runner, _limit

# ------------------------------------------------------------------------------------------------ load_checkpoint
MODEL_KEYS = ['layer1.module.conv1.weight', 'layer1.linear.weight', 'linear.weight']
INITIAL = {k: 0 for k in MODEL_KEYS}
BASE = {'layer1.module.conv1.weight': 1, 'layer1.linear.weight': 2, 'linear.weight': 3}


@pytest.mark.parametrize('overrides,exclude,answer', [
    ({}, [], BASE),
    ({'layer1': {'layer1.module.conv1.weight': 11, 'layer1.linear.weight': 22}}, [],
     {'layer1.module.conv1.weight': 11, 'layer1.linear.weight': 22, 'linear.weight': 3}),
    ({'layer1': {'layer1.module.conv1.weight': 11, 'layer1.linear.weight': 22}}, ['layer1.module'],
     {'layer1.linear.weight': 22, 'linear.weight': 3, 'layer1.module.conv1.weight': 0}),
    ({'layer1': {'module.conv1.weight': 11, 'linear.weight': 22}, 'layer1.linear': {'weight': 222}}, ['layer1.module'],
     {'layer1.linear.weight': 222, 'linear.weight': 3, 'layer1.module.conv1.weight': 0}),
])
def test_generate_required_state_dict_known_answers(overrides, exclude, answer):
    assert generate_required_state_dict(BASE, overrides, exclude, MODEL_KEYS, INITIAL) == answer


def test_generate_required_state_dict_deepest_override_wins():
    """The scenario of the docstring example (torchok/constructor/load.py:110-139) with the deeper override addressed
    as the code requires (a module prefix + relative key; taken literally the docstring's `'backbone.linear.1'` module
    name would be prefixed onto its own key by get_state_dict_with_prefix, load.py:62-70)."""
    keys = ['backbone.linear.1', 'backbone.linear.2', 'head.linear.1', 'head.linear.2']
    got = generate_required_state_dict({k: 1 for k in keys},
                                       {'backbone.linear': {'1': 10},
                                        'backbone': {'backbone.linear.1': 5, 'backbone.linear.2': 3}},
                                       ['head.linear.2'], keys, {k: 0 for k in keys})
    assert got == {'backbone.linear.1': 10, 'backbone.linear.2': 3, 'head.linear.1': 1, 'head.linear.2': 0}


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(2, 2, 1)
        self.linear = nn.Linear(2, 2)


class _Model(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer = nn.Sequential()
        self.layer.add_module('block1', _Block())
        self.layer.add_module('block2', _Block())
        self.linear = nn.Linear(2, 2)


def test_load_checkpoint_files(tmp_path):
    torch.manual_seed(0)
    base, layer, model = _Model(), _Model().layer, _Model()
    initial = {k: v.clone() for k, v in model.state_dict().items()}
    torch.save(base.state_dict(), tmp_path / 'base.pth')
    torch.save({'state_dict': layer.state_dict()}, tmp_path / 'layer.pth')        # Lightning layout
    load_checkpoint(model, str(tmp_path / 'base.pth'), {'layer': str(tmp_path / 'layer.pth')}, ['layer.block2'])
    got = model.state_dict()
    for k in got:
        if k.startswith('layer.block2'):
            want = initial[k]
        elif k.startswith('layer.'):
            want = layer.state_dict()[k[len('layer.'):]]
        else:
            want = base.state_dict()[k]
        assert torch.equal(got[k], want), k
    with pytest.raises(Exception):      # the base checkpoint does not cover the model (strict)
        load_checkpoint(_Model(), str(tmp_path / 'layer.pth'), {}, [])
    with pytest.raises(Exception):      # override module that is not in the model
        load_checkpoint(_Model(), str(tmp_path / 'base.pth'), {'loyer': str(tmp_path / 'layer.pth')}, ['layer.block2'])
    with pytest.raises(ValueError, match='exclude key'):
        load_checkpoint(_Model(), str(tmp_path / 'base.pth'), {'layer': str(tmp_path / 'layer.pth')}, ['loyer.block2'])
    load_checkpoint(_Model())           # nothing to load: no-op


# ------------------------------------------------------------------------------------------------ metrics manager
def _register(cls):
    if cls.__name__ not in tb.METRICS:
        tb.METRICS.register_class(cls)
    return cls


@_register
class MockSumMetric(Metric):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.add_state('sum', default=torch.tensor(0), dist_reduce_fx=None)

    def update(self, predict, target):
        self.sum += 1

    def compute(self):
        return self.sum


@_register
class MockDictMetric(Metric):
    def update(self, predict, target):
        return

    def compute(self):
        return {'target_shape': torch.tensor(10), 'embedding_size': torch.tensor(512)}


@_register
class MockConstantMetric(Metric):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.add_state('constant', default=torch.tensor(0), dist_reduce_fx=None)

    def update(self, predict, target):
        return

    def compute(self):
        return self.constant


@_register
class MockRaiseMetric(Metric):
    def update(self, predict, target):
        return

    def compute(self):
        return torch.tensor([1, 2])


def _run_manager(names, tags):
    mapping = dict(predict='embedding', target='target')
    params = [dict(name=n, mapping=mapping, tag=t, phases=['TRAIN']) for n, t in zip(names, tags)]
    mm = MetricsManager(params)
    for _ in range(5):
        mm.update(Phase.TRAIN, embedding=torch.rand(4, 512), target=torch.rand(4, 10))
    return mm.on_epoch_end(Phase.TRAIN)


def test_metrics_manager_known_answers():
    assert _run_manager(['MockSumMetric'], [None]) == {'train/MockSumMetric': 5}
    assert _run_manager(['MockSumMetric', 'MockConstantMetric'], ['moc_sum', None]) == \
        {'train/moc_sum': 5, 'train/MockConstantMetric': 0}
    assert _run_manager(['MockDictMetric'], [None]) == \
        {'train/MockDictMetric_target_shape': 10, 'train/MockDictMetric_embedding_size': 512}
    with pytest.raises(ValueError, match='no numeric value'):
        _run_manager(['MockRaiseMetric'], [None])
    with pytest.raises(ValueError, match='identical names'):
        _run_manager(['MockSumMetric', 'MockSumMetric'], [None, None])
    mm = MetricsManager([dict(name='MockSumMetric', mapping={'predict': 'nope', 'target': 'target'})])
    with pytest.raises(ValueError, match='Cannot find nope'):
        mm.update(Phase.VALID, target=torch.zeros(1))
    # reset after on_epoch_end; per-dataloader names in VALID
    mm = MetricsManager([dict(name='MockSumMetric', mapping=dict(predict='p', target='t'), val_dataloader_idxs=[0, 1])])
    mm.update(Phase.VALID, 1, p=0, t=0)
    assert mm.on_epoch_end(Phase.VALID) == {'valid/MockSumMetric_dataloader_0': 0, 'valid/MockSumMetric_dataloader_1': 1}
    assert mm.on_epoch_end(Phase.VALID) == {'valid/MockSumMetric_dataloader_0': 0, 'valid/MockSumMetric_dataloader_1': 0}


LABELS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9,
          9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
PREDICTS = [0, 0, 1, 3, 3, 4, 5, 6, 7, 0, 0, 1, 1, 2, 3, 3, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7,
            8, 8, 9, 9, 7, 7, 8, 8, 8, 8, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 9]


def test_accuracy_known_answer_and_sklearn_cross_check():
    """test_metric_manager_ddp.py:15-24: Accuracy(task='multiclass', num_classes=10) over these labels = 0.18."""
    mm = MetricsManager([dict(name='Accuracy', mapping=dict(preds='predict', target='target'), phases=['TRAIN'],
                              params=dict(task='multiclass', num_classes=10))])
    for _ in range(5):                                      # five epochs of the same data: still 0.18
        for i in range(0, 50, 4):
            mm.update(Phase.TRAIN, predict=torch.tensor(PREDICTS[i:i + 4]), target=torch.tensor(LABELS[i:i + 4]))
    out = mm.on_epoch_end(Phase.TRAIN)
    assert float(out['train/Accuracy']) == pytest.approx(0.18, abs=1e-7)

    from sklearn.metrics import accuracy_score, f1_score, jaccard_score, precision_score, recall_score
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(500, 7, generator=g)
    target = torch.randint(0, 7, (500,), generator=g)
    logits[torch.arange(500), target] += 1.0
    pred = logits.argmax(1).numpy()
    cases = [('Accuracy', {}, accuracy_score(target, pred)),
             ('F1Score', {}, f1_score(target, pred, average='micro')),
             ('F1Score', {'average': 'macro'}, f1_score(target, pred, average='macro')),
             ('F1Score', {'average': 'weighted'}, f1_score(target, pred, average='weighted')),
             ('Precision', {'average': 'macro'}, precision_score(target, pred, average='macro')),
             ('Recall', {'average': 'macro'}, recall_score(target, pred, average='macro')),
             ('Accuracy', {'average': 'macro'}, recall_score(target, pred, average='macro')),
             ('JaccardIndex', {}, jaccard_score(target, pred, average='macro')),
             ('JaccardIndex', {'average': 'micro'}, jaccard_score(target, pred, average='micro'))]
    for name, kw, want in cases:
        m = tb.METRICS.get(name)(task='multiclass', num_classes=7, **kw)
        for i in range(0, 500, 64):                          # float logits, arg-maxed inside
            m.update(logits[i:i + 64], target[i:i + 64])
        assert float(m.compute()) == pytest.approx(float(want), abs=1e-6), (name, kw)
    # segmentation layout: (N, C, H, W) logits against (N, H, W) labels, ignore_index
    seg_logits = torch.randn(2, 4, 8, 8, generator=g)
    seg_target = torch.randint(0, 4, (2, 8, 8), generator=g)
    seg_target[0, 0] = 255
    m = tb.METRICS.get('JaccardIndex')(task='multiclass', num_classes=4, ignore_index=255)
    m.update(seg_logits, seg_target)
    keep = seg_target.reshape(-1) != 255
    want = jaccard_score(seg_target.reshape(-1)[keep], seg_logits.argmax(1).reshape(-1)[keep], average='macro')
    assert float(m.compute()) == pytest.approx(float(want), abs=1e-6)
    with pytest.raises(NotImplementedError):
        tb.METRICS.get('Accuracy')(task='binary')


# ------------------------------------------------------------------------------------------------ schedulers
class _FakeOpt:
    def __init__(self, lr):
        self.lr = lr


@pytest.mark.parametrize('name,params,interval', [('ExponentialLR', dict(gamma=0.97), 'epoch'),
                                                  ('StepLR', dict(step_size=2, gamma=0.1), 'epoch'),
                                                  ('CosineAnnealingLR', dict(T_max=7), 'step'),
                                                  ('OneCycleLR', dict(max_lr=0.5, total_steps=12), 'step')])
def test_lr_driver_tracks_torch_scheduler(name, params, interval):
    ref_opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=0.1)
    ref = tb.SCHEDULERS.get(name)(ref_opt, **params)
    opt = _FakeOpt(0.1)
    drv = LrDriver(opt, name, params, {'interval': interval})
    assert opt.lr == pytest.approx(ref_opt.param_groups[0]['lr'])
    for _ in range(10):
        ref_opt.step()
        ref.step()
        (drv.epoch_end if interval == 'epoch' else drv.step_end)()
        (drv.step_end if interval == 'epoch' else drv.epoch_end)()      # the other interval must not tick
        assert opt.lr == pytest.approx(ref_opt.param_groups[0]['lr'], rel=1e-12)
    state = drv.state_dict()
    opt2 = _FakeOpt(0.1)
    drv2 = LrDriver(opt2, name, params, {'interval': interval})
    drv2.load_state_dict(state)
    assert opt2.lr == pytest.approx(opt.lr)


def test_reduce_on_plateau_uses_monitor():
    opt = _FakeOpt(1.0)
    drv = LrDriver(opt, 'ReduceLROnPlateau', dict(patience=0, factor=0.5), {'monitor': 'valid/loss'})
    drv.epoch_end({'valid/loss': 1.0})
    drv.epoch_end({'valid/loss': 2.0})
    assert opt.lr == 0.5
    with pytest.raises(KeyError):
        drv.epoch_end({'other': 1.0})


# ------------------------------------------------------------------------------------------------ FreezeUnfreeze
CIFAR_RULES = [dict(module_name='backbone', epoch=2), dict(module_name='backbone', stages=1),
               dict(module_name='backbone', module_class='_BatchNorm', bn_requires_grad=False,
                    bn_track_running_stats=False)]


def _cifar_task():
    cfg = tb.load_config({
        'task': {'name': 'ClassificationTask',
                 'params': {'backbone_name': 'resnet18', 'backbone_params': {'in_channels': 3},
                            'pooling_name': 'Pooling', 'head_name': 'ClassificationHead',
                            'head_params': {'num_classes': 10}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]}})
    return tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params)


def test_freeze_unfreeze_policy_of_the_cifar_example():
    """examples/configs/classification_cifar10.yaml:103-121: backbone frozen for 2 epochs, stem + layer1 forever,
    backbone BatchNorms never trained and not tracking statistics."""
    task = _cifar_task()
    cb = FreezeUnfreeze(CIFAR_RULES)
    bns = [m for m in task.backbone.modules() if isinstance(m, nn.BatchNorm2d)]
    cb.apply(task, 0)
    assert not any(p.requires_grad for p in task.backbone.parameters())
    assert all(p.requires_grad for p in task.head.parameters())
    assert not any(m.track_running_stats for m in bns)
    cb.apply(task, 1)
    assert not any(p.requires_grad for p in task.backbone.parameters())
    cb.apply(task, 2)
    frozen_forever = {id(p) for m in (task.backbone.conv1, task.backbone.bn1, task.backbone.layer1)
                      for p in m.parameters()}
    bn_params = {id(p) for m in bns for p in m.parameters()}
    for name, p in task.backbone.named_parameters():
        assert p.requires_grad == (id(p) not in frozen_forever and id(p) not in bn_params), name
    assert not any(m.track_running_stats for m in bns)
    # rule addressing errors (freeze_unfreeze.py:26-45)
    with pytest.raises(ValueError, match='is not found'):
        get_modules(dict(module_name='backbone.nope'), task)
    with pytest.raises(ValueError, match='get_stages'):
        get_modules(dict(module_name='head', stages=1), task)
    with pytest.raises(ValueError, match='does not have submodules'):
        get_modules(dict(module_name='head', module_class='Dropout2d'), task)
    assert len(get_modules(dict(module_name='', module_class='BatchNorm2d'), task)) == len(bns)
    # bottom-up order: the later (shallower) rule overwrites the deeper one
    t2 = _cifar_task()
    FreezeUnfreeze([dict(module_name='backbone.layer4', epoch=0), dict(module_name='backbone')],
                   top_down_freeze_order=False).apply(t2, 0)
    assert not any(p.requires_grad for p in t2.backbone.layer4.parameters())
    t3 = _cifar_task()
    FreezeUnfreeze([dict(module_name='backbone.layer4', epoch=0), dict(module_name='backbone')]).apply(t3, 0)
    assert not any(p.requires_grad for p in t3.backbone.parameters())   # thaw first, then the forever-rule freezes all


# ------------------------------------------------------------------------------------------------ Runner
class _TinyTask(nn.Module):
    """CPU stand-in task with the BaseTask surface the Runner touches."""

    def __init__(self, hparams, in_features=8, num_classes=3, **kwargs):
        super().__init__()
        self._hparams = hparams
        self.backbone = nn.Sequential(nn.Linear(in_features, 16), nn.ReLU())
        self.head = nn.Linear(16, num_classes)
        self.losses = tb.tasks.base.configure_losses(hparams)

    def forward_with_gt(self, batch):
        x = batch['image'].float().flatten(1)
        return {'prediction': self.head(self.backbone(x)), 'target': batch['target']}

    def training_step(self, batch, batch_idx=0):
        out = self.forward_with_gt(batch)
        total, tagged = self.losses(**out)
        self.last_output = {k: v.detach() for k, v in out.items()}
        return {'loss': total, **tagged}

    def validation_step(self, batch, batch_idx=0, dataloader_idx=0):
        out = self.forward_with_gt(batch)
        total, tagged = self.losses(**out)
        return {'loss': total, **tagged}, out

    def predict_step(self, batch, batch_idx=0, dataloader_idx=0):
        return self.forward_with_gt(batch)


class _CpuCrossEntropy(nn.CrossEntropyLoss):
    """Test-only CPU loss for the stand-in task (the registered CrossEntropyLoss is the CUDA kernel)."""


if '_TinyTask' not in tb.TASKS:
    tb.TASKS.register_class(_TinyTask)
    tb.LOSSES.register_class(_CpuCrossEntropy)


class _CpuOptimizer:
    def __init__(self, task, lr):
        self.lr = lr
        self.mults = None
        self.task = task
        self.steps = 0

    def set_param_multipliers(self, mults):
        self.mults = mults

    def state_dict(self, module):
        return {'step': self.steps, 'lr': self.lr, 'state': {}}

    def load_state_dict(self, state, module):
        self.steps, self.lr = state['step'], state['lr']


class _CpuLoop:
    """Stand-in for engine.StreamLoop: plain SGD on the CPU, honouring the (lr_mult, decay_mult) table."""

    class _Arena:
        def refresh_shadow(self):
            pass

    def __init__(self, task, optimizer_cfg):
        self.task, self.arena = task, self._Arena()
        self.optimizer = _CpuOptimizer(task, optimizer_cfg['params']['lr'])
        self.graph_resets = 0

    def reset_graph(self):
        self.graph_resets += 1

    def train_step(self, batch):
        self.task.train()
        for p in self.task.parameters():
            p.grad = None
        out = self.task.training_step(batch)
        out['loss'].backward()
        import torch.distributed as dist
        with torch.no_grad():
            for p in self.task.parameters():
                if p.grad is not None and dist.is_available() and dist.is_initialized():
                    dist.all_reduce(p.grad)                          # what engine.BucketAllReduce does per bucket
                    p.grad /= dist.get_world_size()
                if p.grad is not None:
                    m = (self.optimizer.mults or {}).get(p, (1.0, 1.0))[0]
                    p -= self.optimizer.lr * m * p.grad
        self.optimizer.steps += 1
        self.tagged = {k: v.detach() for k, v in out.items() if k != 'loss'}
        return out['loss'].detach()


def _runner_cfg(tmp_path, **top):
    data = lambda n, seed, bs, shuffle: [{  # noqa: E731
        'dataloader': {'batch_size': bs, 'num_workers': 0, 'shuffle': shuffle, 'drop_last': False},
        'dataset': {'name': 'SyntheticImages',
                    'params': {'num_samples': n, 'shape': [2, 2, 2], 'num_classes': 3, 'seed': seed}}}]
    cfg = {
        'task': {'name': '_TinyTask', 'params': {'in_features': 8, 'num_classes': 3}},
        'joint_loss': {'losses': [{'name': '_CpuCrossEntropy', 'tag': 'ce',
                                   'mapping': {'input': 'prediction', 'target': 'target'}}]},
        'optimization': [{'optimizer': {'name': 'SGD', 'params': {'lr': 0.1}},
                          'scheduler': {'name': 'ExponentialLR', 'params': {'gamma': 0.5}}}],
        'data': {'TRAIN': data(40, 0, 8, True), 'VALID': data(20, 1, 8, False)},
        'trainer': {'max_epochs': 3, 'log_every_n_steps': 2},
        'seed_params': {'seed': 42},
        'logger': {'name': 'CSVLogger', 'log_dir': str(tmp_path), 'experiment_name': 'run'},
        'metrics': [{'name': 'Accuracy', 'params': {'task': 'multiclass', 'num_classes': 3},
                     'mapping': {'preds': 'prediction', 'target': 'target'}}],
        'callbacks': [{'name': 'ModelCheckpoint', 'params': {'monitor': 'valid/loss', 'mode': 'min', 'save_top_k': 1,
                                                             'save_last': True}},
                      {'name': 'FreezeUnfreeze', 'params': {'freeze_modules': [{'module_name': 'backbone', 'epoch': 1}]}},
                      {'name': 'TQDMProgressBar', 'params': {'refresh_rate': 5}}],
    }
    cfg.update(top)
    return cfg


def test_runner_epochs_logging_checkpoints_freeze_and_resume(tmp_path):
    r = Runner(_runner_cfg(tmp_path), loop_factory=_CpuLoop)
    w0 = r.task.backbone[0].weight.detach().clone()
    logs = r.fit()
    assert r.current_epoch == 3 and r.global_step == 15
    assert {'train/loss', 'train/ce', 'valid/loss', 'valid/ce', 'train/Accuracy', 'valid/Accuracy'} <= set(logs)
    assert logs['train/ce'] == pytest.approx(logs['train/loss'], rel=1e-6)       # one loss, weight normalised to 1
    assert 0.0 <= logs['valid/Accuracy'] <= 1.0
    # ExponentialLR per epoch: 0.1 -> 0.0125 after three epochs
    assert r.loop.optimizer.lr == pytest.approx(0.1 * 0.5 ** 3)
    # FreezeUnfreeze: frozen before training (graph reset 1), thawed at epoch 1 (graph reset 2)
    assert r.loop.graph_resets == 2
    assert all(p.requires_grad for p in r.task.parameters())
    assert not torch.equal(w0, r.task.backbone[0].weight)
    out_dir = os.path.join(str(tmp_path), 'run')
    rows = open(os.path.join(out_dir, 'metrics.csv')).read().strip().splitlines()
    assert rows[0].startswith('step,epoch') and 'valid/Accuracy' in rows[0] and len(rows) > 6
    ckpts = sorted(os.listdir(os.path.join(out_dir, 'checkpoints')))
    assert 'last.ckpt' in ckpts and len(ckpts) == 2                     # top-1 + last
    last = torch.load(os.path.join(out_dir, 'checkpoints', 'last.ckpt'), weights_only=False)
    assert last['epoch'] == 2 and last['global_step'] == 15 and 'optimizer_states' in last and 'state_dict' in last

    # the frozen epoch really froze: run one epoch only and compare the backbone
    r1 = Runner(_runner_cfg(tmp_path / 'one', trainer={'max_epochs': 1}), loop_factory=_CpuLoop)
    b0 = r1.task.backbone[0].weight.detach().clone()
    h0 = r1.task.head.weight.detach().clone()
    r1.fit()
    assert torch.equal(b0, r1.task.backbone[0].weight) and not torch.equal(h0, r1.task.head.weight)

    # resume_path: continues at epoch 3 with the stored lr / step, runs two more epochs
    cfg = _runner_cfg(tmp_path / 'resumed', resume_path=os.path.join(out_dir, 'checkpoints', 'last.ckpt'),
                      trainer={'max_epochs': 5})
    r2 = Runner(cfg, loop_factory=_CpuLoop)
    r2.fit()
    assert r2.current_epoch == 5 and r2.global_step == 25
    assert r2.loop.optimizer.lr == pytest.approx(0.1 * 0.5 ** 5)
    # load_checkpoint block: head excluded keeps its own init, backbone comes from the file
    cfg = _runner_cfg(tmp_path / 'lc', trainer={'max_epochs': 0})
    cfg['task']['load_checkpoint'] = {'base_ckpt_path': os.path.join(out_dir, 'checkpoints', 'last.ckpt'),
                                      'exclude_keys': ['head']}
    r3 = Runner(cfg, loop_factory=_CpuLoop)
    head0 = r3.task.head.weight.detach().clone()
    r3.fit()
    assert torch.equal(r3.task.backbone[0].weight, last['state_dict']['backbone.0.weight'])
    assert torch.equal(r3.task.head.weight, head0)


def test_runner_limits_max_steps_test_and_predict(tmp_path):
    cfg = _runner_cfg(tmp_path, trainer={'max_epochs': 10, 'max_steps': 7, 'limit_val_batches': 1})
    cfg['callbacks'] = []
    cfg['data']['TEST'] = cfg['data']['VALID']
    cfg['data']['PREDICT'] = cfg['data']['VALID']
    r = Runner(cfg, loop_factory=_CpuLoop)
    r.fit()
    assert r.global_step == 7 and r.should_stop
    logs = r.test()
    assert set(logs) == {'test/Accuracy'}
    preds = r.predict()
    assert len(preds) == 3 and preds[0]['prediction'].shape == (8, 3)
    with pytest.raises(ValueError, match='does not support'):
        r.run('find_lr')
    assert _limit(10, None) == 10 and _limit(10, 3) == 3 and _limit(10, 0.5) == 5 and _limit(10, 1.0) == 10
    bad = _runner_cfg(tmp_path / 'bad')
    bad['data']['VALID'][0]['dataloader']['drop_last'] = True
    with pytest.raises(ValueError, match='drop_last'):
        Runner(bad, loop_factory=_CpuLoop).fit()
    for key, value in (('accumulate_grad_batches', 4), ('gradient_clip_val', 1.0), ('sync_batchnorm', True)):
        with pytest.raises(NotImplementedError, match=key):
            Runner(_runner_cfg(tmp_path / key, trainer={'max_epochs': 1, key: value}), loop_factory=_CpuLoop)
    two = _runner_cfg(tmp_path / 'two')
    two['optimization'] = two['optimization'] * 2
    with pytest.raises(NotImplementedError):
        Runner(two, loop_factory=_CpuLoop).fit()


def test_runner_two_validation_loaders(tmp_path):
    """classification_cifar10_multi_validation.yaml layout: two VALID loaders, one metric per loader
    (`val_dataloader_idxs`), Lightning's `/dataloader_idx_i` suffix on the per-loader losses."""
    cfg = _runner_cfg(tmp_path, trainer={'max_epochs': 1})
    cfg['callbacks'] = []
    cfg['data']['VALID'] = cfg['data']['VALID'] + [dict(cfg['data']['VALID'][0])]
    acc = dict(cfg['metrics'][0], phases=['VALID'])
    cfg['metrics'] = [dict(acc, val_dataloader_idxs=[0]), dict(acc, tag='acc_second', val_dataloader_idxs=[1]),
                      dict(acc, tag='acc_both', val_dataloader_idxs=[0, 1])]
    logs = Runner(cfg, loop_factory=_CpuLoop).fit()
    for key in ('valid/loss/dataloader_idx_0', 'valid/loss/dataloader_idx_1', 'valid/Accuracy', 'valid/acc_second',
                'valid/acc_both_dataloader_0', 'valid/acc_both_dataloader_1'):
        assert key in logs, (key, sorted(logs))
    assert 'valid/loss' not in logs
    assert logs['valid/Accuracy'] == logs['valid/acc_second'] == logs['valid/acc_both_dataloader_1']   # same data twice
    assert logs['valid/loss/dataloader_idx_0'] == pytest.approx(logs['valid/loss/dataloader_idx_1'])


def test_model_checkpoint_top_k(tmp_path):
    class R:
        output_dir, current_epoch, global_step, has_validation = str(tmp_path), 0, 0, True
        saved, removed = [], []

        def save_checkpoint(self, path, weights_only=False):
            self.saved.append(os.path.basename(path))

        def remove_checkpoint(self, path):
            self.removed.append(os.path.basename(path))
    r = R()
    cb = ModelCheckpoint(monitor='valid/F1Score', mode='max', save_top_k=2, filename='{epoch}-best')
    cb.setup(r)
    for epoch, score in enumerate([0.1, 0.3, 0.2, 0.05, 0.4]):
        r.current_epoch, r.global_step = epoch, 10 * (epoch + 1)
        cb.on_validation_end(r, {'valid/F1Score': score})
    assert r.saved == ['0-best.ckpt', '1-best.ckpt', '2-best.ckpt', '4-best.ckpt']
    assert r.removed == ['0-best.ckpt', '2-best.ckpt']
    assert os.path.basename(cb.best_model_path) == '4-best.ckpt' and cb.best_model_score == 0.4
    with pytest.raises(KeyError):
        r.current_epoch = 9
        cb.on_validation_end(r, {'valid/other': 1.0})


# ------------------------------------------------------------------------------------------------ CLI + example YAML
CIFAR_YAML = textwrap.dedent('''
    task:
      name: ClassificationTask
      params:
        backbone_name: resnet18
        backbone_params: {pretrained: false, in_channels: 3}
        pooling_name: Pooling
        head_name: ClassificationHead
        head_params: {num_classes: &num_classes 10}
        inputs:
          - shape: [3, &height 32, &width 32]
            dtype: &input_dtype float16
    joint_loss:
      losses:
        - name: CrossEntropyLoss
          mapping: {input: prediction, target: target}
    optimization:
      - optimizer: {name: Adam, params: {lr: 0.0001}}
        scheduler: {name: ExponentialLR, params: {gamma: 0.97}}
    data:
      TRAIN:
        - dataloader: {batch_size: 128, num_workers: 8, drop_last: true, shuffle: true}
          dataset:
            name: CIFAR10
            params: {input_dtype: *input_dtype, train: true, download: true, data_folder: &folder '${oc.env:HOME}/.cache/torchok/cifar10/data'}
            transform:
              - &resize {name: Resize, params: {height: *height, width: *width}}
              - &normalize {name: Normalize, params: {mean: [0.485, 0.456, 0.406], std: [0.229, 0.224, 0.225]}}
              - &totensor {name: ToTensorV2}
    trainer: {accelerator: gpu, max_epochs: 30, precision: 16, num_sanity_val_steps: 0}
    seed_params: {seed: 42, workers: true}
    logger:
      log_dir: '${oc.env:HOME}/.cache/torchok/cifar10/logs'
      experiment_name: resnet18
      timestamp: '${now:%Y-%m-%d}/${now:%H-%M-%S}'
      name: TensorBoardLogger
    callbacks:
      - name: FreezeUnfreeze
        params:
          freeze_modules:
            - {module_name: backbone, epoch: 2}
            - {module_name: backbone, stages: 1}
    metrics:
      - name: Accuracy
        params: {task: multiclass, num_classes: 10}
        mapping: {preds: prediction, target: target}
''')


def test_cli_finds_config_applies_overrides_and_fails_loudly_without_gpu(tmp_path, monkeypatch):
    from torchok_b200.__main__ import entrypoint, find_config, parse_args
    (tmp_path / 'configs').mkdir()
    (tmp_path / 'configs' / 'classification_cifar10.yaml').write_text(CIFAR_YAML)
    monkeypatch.chdir(tmp_path)
    assert find_config('configs', 'classification_cifar10').endswith('classification_cifar10.yaml')
    assert find_config(str(tmp_path / 'configs'), 'classification_cifar10.yaml')
    with pytest.raises(FileNotFoundError):
        find_config('configs', 'nope')
    a = parse_args(['-cp', 'configs', '-cn', 'x', 'trainer.max_epochs=1', '+mode=test'])
    assert a.overrides == ['trainer.max_epochs=1', '+mode=test']
    cfg = tb.load_config(str(tmp_path / 'configs' / 'classification_cifar10.yaml'),
                         ['trainer.max_epochs=2', 'optimization.0.optimizer.params.lr=0.01'])
    assert cfg.trainer.max_epochs == 2 and cfg.optimization[0].optimizer.params.lr == 0.01
    assert cfg.data.TRAIN[0].dataset.transform[0].params.height == 32
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, tb._lib.TokLibraryError)):     # no CPU path
            entrypoint(['-cp', 'configs', '-cn', 'classification_cifar10', 'trainer.max_epochs=1'])


def test_transforms_and_cifar_reader(tmp_path):
    import pickle

    import numpy as np
    from torchok_b200.data import create_dataset, create_transforms
    rng = np.random.RandomState(0)
    root = tmp_path / 'cifar-10-batches-py'
    root.mkdir()
    for name in ['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5', 'test_batch']:
        with open(root / name, 'wb') as f:
            pickle.dump({'data': rng.randint(0, 256, (4, 3072), dtype=np.uint8), 'labels': list(rng.randint(0, 10, 4))}, f)
    with open(root / 'batches.meta', 'wb') as f:
        pickle.dump({'label_names': [f'c{i}' for i in range(10)]}, f)
    spec = {'name': 'CIFAR10', 'params': {'input_dtype': 'float16', 'train': True, 'download': True,
                                          'data_folder': str(tmp_path)},
            'transform': [{'name': 'Resize', 'params': {'height': 32, 'width': 32}},
                          {'name': 'Normalize', 'params': {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225]}},
                          {'name': 'ToTensorV2'}]}
    ds = create_dataset(spec)
    assert len(ds) == 20
    s = ds[3]
    assert s['image'].dtype == torch.float16 and s['image'].shape == (3, 32, 32) and s['index'] == 3
    raw = ds.images[3].astype(np.float32) / 255.0
    want = (raw - np.array([0.485, 0.456, 0.406], np.float32)) / np.array([0.229, 0.224, 0.225], np.float32)
    assert torch.allclose(s['image'].float(), torch.from_numpy(want.transpose(2, 0, 1)), atol=2e-3)
    assert int(s['target']) == int(ds.targets[3])
    with pytest.raises(RuntimeError, match='Dataset not found or corrupted'):
        create_dataset(dict(spec, params=dict(spec['params'], data_folder=str(tmp_path / 'missing'))))
    t = create_transforms([{'name': 'Compose', 'params': {'transforms': [{'name': 'HorizontalFlip', 'params': {'p': 1.0}},
                                                                          {'name': 'CenterCrop', 'params': {'height': 2, 'width': 2}}]}}])
    img = np.arange(16, dtype=np.uint8).reshape(4, 4, 1)
    out = t(image=img, mask=img[..., 0])
    assert out['image'][..., 0].tolist() == [[6, 5], [10, 9]] and out['mask'].tolist() == [[6, 5], [10, 9]]


# ------------------------------------------------------------------------------------------------ world size 2 (gloo)
class MetricMemoryBlock(Metric):
    """test_metric_manager_ddp.py:43-53: a list state gathered over ranks at compute time."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.add_state('memory_list', default=[], dist_reduce_fx=None)

    def update(self, state):
        self.memory_list.append(state)

    def compute(self):
        return torch.tensor(torch.cat(self.synced_states()['memory_list']).shape[0])


def _ddp_worker(rank, world, port, tmp, ret):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    if 'MetricMemoryBlock' not in tb.METRICS:
        tb.METRICS.register_class(MetricMemoryBlock)
    # (1) the reference's DDP metric case: 50 samples x 5 epochs sharded over the ranks
    mm = MetricsManager([dict(name='Accuracy', mapping=dict(preds='predict', target='target'), phases=['TRAIN'],
                              params=dict(task='multiclass', num_classes=10)),
                         dict(name='MetricMemoryBlock', mapping=dict(state='predict'), phases=['TRAIN'])])
    for _ in range(5):
        for i in range(rank * 4, 50, 4 * world):            # DistributedSampler-like interleaving of batches of 4
            mm.update(Phase.TRAIN, predict=torch.tensor(PREDICTS[i:i + 4]), target=torch.tensor(LABELS[i:i + 4]))
    out = {k: float(v) for k, v in mm.on_epoch_end(Phase.TRAIN).items()}
    # (2) the Runner under two ranks: sharded TRAIN data, gradients averaged by the stand-in loop, metrics / losses
    # combined over ranks, files written by rank 0 only
    cfg = _runner_cfg(tmp, trainer={'max_epochs': 2})
    r = Runner(cfg, loop_factory=_CpuLoop)
    logs = r.fit()
    ret[rank] = dict(metrics=out, logs={k: float(v) for k, v in logs.items()}, steps=r.global_step,
                     weight=r.task.head.weight.detach().clone(), wrote=os.path.exists(os.path.join(tmp, 'run', 'metrics.csv')))
    dist.barrier()
    dist.destroy_process_group()


def test_front_door_world_size_two_gloo(tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29300 + os.getpid() % 500
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, str(tmp_path), ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    a, b = ret[0], ret[1]
    # test_metric_manager_ddp.py:15-24,100-108: Accuracy 0.18, memory block = all samples of all epochs
    for r in (a, b):
        assert r['metrics']['train/Accuracy'] == pytest.approx(0.18, abs=1e-7)
        assert r['metrics']['train/MetricMemoryBlock'] == len(LABELS) * 5
    # 40 samples / (2 ranks x batch 8) = 3 steps per rank and epoch (DistributedSampler pads 20 -> 3 batches of 8,8,4)
    assert a['steps'] == b['steps'] == 6
    assert torch.equal(a['weight'], b['weight'])                       # replicas stay in step
    for k in ('train/loss', 'train/Accuracy', 'valid/loss', 'valid/Accuracy'):
        assert a['logs'][k] == pytest.approx(b['logs'][k], rel=1e-6), k
    assert a['wrote'] and b['wrote']                                   # same tmp dir: the file exists, written once
    rows = open(os.path.join(str(tmp_path), 'run', 'metrics.csv')).read().strip().splitlines()
    assert len(rows) == len(set(rows))                                 # no duplicated rows from a second writer


def test_file_backed_example_datasets(tmp_path):
    """ImageClassificationDataset / SOP / SweetPepper (what classification_imagenet.yaml, pairwise_sop.yaml and
    segmentation_sweet_pepper.yaml name) on small generated image files: annotation formats, target conventions,
    dtypes, missing-folder error."""
    import cv2
    import numpy as np
    from torchok_b200.data import create_dataset
    rng = np.random.RandomState(1)
    tf = [{'name': 'Resize', 'params': {'height': 8, 'width': 8}}, {'name': 'Normalize'}, {'name': 'ToTensorV2'}]

    def write(path, shape):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        img = rng.randint(0, 256, shape, dtype=np.uint8)
        assert cv2.imwrite(path, img)
        return img
    # --- csv classification (multiclass + multilabel)
    root = tmp_path / 'cls'
    imgs = [write(str(root / f'im{i}.png'), (10, 12, 3)) for i in range(3)]
    (root / 'ann.csv').write_text('image_path,label\nim0.png,2\nim1.png,0\nim2.png,1\n')
    ds = create_dataset({'name': 'ImageClassificationDataset', 'transform': tf,
                         'params': {'data_folder': str(root), 'annotation_path': 'ann.csv', 'num_classes': 3,
                                    'input_dtype': 'float16'}})
    s = ds[0]
    assert len(ds) == 3 and s['image'].shape == (3, 8, 8) and s['image'].dtype == torch.float16
    assert s['target'].dtype == torch.long and int(s['target']) == 2 and s['index'] == 0
    raw = ds.get_raw(1)['image']
    assert raw.shape == (10, 12, 3) and np.array_equal(raw, imgs[1][..., ::-1])     # stored BGR by cv2, returned RGB
    (root / 'multi.csv').write_text('image_path,label\nim0.png,0 2\nim1.png,1\n')
    ml = create_dataset({'name': 'ImageClassificationDataset', 'transform': tf,
                         'params': {'data_folder': str(root), 'annotation_path': 'multi.csv', 'num_classes': 3,
                                    'multilabel': True, 'target_dtype': 'float32'}})
    assert ml[0]['target'].tolist() == [1.0, 0.0, 1.0]
    with pytest.raises(ValueError, match='more than num_classes'):
        create_dataset({'name': 'ImageClassificationDataset', 'transform': tf,
                        'params': {'data_folder': str(root), 'annotation_path': 'ann.csv', 'num_classes': 2}})
    with pytest.raises(ValueError, match='annotation_path'):
        create_dataset({'name': 'ImageClassificationDataset', 'transform': tf, 'params': {'data_folder': str(root)}})
    # --- SOP: space-separated txt, zero-based targets per split
    sop = tmp_path / 'Stanford_Online_Products'
    write(str(sop / 'bicycle_final' / 'a.JPG'), (9, 9, 3))
    write(str(sop / 'chair_final' / 'b.JPG'), (9, 9))                                # gray file -> RGB
    (sop / 'Ebay_train.txt').write_text('image_id class_id super_class_id path\n1 1 1 bicycle_final/a.JPG\n2 7 2 chair_final/b.JPG\n')
    (sop / 'Ebay_test.txt').write_text('image_id class_id super_class_id path\n3 11319 1 bicycle_final/a.JPG\n4 11325 2 chair_final/b.JPG\n')
    tr = create_dataset({'name': 'SOP', 'transform': tf, 'params': {'train': True, 'download': True, 'data_folder': str(tmp_path)}})
    te = create_dataset({'name': 'SOP', 'transform': tf, 'params': {'train': False, 'download': False, 'data_folder': str(tmp_path)}})
    assert [int(tr[i]['target']) for i in range(2)] == [0, 6] and [int(te[i]['target']) for i in range(2)] == [0, 6]
    assert tr[1]['image'].shape == (3, 8, 8)
    with pytest.raises(RuntimeError, match='Dataset not found or corrupted'):
        create_dataset({'name': 'SOP', 'transform': tf, 'params': {'train': True, 'download': True,
                                                                 'data_folder': str(tmp_path / 'nowhere')}})
    # --- SweetPepper: image + mask files, mask resized with nearest neighbour, int64 target
    sp = tmp_path / 'sweet_pepper'
    write(str(sp / 'img' / '0.png'), (16, 16, 3))
    mask = (rng.randint(0, 3, (16, 16))).astype(np.uint8)
    os.makedirs(str(sp / 'msk'), exist_ok=True)
    assert cv2.imwrite(str(sp / 'msk' / '0.png'), mask)
    for name in ('train.csv', 'valid.csv'):
        (sp / name).write_text('image_path,mask\nimg/0.png,msk/0.png\n')
    seg = create_dataset({'name': 'SweetPepper', 'transform': tf,
                          'params': {'train': True, 'download': True, 'data_folder': str(tmp_path)}})
    s = seg[0]
    assert s['image'].shape == (3, 8, 8) and s['target'].shape == (8, 8) and s['target'].dtype == torch.int64
    assert set(s['target'].unique().tolist()) <= {0, 1, 2} and 'mask' not in s
    want = cv2.resize(mask, (8, 8), interpolation=0)
    assert torch.equal(s['target'], torch.from_numpy(want).long())
    # --- FancyPCA keeps shape / dtype, changes colours, leaves masks alone
    from torchok_b200.data import FancyPCA
    np.random.seed(0)
    img = rng.randint(0, 256, (12, 12, 3), dtype=np.uint8)
    out = FancyPCA(alpha=0.5, p=1.0)(image=img, mask=mask)
    assert out['image'].shape == img.shape and out['image'].dtype == np.uint8 and not np.array_equal(out['image'], img)
    assert np.array_equal(out['mask'], mask)
    assert 'ModelCheckpointWithOnnx' in tb.CALLBACKS


REF_CONFIGS = '/root/reference/examples/configs'


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason='the reference checkout only exists in the build container')
@pytest.mark.parametrize('name,task_cls', [('classification_cifar10', 'ClassificationTask'),
                                           ('classification_cifar10_multi_validation', 'ClassificationTask'),
                                           ('classification_imagenet', 'ClassificationTask'),
                                           ('pairwise_sop', 'PairwiseLearnTask'),
                                           ('segmentation_sweet_pepper', 'SegmentationTask')])
def test_reference_example_configs_drop_in(name, task_cls, monkeypatch):
    """The reference's own example YAMLs, unchanged: the file loads (anchors, ${oc.env:…}, ${now:…}), every name
    resolves in the registry its position implies, the task assembles from `task.params`, and the metrics / callbacks /
    scheduler blocks construct.  (Datasets are not instantiated: their files are not on this box.)"""
    import importlib.util
    monkeypatch.setenv('HOME', '/root')
    spec = importlib.util.spec_from_file_location(
        '_cov', os.path.join(os.path.dirname(__file__), '..', 'scripts', 'config_coverage.py'))
    cov = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cov)
    cfg = tb.load_config(os.path.join(REF_CONFIGS, name + '.yaml'))
    missing = [(reg, n) for reg, n in cov.names_of(cfg) if n is not None and n not in getattr(tb, reg)]
    assert not missing, missing
    task = tb.TASKS.get(cfg.task.name)(cfg, **cfg.task.params)
    assert type(task).__name__ == task_cls and sum(p.numel() for p in task.parameters()) > 1e6
    MetricsManager(cfg.get('metrics') or [])
    for c in cfg.get('callbacks') or []:
        tb.CALLBACKS.get(c['name'])(**dict(c.get('params') or {}))
    for o in cfg.optimization:
        assert o.optimizer.name in tb.OPTIMIZERS
        if o.get('scheduler'):
            params = dict(o.scheduler.get('params') or {})
            LrDriver(_FakeOpt(o.optimizer.params.lr), o.scheduler.name, params, o.scheduler.get('pl_params'))


def test_jaccard_index_with_in_range_ignore_index():
    """torchmetrics 0.11.4 `_jaccard_index_reduce` takes an in-range ignore_index out of the score (macro:
    weight 0; micro: its denominator subtracted) — segmentation_sweet_pepper.yaml uses num_classes 3 / ignore_index 0."""
    import torch
    from torchok_b200.metrics.metrics_manager import JaccardIndex
    t, p = torch.tensor([0, 0, 1, 1, 2, 2]), torch.tensor([0, 1, 1, 1, 2, 1])
    m = JaccardIndex(num_classes=3, ignore_index=0)
    m.update(p, t)
    # the two ignored targets are dropped: class 1 = 2/3, class 2 = 1/2, class 0 weight 0
    assert abs(float(m.compute()) - (2 / 3 + 1 / 2) / 2) < 1e-6
    m = JaccardIndex(num_classes=3, ignore_index=0, average='micro')
    m.update(p, t)
    assert abs(float(m.compute()) - 3 / 5) < 1e-6          # tp 3 / (unions 3 + 2; class 0's denominator is taken out)
    m = JaccardIndex(num_classes=3, ignore_index=255)      # out of range: every class counts
    m.update(p, t)
    assert abs(float(m.compute()) - (1 / 2 + 2 / 4 + 1 / 2) / 3) < 1e-6


def test_optimizer_rejects_arithmetic_changing_keys():
    from torchok_b200.engine import _check_optimizer_kwargs
    _check_optimizer_kwargs('Adam', {'foreach': True, 'amsgrad': False, 'fused': None})
    with pytest.raises(NotImplementedError):
        _check_optimizer_kwargs('Adam', {'amsgrad': True})
    with pytest.raises(NotImplementedError):
        _check_optimizer_kwargs('SGD', {'maximize': True})
