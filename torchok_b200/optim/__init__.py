"""OPTIMIZERS / SCHEDULERS registries (torchok/optim/optimizers/__init__.py:1-19, optim/schedulers/__init__.py:1-30).

Optimizers: the reference registers torch.optim classes and steps them per parameter group from Lightning.  Here the
names resolve to builders of the flat-arena step kernels (engine.ArenaSGD / ArenaAdam: ONE launch updates every
parameter, refreshes the bf16 shadow weights and clears the gradients); names without an arena kernel raise
NotImplementedError when built, not a silent torch.optim fallback.

Schedulers: the torch.optim.lr_scheduler classes themselves (pure host arithmetic on a learning rate).  They need a
torch.optim.Optimizer to hold `param_groups`; `LrDriver` gives them a one-parameter stand-in and forwards the resulting
learning rate to the arena optimizer's device-side lr cell, which the captured step graph reads.  timm's scheduler
family (CosineLRScheduler, ...) is not in this image and is not registered.
"""
import torch
from torch.optim import lr_scheduler as _sched

from ..constructor import OPTIMIZERS, SCHEDULERS

ARENA_OPTIMIZERS = ('SGD', 'Adam', 'AdamW')
_UNBUILT = ('Adadelta', 'Adagrad', 'Adamax', 'ASGD', 'LBFGS', 'RMSprop', 'Rprop', 'SparseAdam')


def _arena_builder(name):
    def build(arena, module=None, paramwise_cfg=None, **params):
        from ..engine import build_optimizer
        return build_optimizer(arena, name, params, module, paramwise_cfg)
    build.__name__ = name
    return build


def _unbuilt(name):
    def build(*args, **kwargs):
        raise NotImplementedError(f'optimizer {name}: the arena step kernels cover {", ".join(ARENA_OPTIMIZERS)}')
    build.__name__ = name
    return build


for _n in ARENA_OPTIMIZERS:
    OPTIMIZERS.register_class(_arena_builder(_n))
for _n in _UNBUILT:
    OPTIMIZERS.register_class(_unbuilt(_n))

for _n in ('LambdaLR', 'MultiplicativeLR', 'StepLR', 'MultiStepLR', 'ExponentialLR', 'CosineAnnealingLR',
           'ReduceLROnPlateau', 'CyclicLR', 'OneCycleLR', 'CosineAnnealingWarmRestarts'):
    SCHEDULERS.register_class(getattr(_sched, _n))
del _n


class LrDriver:
    """Runs a registered scheduler against an arena optimizer.

    `pl_params` follows SchedulerPLParams (config_structure.py:25-33): `interval` 'epoch' | 'step', `frequency`,
    `monitor` (ReduceLROnPlateau).  `step_end()` / `epoch_end(logs)` are called by the runner."""

    def __init__(self, optimizer, name, params=None, pl_params=None):
        self.optimizer = optimizer
        self.proxy = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=optimizer.lr)
        self.scheduler = SCHEDULERS.get(name)(self.proxy, **dict(params or {}))
        pl = dict(pl_params or {})
        self.interval = pl.get('interval', 'epoch') or 'epoch'
        self.frequency = int(pl.get('frequency', 1) or 1)
        self.monitor = pl.get('monitor', 'val_loss')
        self.strict = pl.get('strict', True)
        self._count = {'step': 0, 'epoch': 0}
        self.on_plateau = isinstance(self.scheduler, _sched.ReduceLROnPlateau)
        self._push()

    def _push(self):
        lr = float(self.proxy.param_groups[0]['lr'])
        if lr != self.optimizer.lr:
            self.optimizer.lr = lr

    def _tick(self, kind, logs=None):
        if self.interval != kind:
            return
        self._count[kind] += 1
        if self._count[kind] % self.frequency:
            return
        if isinstance(self.scheduler, _sched.ReduceLROnPlateau):
            if logs is None or self.monitor not in logs:
                if self.strict:
                    raise KeyError(f'ReduceLROnPlateau conditioned on metric {self.monitor} which is not available. '
                                   f'Available metrics are: {sorted(logs or {})}')
                return
            self.scheduler.step(float(logs[self.monitor]))
        else:
            self.proxy.step()        # keeps torch's "optimizer.step() before lr_scheduler.step()" bookkeeping quiet
            self.scheduler.step()
        self._push()

    def step_end(self, logs=None):
        self._tick('step', logs)

    def epoch_end(self, logs=None):
        self._tick('epoch', logs)

    def state_dict(self):
        return {'scheduler': self.scheduler.state_dict(), 'count': dict(self._count),
                'lr': float(self.proxy.param_groups[0]['lr'])}

    def load_state_dict(self, state):
        self.scheduler.load_state_dict(state['scheduler'])
        self._count = dict(state['count'])
        self.proxy.param_groups[0]['lr'] = state['lr']
        self._push()
