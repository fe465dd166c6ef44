"""CALLBACKS registry (torchok/callbacks/__init__.py:1-23) for the stream-loop runner.

Lightning is replaced by runner.Runner, so a callback here is a plain object with optional hooks

    setup(runner)  on_train_epoch_start(runner)  on_train_epoch_end(runner, logs)  on_validation_end(runner, logs)
    teardown(runner)  state_dict()  load_state_dict(state)

Built: `FreezeUnfreeze` (torchok/callbacks/freeze_unfreeze.py:51-184, the policy the example configs use through
`get_stages`) and `ModelCheckpoint` (monitor / mode / save_top_k / save_last / dirpath / save_weights_only, the
options of examples/configs/*.yaml).  Progress bars, model summaries, loggers' finalizers and the other Lightning
callbacks the reference re-exports are accepted by name and do nothing (`_Accepted`), so existing YAMLs load.
"""
import inspect
import os

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from ..constructor import CALLBACKS


class Callback:
    def setup(self, runner): ...
    def on_train_epoch_start(self, runner): ...
    def on_train_epoch_end(self, runner, logs): ...
    def on_validation_end(self, runner, logs): ...
    def teardown(self, runner): ...

    def state_dict(self):
        return {}

    def load_state_dict(self, state): ...


# ------------------------------------------------------------------------------------------------ FreezeUnfreeze
def get_modules(module_dict, module):
    """Modules a freeze rule addresses (freeze_unfreeze.py:13-47): `module_name` is a dotted path of child names
    ('' = the whole task), `stages` narrows it through `get_stages(int)`, `module_class` keeps the sub-modules whose
    class or any base class carries that NAME."""
    name = module_dict['module_name']
    target = module
    if name != '':
        for part in name.split('.'):
            target = dict(target.named_children()).get(part)
            if target is None:
                raise ValueError(f'Module `{name}` is not found')
    if 'stages' in module_dict:
        if not hasattr(target, 'get_stages'):
            raise ValueError(f'You specified `stages` in `{name}` but this module does not have `get_stages` method')
        target = target.get_stages(module_dict['stages'])
    if 'module_class' in module_dict:
        wanted = module_dict['module_class']
        picked = nn.ModuleList(m for m in target.modules()
                               if wanted in [c.__name__ for c in inspect.getmro(type(m))])
        if len(picked) == 0:
            raise ValueError(f'Module `{name}` does not have submodules of `{wanted}` type.')
        target = picked
    return target


def _leaves(modules):
    """Lightning's BaseFinetuning.flatten_modules: every module without children, plus modules that own parameters
    directly."""
    if isinstance(modules, nn.Module):
        modules = [modules]
    out, seen = [], set()
    for top in modules:
        for m in top.modules():
            if (not list(m.children()) or m._parameters) and id(m) not in seen:
                seen.add(id(m))
                out.append(m)
    return out


@CALLBACKS.register_class
class FreezeUnfreeze(Callback):
    """Freeze modules before training and thaw them when their `epoch` is reached (rules without `epoch` stay frozen).

    Same rule keys, ordering (`top_down_freeze_order`) and BatchNorm handling (`bn_requires_grad`,
    `bn_track_running_stats`) as the reference.  In the reference every parameter already sits in an optimizer group
    (Constructor.add_params adds frozen ones too), so thawing adds no new group; here the arena optimizer keeps all
    parameters as well and a frozen one has its learning-rate and weight-decay multipliers set to zero — its gradient
    is not computed (the wgrad kernels are skipped on `requires_grad=False`) and its value cannot drift through
    momentum or decay.  Every change of the frozen set drops the captured step graph so the next step recaptures it.
    """

    def __init__(self, freeze_modules, top_down_freeze_order=True):
        self.freeze_modules = sorted((dict(m) for m in freeze_modules), key=lambda m: m['module_name'],
                                     reverse=not top_down_freeze_order)

    @staticmethod
    def make_trainable(modules):
        for m in _leaves(modules):
            if isinstance(m, _BatchNorm):
                m.track_running_stats = True
            for p in m.parameters(recurse=False):
                p.requires_grad = True

    @staticmethod
    def freeze(modules, module_dict):
        for m in _leaves(modules):
            if isinstance(m, _BatchNorm):
                for p in m.parameters(recurse=False):
                    p.requires_grad = module_dict.get('bn_requires_grad', False)
                m.track_running_stats = module_dict.get('bn_track_running_stats', False)
            else:
                for p in m.parameters(recurse=False):
                    p.requires_grad = False

    def apply(self, task, current_epoch):
        """State of the task's `requires_grad` / `track_running_stats` flags for `current_epoch`
        (freeze_before_training at epoch 0 with nothing to thaw, finetune_function afterwards)."""
        for rule in self.freeze_modules:
            if 'epoch' in rule and rule['epoch'] <= current_epoch:
                self.make_trainable(get_modules(rule, task))
        for rule in self.freeze_modules:
            if 'epoch' not in rule or rule['epoch'] > current_epoch:
                self.freeze(get_modules(rule, task), rule)

    # -- runner hooks
    def setup(self, runner):
        for rule in self.freeze_modules:
            self.freeze(get_modules(rule, runner.task), rule)
        runner.frozen_set_changed()

    def on_train_epoch_start(self, runner):
        before = [p.requires_grad for p in runner.task.parameters()]
        stats = [m.track_running_stats for m in runner.task.modules() if isinstance(m, _BatchNorm)]
        self.apply(runner.task, runner.current_epoch)
        if before != [p.requires_grad for p in runner.task.parameters()] or \
                stats != [m.track_running_stats for m in runner.task.modules() if isinstance(m, _BatchNorm)]:
            runner.frozen_set_changed()


# ------------------------------------------------------------------------------------------------ ModelCheckpoint
@CALLBACKS.register_class
class ModelCheckpoint(Callback):
    """The subset of pytorch_lightning.callbacks.ModelCheckpoint the example configs use: after every validation (or
    training epoch when there is no validation data) keep the `save_top_k` best checkpoints by `monitor` (`mode`
    'min' | 'max'; -1 keeps all, 0 none) and, with `save_last`, `last.ckpt`.  Without `monitor` only the most recent
    epoch's file is kept (Lightning's default)."""

    def __init__(self, dirpath=None, filename=None, monitor=None, mode='min', save_top_k=1, save_last=None,
                 save_weights_only=False, every_n_epochs=1, verbose=False, **unused):
        if mode not in ('min', 'max'):
            raise ValueError(f'`mode` can be min, max, got {mode}')
        self.dirpath, self.filename, self.monitor, self.mode = dirpath, filename, monitor, mode
        self.save_top_k, self.save_last, self.save_weights_only = save_top_k, save_last, save_weights_only
        self.every_n_epochs = max(int(every_n_epochs or 1), 1)
        self.best_k = {}            # path -> score
        self.best_model_path, self.best_model_score, self.last_model_path = '', None, ''
        self._saved_epoch = -1

    def setup(self, runner):
        if self.dirpath is None:
            self.dirpath = os.path.join(runner.output_dir, 'checkpoints')

    def _name(self, runner, logs):
        if self.filename:
            fields = {'epoch': runner.current_epoch, 'step': runner.global_step}
            fields.update({k: float(v) for k, v in logs.items()})
            try:
                return self.filename.format(**fields) + '.ckpt'
            except (KeyError, IndexError):
                pass
        return f'epoch={runner.current_epoch}-step={runner.global_step}.ckpt'

    def _better(self, a, b):
        return a < b if self.mode == 'min' else a > b

    def _save(self, runner, logs):
        if runner.current_epoch == self._saved_epoch or (runner.current_epoch + 1) % self.every_n_epochs:
            return
        self._saved_epoch = runner.current_epoch
        if self.save_top_k != 0:
            path = os.path.join(self.dirpath, self._name(runner, logs))
            if self.monitor is None:
                for old in list(self.best_k):
                    if self.save_top_k > 0 and old != path:
                        runner.remove_checkpoint(old)
                        del self.best_k[old]
                runner.save_checkpoint(path, weights_only=self.save_weights_only)
                self.best_k[path] = None
                self.best_model_path = path
            elif self.monitor in logs:
                score = float(logs[self.monitor])
                full = self.save_top_k > 0 and len(self.best_k) >= self.save_top_k
                worst = (max if self.mode == 'min' else min)(self.best_k, key=self.best_k.get) if self.best_k else None
                if not full or self._better(score, self.best_k[worst]):
                    if full:
                        runner.remove_checkpoint(worst)
                        del self.best_k[worst]
                    runner.save_checkpoint(path, weights_only=self.save_weights_only)
                    self.best_k[path] = score
                    best = (min if self.mode == 'min' else max)(self.best_k, key=self.best_k.get)
                    self.best_model_path, self.best_model_score = best, self.best_k[best]
            else:
                raise KeyError(f'ModelCheckpoint(monitor={self.monitor!r}) could not find the monitored key in the '
                               f'returned metrics: {sorted(logs)}')
        if self.save_last:
            self.last_model_path = os.path.join(self.dirpath, 'last.ckpt')
            runner.save_checkpoint(self.last_model_path, weights_only=self.save_weights_only)

    def on_validation_end(self, runner, logs):
        self._save(runner, logs)

    def on_train_epoch_end(self, runner, logs):
        if not runner.has_validation:
            self._save(runner, logs)

    def state_dict(self):
        return {'best_k': dict(self.best_k), 'best_model_path': self.best_model_path,
                'best_model_score': self.best_model_score, 'last_model_path': self.last_model_path}

    def load_state_dict(self, state):
        self.best_k = dict(state.get('best_k', {}))
        self.best_model_path = state.get('best_model_path', '')
        self.best_model_score = state.get('best_model_score')
        self.last_model_path = state.get('last_model_path', '')


@CALLBACKS.register_class
class EarlyStopping(Callback):
    """monitor / mode / patience / min_delta of pytorch_lightning.callbacks.EarlyStopping; sets `runner.should_stop`."""

    def __init__(self, monitor, min_delta=0.0, patience=3, mode='min', strict=True, **unused):
        self.monitor, self.min_delta, self.patience, self.mode, self.strict = monitor, abs(min_delta), patience, mode, strict
        self.best, self.wait = None, 0

    def on_validation_end(self, runner, logs):
        if self.monitor not in logs:
            if self.strict:
                raise RuntimeError(f'Early stopping conditioned on metric `{self.monitor}` which is not available. '
                                   f'Available metrics are: {sorted(logs)}')
            return
        score = float(logs[self.monitor])
        improved = self.best is None or (score < self.best - self.min_delta if self.mode == 'min'
                                         else score > self.best + self.min_delta)
        if improved:
            self.best, self.wait = score, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                runner.should_stop = True

    def state_dict(self):
        return {'best': self.best, 'wait': self.wait}

    def load_state_dict(self, state):
        self.best, self.wait = state.get('best'), state.get('wait', 0)


@CALLBACKS.register_class
class CheckpointONNX(ModelCheckpoint):
    """torchok/callbacks/checkpoint_onnx.py:13-83 writes an .onnx file next to each kept checkpoint.  The modules here
    execute hand-written kernels through a C ABI, which torch.onnx cannot trace, so this keeps the ModelCheckpoint
    behaviour (the .ckpt files load into the reference's torch.nn modules: same state-dict keys) and says once that
    the ONNX file is not produced."""

    def __init__(self, *args, onnx_params=None, remove_head=False, export_to_onnx=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.onnx_params, self.remove_head = onnx_params or {}, remove_head

    def setup(self, runner):
        super().setup(runner)
        import warnings
        warnings.warn('CheckpointONNX: ONNX export is not available for the sm_100a kernel modules; '
                      'checkpoints (.ckpt) are written as with ModelCheckpoint')


@CALLBACKS.register_class
class ModelCheckpointWithOnnx(CheckpointONNX):
    """Name used by examples/configs/segmentation_sweet_pepper.yaml:145 and representation_arcface_sop.yaml:161 (the
    reference itself registers only `CheckpointONNX`, so those two files fail there with a KeyError); same behaviour
    as CheckpointONNX here."""


def _accepted(name):
    cls = type(name, (Callback,), {'__init__': lambda self, *a, **k: None,
                                   '__doc__': f'{name}: accepted for config compatibility; no effect in the stream loop.'})
    return cls


for _n in ('FinalizeLogger', 'TQDMProgressBar', 'RichProgressBar', 'ModelSummary', 'RichModelSummary',
           'LearningRateMonitor', 'DeviceStatsMonitor', 'Timer'):
    CALLBACKS.register_class(_accepted(_n))
del _n
