"""MetricsManager and the classification / segmentation meters the example configs name.

Host logic around the hot path: the reference wraps torchmetrics 0.11.4 objects
(torchok/metrics/__init__.py:3-31) in `MetricWithUtils` / `MetricsManager` (torchok/metrics/metrics_manager.py:12-206),
routes task outputs to them through the YAML `mapping`, and logs `<phase>/<name>` values at epoch end.  torchmetrics is
not in this image, so this file carries

* `Metric` — the small part of torchmetrics.Metric the reference relies on: `add_state(name, default, dist_reduce_fx)`,
  `update`, `compute`, `reset`, and synchronisation of the states over `torch.distributed` ranks at compute time
  ('sum' states are all-reduced, list / 'cat' states all-gathered);
* `Accuracy`, `F1Score`, `JaccardIndex`, `Precision`, `Recall` for `task: multiclass` (what
  examples/configs/classification_*.yaml and segmentation_*.yaml use), all derived from ONE confusion matrix that is
  accumulated where the predictions live (on the GPU: a `bincount` over `target * C + pred`, no host sync per step);
* `MetricsManager` with the reference's naming / phase / dataloader-index rules and error messages.

Pinned by the reference's own tests: tests/base_tests/metrics/metric_manager/test_metric_manager.py:100-185 (names,
tags, dict outputs, non-numeric results) and test_metric_manager_ddp.py:15-24 (Accuracy = 0.18), mirrored in
tests/test_front_door.py.  The averaging rules follow torchmetrics 0.11.4 (micro default for Accuracy / F1 /
Precision / Recall, macro default for JaccardIndex; macro skips classes with no support and no predictions for the
stat-score metrics) and are cross-checked against scikit-learn in the tests — torchmetrics itself cannot be executed
here, so those rules are otherwise unpinned.
"""
import numbers
from enum import Enum

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from ..constructor import METRICS


class Phase(Enum):
    """torchok/constructor/config_structure.py:7-11."""
    TRAIN = 'train'
    VALID = 'valid'
    TEST = 'test'
    PREDICT = 'predict'


def as_phase(p):
    if isinstance(p, Phase):
        return p
    return Phase[str(p).upper()] if str(p).upper() in Phase.__members__ else Phase(str(p).lower())


class Metric(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self._defaults = {}
        self._reductions = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        if not (isinstance(default, torch.Tensor) or (isinstance(default, list) and not default)):
            raise ValueError('state variable must be a tensor or an empty list (where you can append tensors)')
        self._defaults[name] = default.clone() if isinstance(default, torch.Tensor) else []
        self._reductions[name] = dist_reduce_fx
        setattr(self, name, default.clone() if isinstance(default, torch.Tensor) else [])

    def reset(self):
        for name, default in self._defaults.items():
            cur = getattr(self, name)
            if isinstance(default, torch.Tensor):
                setattr(self, name, default.clone().to(cur.device if isinstance(cur, torch.Tensor) else default.device))
            else:
                setattr(self, name, [])

    def update(self, *args, **kwargs):
        raise NotImplementedError

    def compute(self):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        self.update(*args, **kwargs)

    # ------------------------------------------------------------------------------------------------------------
    def synced_states(self):
        """{name: value} with the states combined over ranks (identity when not distributed)."""
        out = {n: getattr(self, n) for n in self._defaults}
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return out
        world = dist.get_world_size()
        for name, value in out.items():
            red = self._reductions[name]
            if isinstance(value, list):
                local = torch.cat([v.reshape(-1, *v.shape[1:]) if v.dim() else v.reshape(1) for v in value]) \
                    if value else torch.zeros(0)
                sizes = [None] * world
                dist.all_gather_object(sizes, tuple(local.shape))
                parts = []
                for r, shape in enumerate(sizes):
                    buf = local if r == dist.get_rank() else torch.empty(shape, dtype=local.dtype, device=local.device)
                    dist.broadcast(buf, r)
                    parts.append(buf)
                out[name] = [torch.cat(parts)] if parts else []
            elif red == 'sum':
                t = value.clone()
                dist.all_reduce(t)
                out[name] = t
            elif red in ('cat', None):
                parts = [torch.empty_like(value) for _ in range(world)]
                dist.all_gather(parts, value.contiguous())
                out[name] = torch.stack(parts) if red is None else torch.cat([p.reshape(-1) for p in parts])
            else:
                raise ValueError(f'unsupported dist_reduce_fx {red!r}')
        return out


# ---------------------------------------------------------------------------------------------- confusion matrix
def _labels_from(preds, target, num_classes):
    """(pred labels, target labels) flattened; float `preds` with a class dimension (N, C, ...) are arg-maxed over
    dim 1 (torchmetrics `_multiclass_*_format`), integer `preds` are labels already."""
    if preds.is_floating_point():
        if preds.dim() == target.dim() + 1:
            preds = preds.argmax(dim=1)
        else:
            raise ValueError('float predictions must have one more dimension (classes, dim 1) than the target')
    elif preds.dim() != target.dim():
        raise ValueError('integer predictions must have the shape of the target')
    return preds.reshape(-1).long(), target.reshape(-1).long()


class _ConfusionMetric(Metric):
    def __init__(self, task='multiclass', num_classes=None, average='micro', ignore_index=None, top_k=1,
                 multidim_average='global', validate_args=True, **kwargs):
        super().__init__()
        if task != 'multiclass':
            raise NotImplementedError(f'{type(self).__name__}: only task="multiclass" is built (the example configs '
                                      f'use nothing else); got {task!r}')
        if not isinstance(num_classes, int) or num_classes < 2:
            raise ValueError('num_classes must be an integer larger than 1')
        if top_k != 1 or multidim_average != 'global':
            raise NotImplementedError('top_k=1 and multidim_average="global" only')
        if average not in ('micro', 'macro', 'weighted', 'none', None):
            raise ValueError(f'average={average!r}')
        self.num_classes, self.average, self.ignore_index = num_classes, average, ignore_index
        self.add_state('confmat', torch.zeros(num_classes, num_classes, dtype=torch.long), dist_reduce_fx='sum')

    def update(self, preds, target):
        p, t = _labels_from(preds, target, self.num_classes)
        if self.ignore_index is not None:
            keep = t != self.ignore_index
            p, t = p[keep], t[keep]
        c = self.num_classes
        if self.confmat.device != p.device:
            self.confmat = self.confmat.to(p.device)
        self.confmat += torch.bincount(t * c + p, minlength=c * c)[:c * c].reshape(c, c)

    def _stats(self):
        cm = self.synced_states()['confmat'].double()
        tp = cm.diag()
        fp = cm.sum(0) - tp
        fn = cm.sum(1) - tp
        return cm, tp, fp, fn

    @staticmethod
    def _safe_div(num, den):
        return torch.where(den == 0, torch.zeros_like(num), num / torch.where(den == 0, torch.ones_like(den), den))

    def _reduce_scores(self, score, tp, fp, fn, skip_absent=True):
        if self.average in (None, 'none'):
            return score.float()
        if self.average == 'weighted':
            w = tp + fn
        else:
            w = torch.ones_like(score)
            if skip_absent:
                w[(tp + fp + fn) == 0] = 0.0
        return (self._safe_div(w * score, w.sum())).sum().float()


@METRICS.register_class
class Accuracy(_ConfusionMetric):
    def compute(self):
        cm, tp, fp, fn = self._stats()
        if self.average == 'micro':
            return self._safe_div(tp.sum(), cm.sum()).float()
        return self._reduce_scores(self._safe_div(tp, tp + fn), tp, fp, fn)


@METRICS.register_class
class Recall(Accuracy):
    """Multiclass recall = per-class accuracy; the micro average equals the accuracy."""


@METRICS.register_class
class Precision(_ConfusionMetric):
    def compute(self):
        cm, tp, fp, fn = self._stats()
        if self.average == 'micro':
            return self._safe_div(tp.sum(), (tp + fp).sum()).float()
        return self._reduce_scores(self._safe_div(tp, tp + fp), tp, fp, fn)


@METRICS.register_class
class F1Score(_ConfusionMetric):
    def compute(self):
        cm, tp, fp, fn = self._stats()
        if self.average == 'micro':
            tp, fp, fn = tp.sum(), fp.sum(), fn.sum()
            return self._safe_div(2 * tp, 2 * tp + fp + fn).float()
        return self._reduce_scores(self._safe_div(2 * tp, 2 * tp + fp + fn), tp, fp, fn)


@METRICS.register_class
class JaccardIndex(_ConfusionMetric):
    def __init__(self, task='multiclass', num_classes=None, average='macro', ignore_index=None, **kwargs):
        super().__init__(task=task, num_classes=num_classes, average=average, ignore_index=ignore_index, **kwargs)

    def compute(self):
        cm, tp, fp, fn = self._stats()
        denom = tp + fp + fn
        # torchmetrics 0.11.4 `_jaccard_index_reduce(confmat, average, ignore_index)`: an IN-RANGE ignore_index (e.g.
        # `num_classes: 3, ignore_index: 0` of examples/configs/segmentation_sweet_pepper.yaml) is taken out of the
        # score: its denominator is subtracted for 'micro', its weight is zero for 'macro'
        ign = self.ignore_index if (self.ignore_index is not None and 0 <= self.ignore_index < self.num_classes) else None
        if self.average == 'micro':
            total = denom.sum() - (denom[ign] if ign is not None else 0.0)
            return self._safe_div(tp.sum(), total).float()
        score = self._safe_div(tp, denom)
        if self.average in (None, 'none'):
            return score.float()
        if self.average == 'weighted':
            w = tp + fn
        else:   # macro: weight one for every class (absent ones score 0), zero for the ignored class
            w = torch.ones_like(score)
            if ign is not None:
                w[ign] = 0.0
        return self._safe_div(w * score, w.sum()).sum().float()


# ---------------------------------------------------------------------------------------------- the manager
class MetricWithUtils(nn.Module):
    """A metric, its output→argument mapping, its log name and the dataloader it listens to
    (metrics_manager.py:12-72)."""

    def __init__(self, metric, mapping, log_name, dataloader_idx):
        super().__init__()
        self.metric, self.mapping, self.log_name, self.dataloader_idx = metric, dict(mapping), log_name, dataloader_idx

    def map_arguments(self, task_output):
        picked = {}
        for dst, src in self.mapping.items():
            if src not in task_output:
                raise ValueError(f'Cannot find {src} for your mapping {dst} : {src}. You should either add {src} '
                                 f'output to your model or remove the mapping from configuration')
            picked[dst] = task_output[src]
        return picked

    def update(self, dataloader_idx=0, **kwargs):
        if dataloader_idx == self.dataloader_idx:
            self.metric.update(**self.map_arguments(kwargs))

    def compute(self):
        return self.metric.compute()

    def reset(self):
        self.metric.reset()


def _field(params, key, default):
    if isinstance(params, dict):
        value = params.get(key, default)
    else:
        value = getattr(params, key, default)
    return default if value is None and key in ('phases', 'val_dataloader_idxs', 'test_dataloader_idxs', 'params') \
        else value


class MetricsManager(nn.Module):
    """metrics_manager.py:75-206.  `params`: the `metrics:` list of a config (dicts or objects with name / mapping /
    params / phases / val_dataloader_idxs / test_dataloader_idxs / tag)."""

    def __init__(self, params):
        super().__init__()
        self.phase2metrics = nn.ModuleDict()
        for phase in Phase:
            self.phase2metrics[phase.name] = self._phase_metrics(params or [], phase)

    @staticmethod
    def _phase_metrics(params, phase):
        seen, metrics = set(), []
        for mp in params:
            phases = [as_phase(p) for p in _field(mp, 'phases', list(Phase))]
            if phase not in phases:
                continue
            name, tag = _field(mp, 'name', None), _field(mp, 'tag', None)
            base = name if tag is None else tag
            if phase == Phase.VALID:
                idxs = list(_field(mp, 'val_dataloader_idxs', [0]))
            elif phase == Phase.TEST:
                idxs = list(_field(mp, 'test_dataloader_idxs', [0]))
            else:
                idxs = [0]
            if phase in (Phase.VALID, Phase.TEST) and len(idxs) > 1:
                log_names = [f'{base}_dataloader_{i}' for i in idxs]
            else:
                log_names = [base]
            for log_name in log_names:
                if log_name in seen:
                    raise ValueError(f'Got two metrics with identical names: {log_name}. '
                                     f'Please, set different prefixes for identical metrics in the config file.')
                seen.add(log_name)
            for idx, log_name in zip(idxs, log_names):
                metric = METRICS.get(name)(**dict(_field(mp, 'params', {}) or {}))
                metrics.append(MetricWithUtils(metric, _field(mp, 'mapping', {}), log_name, idx))
        return nn.ModuleList(metrics)

    def update(self, phase, dataloader_idx=0, **kwargs):
        for m in self.phase2metrics[as_phase(phase).name]:
            m.update(dataloader_idx, **kwargs)

    @staticmethod
    def is_number(num):
        if isinstance(num, np.ndarray):
            return num.ndim == 0 and np.issubdtype(num.dtype, np.number)
        if isinstance(num, torch.Tensor):
            return num.dim() == 0
        return isinstance(num, numbers.Number)

    def on_epoch_end(self, phase):
        phase = as_phase(phase)
        log = {}
        for m in self.phase2metrics[phase.name]:
            value = m.compute()
            if isinstance(value, dict):
                numeric = {f'{phase.value}/{m.log_name}_{k}': v for k, v in value.items() if self.is_number(v)}
                if not numeric:
                    raise ValueError(f'Metric manager on_epoch_end method. Metric {m.log_name}'
                                     f'return dict with has no numeric values.')
                log.update(numeric)
            elif self.is_number(value):
                log[f'{phase.value}/{m.log_name}'] = value
            else:
                raise ValueError(f'Metric manager on_epoch_end method. Metric {m.log_name} '
                                 f'return no numeric value.')
            m.reset()
        return log
