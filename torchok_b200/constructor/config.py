"""YAML experiment configs without hydra/omegaconf.

Accepts the reference's example files unchanged: YAML anchors/aliases are native to PyYAML; the
interpolations the examples use — `${oc.env:VAR}`, `${now:%fmt}`, `${a.b.c}` — are resolved here; `key.sub=value`
overrides follow hydra's CLI form (torchok/__main__.py:13-31).  `Config` gives attribute + item access like DictConfig
and `.get()`; missing optional blocks resolve to None like the dataclass defaults in
torchok/constructor/config_structure.py:185-196.
"""
import os
import re
from datetime import datetime

import yaml

_TOP_DEFAULTS = dict(task=None, data=None, optimization=None, joint_loss=None, trainer=None, logger=None,
                     callbacks=None, metrics=None, resume_path=None, seed_params=None, hydra=None)
_TASK_DEFAULTS = dict(compute_loss_on_valid=True, load_checkpoint=None, params=None)


class Config(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return Config({k: Config.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return [Config.wrap(v) for v in obj]
        return obj

    def to_dict(self):
        def un(o):
            if isinstance(o, dict):
                return {k: un(v) for k, v in o.items()}
            if isinstance(o, list):
                return [un(v) for v in o]
            return o
        return un(self)


_PAT = re.compile(r'\$\{([^${}]+)\}')


def _lookup(root, dotted):
    node = root
    for part in dotted.split('.'):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def _resolve_str(text, root, now, depth=0):
    if depth > 16:
        raise ValueError(f'interpolation too deep in {text!r}')

    def one(expr):
        expr = expr.strip()
        if expr.startswith('oc.env:'):
            name, _, default = expr[len('oc.env:'):].partition(',')
            if name in os.environ:
                return os.environ[name]
            if default:
                return default.strip()
            raise KeyError(f'environment variable {name} is not set')
        if expr.startswith('now:'):
            return now.strftime(expr[len('now:'):])
        return _lookup(root, expr)

    m = _PAT.fullmatch(text)
    if m:  # whole-value interpolation keeps the type
        val = one(m.group(1))
        return _resolve_str(val, root, now, depth + 1) if isinstance(val, str) else val
    out = _PAT.sub(lambda mm: str(one(mm.group(1))), text)
    return _resolve_str(out, root, now, depth + 1) if _PAT.search(out) else out


def _resolve(node, root, now):
    if isinstance(node, dict):
        return {k: _resolve(v, root, now) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, now) for v in node]
    if isinstance(node, str) and '${' in node:
        return _resolve_str(node, root, now)
    return node


def apply_overrides(raw, overrides):
    for item in overrides or ():
        key, _, value = item.partition('=')
        key = key.lstrip('+')
        node = raw
        parts = key.split('.')
        for part in parts[:-1]:
            if isinstance(node, list):
                node = node[int(part)]
            else:
                node = node.setdefault(part, {})
        leaf = parts[-1]
        value = yaml.safe_load(value)
        if isinstance(node, list):
            node[int(leaf)] = value
        else:
            node[leaf] = value
    return raw


# Field names of the reference's structured schema (torchok/constructor/config_structure.py:14-196).  hydra merges every
# config into `OmegaConf.structured(ConfigParams)` (torchok/__main__.py:27-31), so a key the schema does not know is an
# error there (`ConfigKeyError`, a KeyError: "Key 'bag' not in 'JointLossParams'"); `validate_schema` reproduces that.
# `Dict` / `Any` typed fields (params, mapping, dataloader, paramwise_cfg, ...) are free-form.  The table is compared
# with the reference's dataclasses in tests/test_front_door_goldens.py.
SCHEMA = {
    'ConfigParams': {'task': 'TaskParams', 'data': 'data', 'trainer': 'TrainerParams', 'optimization': ['OptimizationParams'],
                     'joint_loss': 'JointLossParams', 'logger': 'LoggerParams', 'metrics': ['MetricParams'],
                     'callbacks': ['CallbacksParams'], 'resume_path': None, 'seed_params': 'SeedParams'},
    'TaskParams': {'name': None, 'compute_loss_on_valid': None, 'params': None, 'load_checkpoint': 'LoadCheckpointParams'},
    'LoadCheckpointParams': {'base_ckpt_path': None, 'overridden_name2ckpt_path': None, 'exclude_keys': None, 'strict': None},
    'JointLossParams': {'losses': ['LossParams'], 'normalize_weights': None},
    'LossParams': {'name': None, 'mapping': None, 'params': None, 'tag': None, 'weight': None},
    'OptimizationParams': {'optimizer': 'OptmizerParams', 'scheduler': 'SchedulerParams'},
    'OptmizerParams': {'name': None, 'params': None, 'paramwise_cfg': None},
    'SchedulerParams': {'name': None, 'params': None, 'pl_params': 'SchedulerPLParams'},
    'SchedulerPLParams': {'interval': None, 'frequency': None, 'monitor': None, 'strict': None, 'name': None},
    'DataParams': {'dataset': 'DatasetParams', 'dataloader': None, 'sampler': 'SamplerParams'},
    'DatasetParams': {'name': None, 'params': None, 'transform': ['AugmentationParams'], 'augment': ['AugmentationParams']},
    'AugmentationParams': {'name': None, 'params': None},
    'SamplerParams': {'name': None, 'params': None},
    'MetricParams': {'name': None, 'mapping': None, 'params': None, 'phases': None, 'val_dataloader_idxs': None,
                     'test_dataloader_idxs': None, 'tag': None},
    'CallbacksParams': {'name': None, 'params': None},
    'SeedParams': {'seed': None, 'workers': None},
    'LoggerParams': {'name': None, 'log_dir': None, 'experiment_name': None, 'timestamp': None, 'params': None},
    'TrainerParams': {k: None for k in (
        'accelerator', 'strategy', 'devices', 'num_nodes', 'precision', 'fast_dev_run', 'max_epochs', 'min_epochs',
        'max_steps', 'min_steps', 'max_time', 'limit_train_batches', 'limit_val_batches', 'limit_test_batches',
        'limit_predict_batches', 'overfit_batches', 'val_check_interval', 'check_val_every_n_epoch',
        'num_sanity_val_steps', 'log_every_n_steps', 'enable_checkpointing', 'enable_progress_bar',
        'enable_model_summary', 'accumulate_grad_batches', 'gradient_clip_val', 'gradient_clip_algorithm',
        'deterministic', 'benchmark', 'inference_mode', 'use_distributed_sampler', 'profiler', 'detect_anomaly',
        'barebones', 'sync_batchnorm', 'reload_dataloaders_every_n_epochs')},
}
PHASES = ('TRAIN', 'VALID', 'TEST', 'PREDICT')
_HYDRA_KEYS = ('hydra', 'defaults', 'mode')      # consumed by hydra / __main__ before the schema merge


def validate_schema(raw):
    """Raise KeyError for keys the reference's structured config would reject."""
    def check(node, cls):
        if node is None:
            return
        if cls == 'data':
            for phase, entries in node.items():
                if phase not in PHASES:
                    raise KeyError(f"Invalid value '{phase}', expected one of [{', '.join(PHASES)}]")
                for entry in entries or []:
                    check(entry, 'DataParams')
            return
        if not isinstance(node, dict):
            raise TypeError(f'{cls}: expected a mapping, got {type(node).__name__}')
        fields = SCHEMA[cls]
        for key, value in node.items():
            if cls == 'ConfigParams' and key in _HYDRA_KEYS:
                continue
            if key not in fields:
                raise KeyError(f"Key '{key}' not in '{cls}'")
            sub = fields[key]
            if isinstance(sub, list):
                for item in value or []:
                    check(item, sub[0])
            elif sub is not None:
                check(value, sub)
    check(raw, 'ConfigParams')


def load_config(path_or_dict, overrides=None, strict=None):
    """`strict` (default: True for files, False for dicts built in code) applies `validate_schema`."""
    if isinstance(path_or_dict, (str, os.PathLike)):
        with open(path_or_dict) as f:
            raw = yaml.safe_load(f)
    else:
        raw = Config.wrap(path_or_dict).to_dict()
    raw = apply_overrides(raw, overrides)
    raw = _resolve(raw, raw, datetime.now())
    if strict if strict is not None else isinstance(path_or_dict, (str, os.PathLike)):
        validate_schema(raw)
    for k, v in _TOP_DEFAULTS.items():
        raw.setdefault(k, v)
    if isinstance(raw.get('task'), dict):
        for k, v in _TASK_DEFAULTS.items():
            raw['task'].setdefault(k, v)
    jl = raw.get('joint_loss')
    if isinstance(jl, dict):
        jl.setdefault('normalize_weights', True)
        for loss in jl.get('losses', []):
            loss.setdefault('params', {})
            loss.setdefault('tag', None)
            loss.setdefault('weight', None)
            loss.setdefault('mapping', {})
    return Config.wrap(raw)
