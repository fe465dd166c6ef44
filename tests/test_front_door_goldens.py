"""Front-door host logic replayed against vectors produced by the REFERENCE's own files
(tests/golden/make_front_door_goldens.py executes torchok/constructor/load.py, torchok/metrics/metrics_manager.py and
torchok/callbacks/freeze_unfreeze.py by path under stubbed framework imports; fixture
tests/golden/front_door_goldens.pt)."""
import importlib.util
import os

import pytest
import torch
import torch.nn as nn

import torchok_b200 as tb
from torchok_b200.callbacks import FreezeUnfreeze
from torchok_b200.constructor.load import load_checkpoint
from torchok_b200.metrics.metrics_manager import MetricsManager

HERE = os.path.join(os.path.dirname(__file__), 'golden')
G = torch.load(os.path.join(HERE, 'front_door_goldens.pt'), weights_only=False)
_spec = importlib.util.spec_from_file_location('_mk_front_door', os.path.join(HERE, 'make_front_door_goldens.py'))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)            # helpers only: nothing under /root/reference is touched at import time

for _cls in (mk.MockSum, mk.MockDict):
    if _cls.__name__ not in tb.METRICS:
        tb.METRICS.register_class(_cls)


@pytest.mark.parametrize('case', G['MetricsManager'], ids=lambda c: '+'.join(p['name'] for p in c['spec']))
def test_metrics_manager_matches_reference(case):
    mgr = MetricsManager(case['spec'])
    for phase, idx in case['updates']:
        mgr.update(phase, idx, emb=torch.zeros(1), y=torch.zeros(1))
    for phase, want in case['logs'].items():
        got = {k: int(v) for k, v in mgr.on_epoch_end(phase).items()}
        assert got == want, phase
    assert {k: int(v) for k, v in mgr.on_epoch_end('VALID').items()} == case['valid_after_reset']


@pytest.mark.parametrize('case', G['load_checkpoint'], ids=lambda c: f"{sorted(c['overrides'])}-{c['exclude_keys']}")
def test_load_checkpoint_matches_reference(case, tmp_path):
    task = mk.tree_from_spec(case['spec'])
    task.load_state_dict(case['initial'])
    torch.save({'state_dict': case['base']}, tmp_path / 'base.ckpt')
    paths = {}
    for name, state in case['overrides'].items():
        torch.save(state, tmp_path / f'{name}.pth')
        paths[name] = str(tmp_path / f'{name}.pth')
    load_checkpoint(task, str(tmp_path / 'base.ckpt'), paths or None, case['exclude_keys'] or None)
    got = task.state_dict()
    assert set(got) == set(case['loaded'])
    for k, v in case['loaded'].items():
        assert torch.equal(got[k], v), k


@pytest.mark.parametrize('case', G['FreezeUnfreeze'], ids=lambda c: str([r['module_name'] for r in c['rules']]))
def test_freeze_unfreeze_matches_reference(case):
    """Epoch-by-epoch `requires_grad` of every parameter and `track_running_stats` of every BatchNorm: the
    reference's freeze_before_training, then finetune_function(epoch) for epochs 0..3."""
    task = mk.tree_from_spec(case['spec'])
    cb = FreezeUnfreeze(case['rules'], top_down_freeze_order=case['top_down'])

    class R:                                    # the two things the hooks touch on a runner
        current_epoch = 0
        changes = 0

        def frozen_set_changed(self):
            self.changes += 1
    r = R()
    r.task = task
    cb.setup(r)
    assert mk.flags(task) == case['history'][0]
    for epoch in range(4):
        r.current_epoch = epoch
        cb.on_train_epoch_start(r)
        assert mk.flags(task) == case['history'][epoch + 1], epoch
    distinct = sum(a != b for a, b in zip(case['history'], case['history'][1:]))
    assert r.changes == 1 + distinct            # the step graph is recaptured exactly when the flags change
    assert case['n_groups'] == 1                # thawing adds no optimizer group in the reference either (all present)


def test_registry_matches_reference():
    """The same scenario through torchok/constructor/registry.py (executed by path; timm's `_natural_key` stubbed with
    its one-line definition) and through torchok_b200.constructor.Registry: listings under every filter form, `in`,
    repr, module bookkeeping, `__all__` export, and the type + text of all four errors."""
    from torchok_b200.constructor.registry import Registry
    got = mk.registry_scenario(Registry)
    want = G['Registry']
    assert got == want


def test_config_schema_matches_reference_dataclasses(tmp_path):
    """constructor/config.py::SCHEMA (what `validate_schema` enforces on YAML files) against the field names of the
    reference's dataclasses (torchok/constructor/config_structure.py executed as is), and the error behaviour hydra's
    structured merge has for unknown keys (tests/base_tests/constructor/test_config_structure_load.py:58-59: the
    `bag` key under joint_loss raises KeyError)."""
    from torchok_b200.constructor.config import PHASES, SCHEMA, load_config, validate_schema
    ref = G['config_structure']
    assert set(SCHEMA) == set(ref)
    for cls, fields in ref.items():
        assert list(SCHEMA[cls]) == fields or set(SCHEMA[cls]) == set(fields), cls
    assert list(PHASES) == G['phases']
    good = {'task': {'name': 'ClassificationTask', 'params': {'anything': {'goes': 1}}},
            'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction'}}]},
            'data': {'TRAIN': [{'dataset': {'name': 'X', 'params': {}, 'transform': []}, 'dataloader': {'whatever': 1}}]},
            'trainer': {'max_epochs': 1}, 'hydra': {'run': {'dir': 'x'}}}
    validate_schema(good)
    for path, key, cls in [(('joint_loss',), 'bag', 'JointLossParams'), ((), 'log_dir', 'ConfigParams'),
                           (('trainer',), 'gpus', 'TrainerParams'), (('task',), 'input_size', 'TaskParams'),
                           (('joint_loss', 'losses', 0), 'wieght', 'LossParams')]:
        import copy
        bad = copy.deepcopy(good)
        node = bad
        for part in path:
            node = node[part]
        node[key] = 1
        with pytest.raises(KeyError, match=f"Key '{key}' not in '{cls}'"):
            validate_schema(bad)
    with pytest.raises(KeyError, match='expected one of'):
        validate_schema({'data': {'TRIAN': []}})
    # files are validated, dicts built in code are not (unless asked to)
    (tmp_path / 'bad.yaml').write_text('task: {name: X}\njoint_loss: {bag: big_bag, losses: []}\n')
    with pytest.raises(KeyError, match="Key 'bag' not in 'JointLossParams'"):
        load_config(str(tmp_path / 'bad.yaml'))
    assert load_config({'task': {'name': 'X'}, 'joint_loss': {'bag': 1, 'losses': []}}).joint_loss.bag == 1
    with pytest.raises(KeyError):
        load_config({'task': {'name': 'X'}, 'bag': 1}, strict=True)
