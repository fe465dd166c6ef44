"""Plug-in registries of the package.

The drop-in boundary is the SET OF NAMES the reference exposes from `torchok/constructor/__init__.py:4-17`:
configs and user code say `BACKBONES.get('resnet50')`, `TASKS.get(cfg.task.name)`, ...  Each name is a `Registry`
labelled with its lower-case spelling; they are created from one table so that the contract is visible in one place
(tests/test_host_logic.py checks all fourteen).
"""
from .registry import Registry

REGISTRY_NAMES = (
    # model parts, in the order a task assembles them
    'backbones', 'necks', 'detection_necks', 'poolings', 'heads',
    # what a config wires around the model
    'tasks', 'losses', 'metrics', 'optimizers', 'schedulers', 'callbacks',
    # data side (kept for config compatibility; the data pipeline itself is out of scope)
    'datasets', 'transforms', 'samplers',
)

for _name in REGISTRY_NAMES:
    globals()[_name.upper()] = Registry(_name)
del _name

__all__ = ['Registry', 'REGISTRY_NAMES'] + [n.upper() for n in REGISTRY_NAMES]
