"""BaseTask: the object the step loop drives.

The reference's BaseTask is a pytorch_lightning.LightningModule (torchok/tasks/base.py:17-204); Lightning's fit loop
is replaced here by torchok_b200.engine.StreamLoop, so this class is a plain nn.Module that keeps the reference's
method surface: `forward`, `forward_with_gt(batch) -> dict`, `training_step` / `validation_step` (same returned
dicts), `configure_optimizers`, `as_module`, the `inputs` example buffers and `self.losses` (JointLoss built from
`hparams.joint_loss`, constructor.py:367-382).
"""
from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from ..constructor import LOSSES
from ..constructor.config import Config
from ..losses.base import JointLoss


def configure_losses(hparams):
    """Constructor.configure_losses (torchok/constructor/constructor.py:367-382)."""
    modules, mappings, tags, weights = [], [], [], []
    for cfg in hparams.joint_loss.losses:
        modules.append(LOSSES.get(cfg.name)(**(cfg.get('params') or {})))
        mappings.append(cfg.mapping)
        tags.append(cfg.get('tag'))
        weights.append(cfg.get('weight'))
    return JointLoss(modules, mappings, tags, weights, hparams.joint_loss.get('normalize_weights', True))


class BaseTask(nn.Module, ABC):
    def __init__(self, hparams, inputs=None, **kwargs):
        super().__init__()
        self._hparams = hparams if isinstance(hparams, Config) else Config.wrap(hparams or {})
        self.input_tensor_names = []
        self.losses = configure_losses(self._hparams) if self._hparams.get('joint_loss') is not None else None
        self.example_input_array = []
        if inputs is not None:
            for i, spec in enumerate(inputs):
                name = f'input_tensors_{i}'
                self.input_tensor_names.append(name)
                t = torch.rand(1, *spec['shape']).type(getattr(torch, spec['dtype']))
                self.example_input_array.append(t)
                self.register_buffer(name, t)

    @property
    def hparams(self):
        return self._hparams

    @abstractmethod
    def forward(self, *args, **kwargs):
        ...

    @abstractmethod
    def forward_with_gt(self, batch):
        ...

    @abstractmethod
    def as_module(self):
        ...

    def training_step(self, batch, batch_idx=0):
        """torchok/tasks/base.py:125-133 minus logging/metrics (done by the loop on device-side accumulators)."""
        output = self.forward_with_gt(batch)
        total_loss, tagged = self.losses(**output)
        # what `metrics_manager.update(Phase.TRAIN, **output)` reads (base.py:130); under graph replay these are the
        # graph's static output tensors, refreshed by every replay
        self.last_output = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in output.items()}
        out = {'loss': total_loss}
        out.update(tagged)
        return out

    def validation_step(self, batch, batch_idx=0, dataloader_idx=0):
        output = self.forward_with_gt(batch)
        if self._hparams.get('task') is None or self._hparams.task.get('compute_loss_on_valid', True):
            total_loss, tagged = self.losses(**output)
            out = {'loss': total_loss}
            out.update(tagged)
            return out, output
        return {}, output

    def predict_step(self, batch, batch_idx=0, dataloader_idx=0):
        return self.forward_with_gt(batch)
