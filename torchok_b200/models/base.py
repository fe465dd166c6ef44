"""BaseModel / BaseBackbone / BackboneWrapper: the interface contract every model part obeys.

Mirrors torchok/models/base.py:8-63 and torchok/models/backbones/base_backbone.py:11-64: `in_channels` /
`out_channels` properties that raise ValueError when unset, `no_weight_decay()`, `init_weights()`, and for backbones
`out_encoder_channels`, `forward_features(x) -> [x, f1..fn]`, abstract `get_stages(stage)`.  The reference collects
the intermediate features with timm FeatureHooks; here a backbone returns them from `_forward_collect` directly.
"""
from abc import ABC, abstractmethod

import torch.nn as nn


class BaseModel(nn.Module, ABC):
    def __init__(self, in_channels=None, out_channels=None):
        super().__init__()
        self._in_channels = in_channels
        self._out_channels = out_channels

    @abstractmethod
    def forward(self, *args, **kwargs):
        ...

    def no_weight_decay(self):
        return []

    @property
    def in_channels(self):
        if self._in_channels is None:
            raise ValueError('TorchOk Models must have self._in_channels attribute.')
        return self._in_channels

    @property
    def out_channels(self):
        if self._out_channels is None:
            raise ValueError('TorchOk Models must have self._out_channels attribute.')
        return self._out_channels

    def init_weights(self):
        # torchok/models/base.py:50-63
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode='fan_in', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)


class BaseBackbone(BaseModel, ABC):
    _out_encoder_channels = None

    def create_hooks(self):
        """Record stage names / channel counts from `self.feature_info` (base_backbone.py:14-24)."""
        self.stage_names = [h['module'] for h in self.feature_info]
        self._out_encoder_channels = [h['num_chs'] for h in self.feature_info]

    @abstractmethod
    def _forward_collect(self, x):
        """Run the backbone and return the list of stage outputs named in `feature_info` (last = forward output)."""

    def forward(self, x):
        return self._forward_collect(x)[-1]

    def forward_features(self, x):
        return [x] + list(self._forward_collect(x))

    @property
    def out_encoder_channels(self):
        if self._out_encoder_channels is None:
            raise ValueError('TorchOk Backbones must have self._out_feature_channels attribute.')
        return tuple(self._out_encoder_channels)

    @abstractmethod
    def get_stages(self, stage):
        ...


class BackboneWrapper(nn.Module):
    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, x):
        return self.backbone.forward_features(x)

    @property
    def out_encoder_channels(self):
        return self.backbone.out_encoder_channels
