"""Pooling / PoolingLinear (torchok/models/poolings/classification/pooling.py:7-12, linear.py:8-25).

`Pooling` restates timm's SelectAdaptivePool2d(output_size=1, flatten=True): 'avg' | 'max' | 'avgmax' (= 0.5*(avg+max))
| 'catavgmax' (concat, hence 2*in_channels), computed by tok_gap_fwd.
"""
import torch
import torch.nn as nn

from ... import kernels as K
from ...constructor import POOLINGS
from ..base import BaseModel


@POOLINGS.register_class
class Pooling(BaseModel):
    def __init__(self, in_channels, pooling_type='avg', output_size=1):
        super().__init__(in_channels, in_channels if pooling_type != 'catavgmax' else 2 * in_channels)
        if output_size != 1:
            raise NotImplementedError('Pooling: only global pooling (output_size=1) is on the hot path')
        if pooling_type not in ('avg', 'max', 'avgmax', 'catavgmax'):
            raise ValueError(f'Invalid pool type: {pooling_type}')
        self.pool_type = pooling_type

    def forward(self, x):
        if self.pool_type == 'catavgmax':
            return torch.cat([K.GlobalPoolFn.apply(x, 'avg'), K.GlobalPoolFn.apply(x, 'max')], dim=1)
        return K.GlobalPoolFn.apply(x, self.pool_type)


@POOLINGS.register_class
class PoolingLinear(Pooling):
    def __init__(self, in_channels, out_channels, pooling_type='avg', output_size=1, bias=True):
        super().__init__(in_channels, pooling_type, output_size=output_size)
        self.fc = nn.Linear(self._out_channels, out_channels, bias=bias)
        self._out_channels = out_channels
        self.init_weights()

    def forward(self, x):
        x = super().forward(x)
        return K.linear(x, self.fc.weight, self.fc.bias)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
