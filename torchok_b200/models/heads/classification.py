"""LinearHead / ClassificationHead (torchok/models/heads/representation/linear_head.py:10-36,
torchok/models/heads/classification/classification_head.py:9-40): dropout -> FC -> optional L2 normalise; the
classification variant drops the channel axis when num_classes == 1.  The FC runs on tok_linear_*."""
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels as K
from ...constructor import HEADS
from ..base import BaseModel


@HEADS.register_class
class LinearHead(BaseModel):
    def __init__(self, in_channels, out_channels, drop_rate=0.0, bias=True, normalize=False):
        super().__init__(in_channels, out_channels)
        self.drop_rate = drop_rate
        self.normalize = normalize
        self.fc = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, targets=None):
        if self.drop_rate > 0.:
            x = F.dropout(x, p=self.drop_rate, training=self.training)
        x = K.linear(x, self.fc.weight, self.fc.bias)
        if self.normalize:
            x = K.l2_normalize(x)
        return x


@HEADS.register_class
class ClassificationHead(LinearHead):
    def __init__(self, in_channels, num_classes, drop_rate=0.0, bias=True):
        super().__init__(in_channels, out_channels=num_classes, drop_rate=drop_rate, bias=bias)

    def forward(self, x, target=None):
        x = super().forward(x, target)
        if self.out_channels == 1:
            x = x[..., 0]
        return x
