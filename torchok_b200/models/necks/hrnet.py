"""HRNet necks (torchok/models/necks/segmentation/hrnet.py:16-42, torchok/models/necks/classification/hrnet.py:12-85).

Segmentation neck: the three low-resolution branches are bilinearly resized to branch 0's size and concatenated — one
tok_bilinear_fwd per branch writing straight into its channel segment of the concat buffer — then a fused 1x1
ConvBnReLU.  Classification neck: Bottleneck "incre" modules, stride-2 3x3 ConvBnAct downsamples and a 1x1 to 2048
channels, INCLUDING the reference's overwrite quirk: `y = incre[i+1](x[i+1])` replaces the downsampled sum,
so the result is `final_layer(incre[3](x[3]))` while the discarded modules are still constructed (state-dict parity).
"""
import torch.nn as nn

from ... import kernels as K
from ...constructor import NECKS
from ..backbones.resnet import Bottleneck
from ..base import BaseModel
from ..modules.bricks import ConvBnAct


@NECKS.register_class
class HRNetSegmentationNeck(BaseModel):
    def __init__(self, in_channels):
        out_channels = sum(in_channels)
        super().__init__(in_channels, out_channels)
        self.convbnact = ConvBnAct(out_channels, out_channels, kernel_size=1, padding=0, stride=1, act_layer=nn.ReLU)
        self.convbnact.conv.set_input_layout(list(in_channels))

    def forward(self, features):
        input_image, x0, x1, x2, x3 = features
        feats = K.bilinear_cat([x0, x1, x2, x3], (x0.size(2), x0.size(3)))
        feats = self.convbnact(feats)
        return [input_image, feats]


@NECKS.register_class
class HRNetClassificationNeck(BaseModel):
    def __init__(self, in_channels):
        super().__init__(in_channels, 2048)
        self.head_channels = [32, 64, 128, 256]
        self.incre_modules = nn.ModuleList([self._make_layer(Bottleneck, c, self.head_channels[i], 1, stride=1)
                                            for i, c in enumerate(in_channels)])
        self.downsamp_modules = nn.ModuleList([
            ConvBnAct(self.head_channels[i] * Bottleneck.expansion, self.head_channels[i + 1] * Bottleneck.expansion,
                      kernel_size=3, padding=1, stride=2) for i in range(len(in_channels) - 1)])
        self.final_layer = ConvBnAct(self.head_channels[3] * Bottleneck.expansion, self.out_channels, kernel_size=1,
                                     padding=0, stride=1)

    @staticmethod
    def _make_layer(block, inplanes, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = ConvBnAct(inplanes, planes * block.expansion, kernel_size=1, padding=0, stride=stride,
                                   bias=False, act_layer=None)
        layers = [_NeckBottleneck(inplanes, planes, stride, downsample)]
        inplanes = planes * block.expansion
        layers += [_NeckBottleneck(inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        # necks/classification/hrnet.py:77-85 — y is OVERWRITTEN by every incre module, so only the last branch reaches
        # the output; the discarded modules still run (in training mode they update their BatchNorm running statistics
        # in the reference, and so they do here).
        y = self.incre_modules[0](x[0])
        for i in range(len(self.downsamp_modules)):
            y = self.downsamp_modules[i](y)
            if i + 1 < len(x):
                y = self.incre_modules[i + 1](x[i + 1])
        return self.final_layer(y)


class _NeckBottleneck(Bottleneck):
    """timm Bottleneck whose shortcut is a ConvBnAct module (submodules .conv / .bn) instead of nn.Sequential."""

    def tok_downsample(self):
        ds = self.downsample
        return None if ds is None else (ds.conv, ds.bn)
