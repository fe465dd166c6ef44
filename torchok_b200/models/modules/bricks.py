"""ConvBnAct: convolution + BatchNorm + activation as ONE fused unit
(same constructor as torchok/models/modules/bricks/convbnact.py:8-53; submodule names conv / bn / act preserved)."""
import torch.nn as nn

from .layers import BatchNorm2d, Conv2d, ConvFn, ReLU, conv_bn_act


class ConvBnAct(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, bias=False, use_batchnorm=True,
                 groups=1, act_layer=nn.ReLU):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                           groups=groups, bias=bias)
        self.bn = BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        if act_layer is not None and not (isinstance(act_layer, type) and issubclass(act_layer, nn.ReLU)):
            raise NotImplementedError('ConvBnAct: only ReLU (or no activation) is fused on the hot path')
        self.act = ReLU(inplace=True) if act_layer is not None else nn.Identity()
        if use_batchnorm and bias:
            raise NotImplementedError('ConvBnAct: a conv bias in front of BatchNorm is redundant and not supported')

    def forward(self, x):
        relu = isinstance(self.act, nn.ReLU)
        if isinstance(self.bn, nn.Identity):
            return ConvFn.apply(x, self.conv, relu, self.conv.weight, self.conv.bias)
        return conv_bn_act(x, self.conv, self.bn, relu=relu)
