"""ResNet backbones on the sm_100a kernels.

Mirror of torchok/models/backbones/resnet.py:408-563 (class ResNet, make_blocks :363-405) restricted to the plain
family the hot path names: 7x7 stem, BasicBlock / Bottleneck (timm semantics, SURVEY Appendix A.1: stride on the 3x3,
conv-bn-act ordering, zero_init_last), 1x1 conv+BN downsample, output_stride 32.  Parameter names equal timm's /
torchvision's (`conv1.weight`, `layer1.0.bn2.running_var`, `layer2.0.downsample.1.weight` ...), so reference
checkpoints load unchanged.  Variants that need kernels outside the scope table (grouped / SE / ECA / anti-aliased /
deep-stem / avg-down) raise NotImplementedError at construction.
"""
import math

import torch
import torch.nn as nn

from ...constructor import BACKBONES
from ..base import BaseBackbone
from ..modules.layers import BatchNorm2d, Conv2d, MaxPool2d, ReLU, residual_block, stem


class _Block(nn.Module):
    expansion = 1

    def tok_downsample(self):
        ds = self.downsample
        return None if ds is None else (ds[0], ds[1])

    def forward(self, x):
        return residual_block(x, self)


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.act1 = ReLU(inplace=True)
        self.conv2 = Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.act2 = ReLU(inplace=True)
        self.downsample = downsample

    def zero_init_last(self):
        nn.init.zeros_(self.bn2.weight)

    def tok_units(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2)]


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, base_width=64):
        super().__init__()
        width = int(math.floor(planes * (base_width / 64)))
        outplanes = planes * self.expansion
        self.conv1 = Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = BatchNorm2d(width)
        self.act1 = ReLU(inplace=True)
        self.conv2 = Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = BatchNorm2d(width)
        self.act2 = ReLU(inplace=True)
        self.conv3 = Conv2d(width, outplanes, 1, bias=False)
        self.bn3 = BatchNorm2d(outplanes)
        self.act3 = ReLU(inplace=True)
        self.downsample = downsample

    def zero_init_last(self):
        nn.init.zeros_(self.bn3.weight)

    def tok_units(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)]


def _make_stage(block, inplanes, planes, blocks, stride, **block_kwargs):
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(Conv2d(inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                                   BatchNorm2d(planes * block.expansion))
    layers = [block(inplanes, planes, stride, downsample, **block_kwargs)]
    inplanes = planes * block.expansion
    layers += [block(inplanes, planes, **block_kwargs) for _ in range(1, blocks)]
    return nn.Sequential(*layers), inplanes


class ResNet(BaseBackbone):
    def __init__(self, block, layers, in_channels=3, output_stride=32, cardinality=1, base_width=64, stem_width=64,
                 stem_type='', replace_stem_pool=False, block_reduce_first=1, down_kernel_size=1, avg_down=False,
                 act_layer=None, norm_layer=None, aa_layer=None, drop_path_rate=0., drop_block_rate=0.,
                 zero_init_last=True, block_args=None):
        super().__init__(in_channels=in_channels)
        if output_stride not in (8, 16, 32):
            raise ValueError('`output_stride` must be in (8, 16, 32)')
        unsupported = dict(output_stride=output_stride != 32, cardinality=cardinality != 1, stem_type=bool(stem_type),
                           replace_stem_pool=replace_stem_pool, block_reduce_first=block_reduce_first != 1,
                           down_kernel_size=down_kernel_size != 1, avg_down=avg_down, act_layer=act_layer is not None,
                           norm_layer=norm_layer is not None, aa_layer=aa_layer is not None,
                           drop_path_rate=drop_path_rate != 0, drop_block_rate=drop_block_rate != 0,
                           block_args=bool(block_args))
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f'torchok_b200 ResNet: option(s) {bad} need kernels outside the hot-path scope '
                                      '(plain resnet18/34/50/101/152 are supported)')
        if in_channels > 4:
            raise NotImplementedError('the 7x7 stem kernel takes at most 4 input channels')
        inplanes = 64
        self.conv1 = Conv2d(in_channels, inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm2d(inplanes)
        self.act1 = ReLU(inplace=True)
        self.feature_info = [dict(num_chs=inplanes, reduction=2, module='act1')]
        self.maxpool = MaxPool2d(kernel_size=3, stride=2, padding=1)
        kw = dict(base_width=base_width) if block is Bottleneck else {}
        net_stride = 4
        for i, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
            stride = 1 if i == 0 else 2
            net_stride *= stride
            stage, inplanes = _make_stage(block, inplanes, planes, n, stride, **kw)
            self.add_module(f'layer{i + 1}', stage)
            self.feature_info.append(dict(num_chs=inplanes, reduction=net_stride, module=f'layer{i + 1}'))
        self._out_channels = 512 * block.expansion
        self.create_hooks()
        self.init_weights(zero_init_last=zero_init_last)

    @torch.no_grad()
    def init_weights(self, zero_init_last=True):
        # resnet.py:529-539
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_last:
            for m in self.modules():
                if hasattr(m, 'zero_init_last'):
                    m.zero_init_last()

    def forward(self, x):
        # only the last stage is needed: the stem does not materialise the act1 feature
        return self._forward_collect(x, need_act=False)[-1]

    def _forward_collect(self, x, need_act=True):
        act1, x = stem(x, self.conv1, self.bn1, (3, 2, 1), need_act=need_act)
        feats = [act1]
        for name in ('layer1', 'layer2', 'layer3', 'layer4'):
            x = getattr(self, name)(x)
            feats.append(x)
        return feats

    def get_stages(self, stage):
        output = [self.conv1, self.bn1, self.act1, self.maxpool]
        layers = [self.layer1, self.layer2, self.layer3, self.layer4]
        return nn.ModuleList(output + layers[:stage])


_DROPPED_KWARGS = ('num_classes', 'global_pool', 'in_chans')  # resnet.py:567-568 kwargs_filter


def _create_resnet(variant, pretrained=False, **kwargs):
    for k in _DROPPED_KWARGS:
        kwargs.pop(k, None)
    model = ResNet(**kwargs)
    if pretrained:
        from ...constructor.load import load_pretrained
        load_pretrained(model, variant)
    return model


def _register(name, block, layers, **fixed):
    def factory(pretrained=False, **kwargs):
        return _create_resnet(name, pretrained, **dict(block=block, layers=layers, **fixed, **kwargs))
    factory.__name__ = name
    factory.__qualname__ = name
    factory.__doc__ = f'Constructs a {name} model ({block.__name__}, layers {layers}{", " + str(fixed) if fixed else ""}).'
    factory.__module__ = __name__
    globals()[name] = factory
    return BACKBONES.register_class(factory)


# Same factory names as the reference registers (resnet.py:590-596, 607-613, 624-630, 649-655, 675-681, 691-697,
# 708-714, 725-755, 757-778, 882-900, 942-962); the tv_/ssl_/swsl_ names are the same architectures with other
# pretrained weights.
for _n, _b, _l in [('resnet18', BasicBlock, [2, 2, 2, 2]), ('resnet34', BasicBlock, [3, 4, 6, 3]),
                   ('resnet26', Bottleneck, [2, 2, 2, 2]), ('resnet50', Bottleneck, [3, 4, 6, 3]),
                   ('resnet101', Bottleneck, [3, 4, 23, 3]), ('resnet152', Bottleneck, [3, 8, 36, 3]),
                   ('resnet200', Bottleneck, [3, 24, 36, 3]),
                   ('tv_resnet34', BasicBlock, [3, 4, 6, 3]), ('tv_resnet50', Bottleneck, [3, 4, 6, 3]),
                   ('tv_resnet101', Bottleneck, [3, 4, 23, 3]), ('tv_resnet152', Bottleneck, [3, 8, 36, 3]),
                   ('ssl_resnet18', BasicBlock, [2, 2, 2, 2]), ('ssl_resnet50', Bottleneck, [3, 4, 6, 3]),
                   ('swsl_resnet18', BasicBlock, [2, 2, 2, 2]), ('swsl_resnet50', Bottleneck, [3, 4, 6, 3])]:
    _register(_n, _b, _l)
_register('wide_resnet50_2', Bottleneck, [3, 4, 6, 3], base_width=128)
_register('wide_resnet101_2', Bottleneck, [3, 4, 23, 3], base_width=128)
