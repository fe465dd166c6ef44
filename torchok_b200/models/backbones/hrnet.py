"""HRNet backbones on the sm_100a kernels.

Mirror of torchok/models/backbones/hrnet.py:49-255 (HighResolutionNet: two 3x3 stride-2 stem convs, layer1, three
transitions, stages 2-4) with timm 0.6.13's HighResolutionModule / cfg_cls restated (timm is not
vendored in the reference).  Module / parameter names follow timm (`stage3.1.branches.2.0.conv1.weight`,
`stage2.0.fuse_layers.1.0.0.0.weight`, `transition1.1.0.0.weight` ...) so reference checkpoints load unchanged.
Every conv+BN(+ReLU) is one fused unit, residual blocks are the ResNet ones, and the cross-resolution fuse
`relu(sum_j fuse_ij(x_j))` with its nearest-neighbour upsampling is ONE pass (tok_fuse_sum_fwd) per output branch.
Channel counts that are not multiples of 8 (18, 36, 30 ...) run zero-padded to the next multiple of 8 internally.
"""
import torch
import torch.nn as nn

from ... import kernels as K
from ...constructor import BACKBONES
from ..base import BaseBackbone
from ..modules.layers import BatchNorm2d, Conv2d, ReLU, conv_bn_act
from .resnet import BasicBlock, Bottleneck

_BN_MOMENTUM = 0.1
blocks_dict = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}


def _stage(modules, branches, block, blocks, channels):
    return dict(NUM_MODULES=modules, NUM_BRANCHES=branches, BLOCK=block, NUM_BLOCKS=tuple(blocks),
                NUM_CHANNELS=tuple(channels), FUSE_METHOD='SUM')


def _cfg(c, s1_blocks=4, s1_ch=64, blocks=4, mods=(1, 4, 3)):
    return dict(STEM_WIDTH=64,
                STAGE1=_stage(1, 1, 'BOTTLENECK', (s1_blocks,), (s1_ch,)),
                STAGE2=_stage(mods[0], 2, 'BASIC', (blocks,) * 2, (c, 2 * c)),
                STAGE3=_stage(mods[1], 3, 'BASIC', (blocks,) * 3, (c, 2 * c, 4 * c)),
                STAGE4=_stage(mods[2], 4, 'BASIC', (blocks,) * 4, (c, 2 * c, 4 * c, 8 * c)))


# timm.models.hrnet.cfg_cls (0.6.13)
cfg_cls = dict(
    hrnet_w18_small=_cfg(16, s1_blocks=1, s1_ch=32, blocks=2, mods=(1, 1, 1)),
    hrnet_w18_small_v2=_cfg(18, s1_blocks=2, s1_ch=64, blocks=2, mods=(1, 3, 2)),
    hrnet_w18=_cfg(18), hrnet_w30=_cfg(30), hrnet_w32=_cfg(32), hrnet_w40=_cfg(40), hrnet_w44=_cfg(44),
    hrnet_w48=_cfg(48), hrnet_w64=_cfg(64),
)


class ConvBnSeq(nn.Sequential):
    """nn.Sequential(Conv2d, BatchNorm2d[, ReLU | Upsample]) executed as one fused unit.  A trailing nn.Upsample is a
    marker: the nearest-neighbour upsampling itself happens inside the consumer's fuse pass."""

    def forward(self, x):
        relu = len(self) > 2 and isinstance(self[2], nn.ReLU)
        return conv_bn_act(x, self[0], self[1], relu=relu)


def _conv_bn(cin, cout, k, stride, pad, tail=None):
    mods = [Conv2d(cin, cout, k, stride, pad, bias=False), BatchNorm2d(cout, momentum=_BN_MOMENTUM)]
    if tail == 'relu':
        mods.append(ReLU(inplace=True))
    elif tail is not None:
        mods.append(tail)
    return ConvBnSeq(*mods)


class HighResolutionModule(nn.Module):
    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True):
        super().__init__()
        if num_branches != len(num_blocks) or num_branches != len(num_channels) or num_branches != len(num_inchannels):
            raise ValueError('NUM_BRANCHES does not match NUM_BLOCKS / NUM_CHANNELS / NUM_INCHANNELS')
        self.num_inchannels = list(num_inchannels)
        self.fuse_method = fuse_method
        self.num_branches = num_branches
        self.multi_scale_output = multi_scale_output
        self.branches = nn.ModuleList([self._make_one_branch(i, blocks, num_blocks, num_channels)
                                       for i in range(num_branches)])
        self.fuse_layers = self._make_fuse_layers()
        self.fuse_act = ReLU(False)

    def _make_one_branch(self, i, block, num_blocks, num_channels, stride=1):
        downsample = None
        if stride != 1 or self.num_inchannels[i] != num_channels[i] * block.expansion:
            downsample = nn.Sequential(
                Conv2d(self.num_inchannels[i], num_channels[i] * block.expansion, 1, stride=stride, bias=False),
                BatchNorm2d(num_channels[i] * block.expansion, momentum=_BN_MOMENTUM))
        layers = [block(self.num_inchannels[i], num_channels[i], stride, downsample)]
        self.num_inchannels[i] = num_channels[i] * block.expansion
        for _ in range(1, num_blocks[i]):
            layers.append(block(self.num_inchannels[i], num_channels[i]))
        return nn.Sequential(*layers)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return nn.Identity()
        ch = self.num_inchannels
        fuse_layers = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:
                    row.append(_conv_bn(ch[j], ch[i], 1, 1, 0, nn.Upsample(scale_factor=2 ** (j - i), mode='nearest')))
                elif j == i:
                    row.append(nn.Identity())
                else:
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        chain.append(_conv_bn(ch[j], ch[i] if last else ch[j], 3, 2, 1, None if last else 'relu'))
                    row.append(nn.Sequential(*chain))
            fuse_layers.append(nn.ModuleList(row))
        return nn.ModuleList(fuse_layers)

    def get_num_in_chs(self):
        return self.num_inchannels

    def forward(self, x):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        x = [branch(x[i]) for i, branch in enumerate(self.branches)]
        out = []
        for i, row in enumerate(self.fuse_layers):
            terms = [None] * self.num_branches
            for j in range(self.num_branches):
                terms[j] = x[j] if j == i else row[j](x[j])
            # the full-resolution term first (it defines the output size of the fused pass)
            order = [i] + [j for j in range(self.num_branches) if j != i]
            out.append(K.fuse_sum([terms[j] for j in order], relu=True))
        return out


class HighResolutionNet(BaseBackbone):
    def __init__(self, cfg, in_channels=3):
        super().__init__(in_channels=in_channels, out_channels=cfg['STAGE4']['NUM_CHANNELS'])
        self._out_encoder_channels = cfg['STAGE4']['NUM_CHANNELS']
        stem_width = cfg['STEM_WIDTH']
        self.conv1 = Conv2d(in_channels, stem_width, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = BatchNorm2d(stem_width, momentum=_BN_MOMENTUM)
        self.act1 = ReLU(inplace=True)
        self.conv2 = Conv2d(stem_width, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn2 = BatchNorm2d(64, momentum=_BN_MOMENTUM)
        self.act2 = ReLU(inplace=True)

        self.stage1_cfg = cfg['STAGE1']
        num_channels = self.stage1_cfg['NUM_CHANNELS'][0]
        block = blocks_dict[self.stage1_cfg['BLOCK']]
        self.layer1 = self._make_layer(block, 64, num_channels, self.stage1_cfg['NUM_BLOCKS'][0])
        stage1_out_channel = block.expansion * num_channels

        pre = [stage1_out_channel]
        for idx in (2, 3, 4):
            scfg = cfg[f'STAGE{idx}']
            setattr(self, f'stage{idx}_cfg', scfg)
            block = blocks_dict[scfg['BLOCK']]
            num_channels = [c * block.expansion for c in scfg['NUM_CHANNELS']]
            setattr(self, f'transition{idx - 1}', self._make_transition_layer(pre, num_channels))
            stage, pre = self._make_stage(scfg, num_channels, multi_scale_output=True)
            setattr(self, f'stage{idx}', stage)
        self.init_weights()

    @torch.no_grad()
    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    @staticmethod
    def _make_transition_layer(pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                layers.append(_conv_bn(pre[i], cur[i], 3, 1, 1, 'relu') if cur[i] != pre[i] else nn.Identity())
            else:
                chain = []
                for j in range(i + 1 - len(pre)):
                    cout = cur[i] if j == i - len(pre) else pre[-1]
                    chain.append(_conv_bn(pre[-1], cout, 3, 2, 1, 'relu'))
                layers.append(nn.Sequential(*chain))
        return nn.ModuleList(layers)

    @staticmethod
    def _make_layer(block, in_channels, out_channels, num_blocks, stride=1):
        downsample = None
        if stride != 1 or in_channels != out_channels * block.expansion:
            downsample = nn.Sequential(
                Conv2d(in_channels, out_channels * block.expansion, kernel_size=1, stride=stride, bias=False),
                BatchNorm2d(out_channels * block.expansion, momentum=_BN_MOMENTUM))
        layers = [block(in_channels, out_channels, stride, downsample)]
        in_channels = out_channels * block.expansion
        layers += [block(in_channels, out_channels) for _ in range(1, num_blocks)]
        return nn.Sequential(*layers)

    @staticmethod
    def _make_stage(layer_config, in_channels, multi_scale_output=True):
        block = blocks_dict[layer_config['BLOCK']]
        modules = []
        for i in range(layer_config['NUM_MODULES']):
            reset = multi_scale_output or i < layer_config['NUM_MODULES'] - 1
            modules.append(HighResolutionModule(layer_config['NUM_BRANCHES'], block, layer_config['NUM_BLOCKS'],
                                                in_channels, layer_config['NUM_CHANNELS'],
                                                layer_config['FUSE_METHOD'], reset))
            in_channels = modules[-1].get_num_in_chs()
        return nn.Sequential(*modules), in_channels

    @staticmethod
    def _run_stage(stage, xl):
        for module in stage:
            xl = module(xl)
        return xl

    def forward_stem(self, x):
        x = conv_bn_act(x, self.conv1, self.bn1, relu=True)
        return conv_bn_act(x, self.conv2, self.bn2, relu=True)

    def forward_stages(self, x):
        x = self.layer1(x)
        xl = [t(x) for t in self.transition1]
        yl = self._run_stage(self.stage2, xl)
        xl = [t(yl[-1]) if not isinstance(t, nn.Identity) else yl[i] for i, t in enumerate(self.transition2)]
        yl = self._run_stage(self.stage3, xl)
        xl = [t(yl[-1]) if not isinstance(t, nn.Identity) else yl[i] for i, t in enumerate(self.transition3)]
        return self._run_stage(self.stage4, xl)

    def _forward_collect(self, x):
        return self.forward_stages(self.forward_stem(x))

    def forward(self, x):
        return self._forward_collect(x)

    def forward_features(self, x):
        return [x] + self.forward(x)

    @property
    def out_encoder_channels(self):
        return tuple(self._out_encoder_channels)

    def get_stages(self, stage):
        output = [self.conv1, self.bn1, self.act1, self.conv2, self.bn2, self.act2]
        layers = [[self.layer1], [self.transition1, self.stage2], [self.transition2, self.stage3],
                  [self.transition3, self.stage4]]
        for i in range(stage):
            output += layers[i]
        return nn.ModuleList(output)


def _create_hrnet(variant, pretrained=False, **kwargs):
    for k in ('num_classes', 'global_pool', 'in_chans'):  # hrnet.py:265 kwargs_filter
        kwargs.pop(k, None)
    model = HighResolutionNet(cfg_cls[variant], **kwargs)
    if pretrained:
        from ...constructor.load import load_pretrained
        load_pretrained(model, variant)
    return model


def _register(name):
    def factory(pretrained=False, **kwargs):
        return _create_hrnet(name, pretrained, **kwargs)
    factory.__name__ = factory.__qualname__ = name
    factory.__doc__ = f"It's constructing a {name} model."
    factory.__module__ = __name__
    globals()[name] = factory
    return BACKBONES.register_class(factory)


for _n in cfg_cls:
    _register(_n)
