"""FPN neck (torchok/models/necks/detection/fpn.py:9-117 = mmdet 3.0.0 `necks.FPN` with reversed `in_channels`;
mmdet is not vendored in the reference, semantics restated from SURVEY Appendix A.4).

lateral 1x1 convs (bias, no norm / activation: mmdet ConvModule defaults) -> top-down pathway
`lateral[i-1] += nearest_upsample(lateral[i])` -> 3x3 output convs -> extra levels (stride-2 "max pool" of kernel 1,
or stride-2 3x3 convs on input / lateral / output).  The lateral add with its nearest-neighbour upsampling is ONE
pass (tok_fuse_sum_fwd) per level; convs run on the tcgen05 implicit-GEMM kernel with the bias in the epilogue.
Registered in DETECTION_NECKS (as in the reference) and in NECKS, so that it is reachable from tasks that look necks
up there.  State-dict keys follow mmdet: `lateral_convs.{i}.conv.*`, `fpn_convs.{i}.conv.*`.
"""
import torch.nn as nn

from ... import kernels as K
from ...constructor import DETECTION_NECKS, NECKS
from ..base import BaseModel
from ..modules.layers import Conv2d, ConvFn


class ConvModule(nn.Module):
    """mmcv ConvModule with norm_cfg=None, act_cfg=None: a biased Conv2d under `.conv`."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=True)

    def forward(self, x):
        return ConvFn.apply(x, self.conv, False, self.conv.weight, self.conv.bias)


class FPN(BaseModel):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=None, init_cfg=None):
        in_channels = list(in_channels[::-1])  # fpn.py:62 of the reference
        super().__init__(in_channels, out_channels)
        if conv_cfg is not None or norm_cfg is not None or act_cfg is not None:
            raise NotImplementedError('FPN: conv_cfg / norm_cfg / act_cfg other than the mmdet defaults (None)')
        upsample_cfg = dict(upsample_cfg or dict(mode='nearest'))
        if upsample_cfg.get('mode', 'nearest') != 'nearest':
            raise NotImplementedError('FPN: only nearest-neighbour top-down upsampling (the mmdet default)')
        self.upsample_cfg = upsample_cfg
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.no_norm_on_lateral = no_norm_on_lateral
        if end_level == -1 or end_level == self.num_ins - 1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level + 1
            assert end_level < self.num_ins
            assert num_outs == end_level - start_level + 1
        self.start_level, self.end_level = start_level, end_level
        assert isinstance(add_extra_convs, (str, bool))
        if isinstance(add_extra_convs, str):
            assert add_extra_convs in ('on_input', 'on_lateral', 'on_output')
        elif add_extra_convs:
            add_extra_convs = 'on_input'
        self.add_extra_convs = add_extra_convs
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1))
        extra_levels = num_outs - self.backbone_end_level + self.start_level
        if self.add_extra_convs and extra_levels >= 1:
            for i in range(extra_levels):
                cin = in_channels[self.backbone_end_level - 1] if (i == 0 and self.add_extra_convs == 'on_input') \
                    else out_channels
                self.fpn_convs.append(ConvModule(cin, out_channels, 3, stride=2, padding=1))
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):  # mmdet init_cfg: Xavier uniform on Conv2d
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        laterals = [conv(inputs[i + self.start_level]) for i, conv in enumerate(self.lateral_convs)]
        used = len(laterals)
        for i in range(used - 1, 0, -1):
            # lateral add + nearest upsample in one pass (scale must be a power of two, as it is for every backbone here)
            laterals[i - 1] = K.fuse_sum([laterals[i - 1], laterals[i]], relu=False)
        outs = [self.fpn_convs[i](laterals[i]) for i in range(used)]
        if self.num_outs > len(outs):
            if not self.add_extra_convs:
                for _ in range(self.num_outs - used):
                    outs.append(outs[-1][:, :, ::2, ::2])  # F.max_pool2d(x, 1, stride=2)
            else:
                if self.add_extra_convs == 'on_input':
                    extra_source = inputs[self.backbone_end_level - 1]
                elif self.add_extra_convs == 'on_lateral':
                    extra_source = laterals[-1]
                else:
                    extra_source = outs[-1]
                outs.append(self.fpn_convs[used](extra_source))
                for i in range(used + 1, self.num_outs):
                    src = K.fuse_sum([outs[-1]], relu=True) if self.relu_before_extra_convs else outs[-1]
                    outs.append(self.fpn_convs[i](src))
        return tuple(outs)


DETECTION_NECKS.register_class(FPN)
NECKS.register_class(FPN)
