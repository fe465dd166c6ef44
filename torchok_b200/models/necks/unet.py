"""UnetNeck (torchok/models/necks/segmentation/unet.py:20-131, use_attention=False).

A decoder block is `nearest x2 upsample -> (nearest resize of the skip) -> cat -> ConvBnAct 3x3 -> ConvBnAct 3x3`
(unet.py:40-58).  Upsample, skip resize and concat are ONE gather pass per input (tok_nearest_fwd) writing straight into
the padded NHWC concat buffer the first 3x3 reads; both 3x3 units are the fused conv + BatchNorm + ReLU kernels (or
conv + bias + ReLU when use_batchnorm=False).  Submodule names follow the reference (center.0/1, blocks.i.conv1/conv2).
"""
import torch.nn as nn

from ... import kernels as K
from ...constructor import NECKS
from ..base import BaseModel
from ..modules.bricks import ConvBnAct


class DecoderBlock(nn.Module):
    def __init__(self, in_channels, skip_channels, out_channels, use_attention=False, use_batchnorm=True):
        super().__init__()
        if use_attention:
            raise NotImplementedError('UnetNeck: use_attention=True (SCSEModule) is outside the hot path (SURVEY 2)')
        self.attention1 = nn.Identity()
        self.conv1 = ConvBnAct(in_channels + skip_channels, out_channels, kernel_size=3, padding=1,
                               use_batchnorm=use_batchnorm)
        self.conv2 = ConvBnAct(out_channels, out_channels, kernel_size=3, padding=1, use_batchnorm=use_batchnorm)
        self.attention2 = nn.Identity()
        self._segments = [in_channels, skip_channels] if skip_channels else [in_channels]
        self.conv1.conv.set_input_layout(self._segments)

    def forward(self, x, skip=None):
        size = (2 * x.size(2), 2 * x.size(3))
        xs = [x] if skip is None else [x, skip]
        if len(xs) != len(self._segments):
            raise ValueError('DecoderBlock: skip connection does not match the construction-time skip_channels')
        x = K.nearest_cat(xs, size)
        return self.conv2(self.conv1(x))


@NECKS.register_class
class UnetNeck(BaseModel):
    def __init__(self, in_channels, decoder_channels=(512, 256, 128, 64, 64), use_batchnorm=True, use_attention=False,
                 center=True):
        super().__init__(in_channels=in_channels, out_channels=decoder_channels[-1])
        self.n_blocks = len(decoder_channels)
        encoder_channels = list(in_channels)[::-1]
        head_channels = encoder_channels[0]
        ins = [head_channels] + list(decoder_channels[:-1])
        skips = list(encoder_channels[1:]) + [0]
        if center:
            self.center = nn.Sequential(
                ConvBnAct(head_channels, head_channels, kernel_size=3, padding=1, use_batchnorm=use_batchnorm),
                ConvBnAct(head_channels, head_channels, kernel_size=3, padding=1, use_batchnorm=use_batchnorm))
        else:
            self.center = nn.Identity()
        self.blocks = nn.ModuleList(DecoderBlock(i, s, o, use_attention=use_attention, use_batchnorm=use_batchnorm)
                                    for i, s, o in zip(ins, skips, decoder_channels))
        self.init_weights()

    def forward(self, features):
        head, *skips, input_image = features[::-1]
        x = self.center(head)
        for i, block in enumerate(self.blocks):
            x = block(x, skips[i] if i < len(skips) else None)
        return [input_image, x]
