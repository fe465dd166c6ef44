"""Segmentation heads.

SegmentationHead (torchok/models/heads/segmentation/base.py:11-41): 1x1 classifier conv (with bias) on the neck
features, bilinear resize (align_corners=False) to the input image size, channel squeeze for a single class.
OCRSegmentationHead (torchok/models/heads/segmentation/ocr.py:22-192): see the class docstring."""
import torch.nn as nn

from ... import kernels as K
from ...constructor import HEADS
from ..base import BaseModel
from ..modules.bricks import ConvBnAct
from ..modules.layers import Conv2d


@HEADS.register_class
class SegmentationHead(BaseModel):
    def __init__(self, in_channels, num_classes, do_interpolate=True):
        super().__init__(in_channels, num_classes)
        self.num_classes = num_classes
        self.do_interpolate = do_interpolate
        self.classifier = Conv2d(in_channels, num_classes, kernel_size=1)
        self.init_weights()

    def forward(self, x):
        input_image, features = x
        segm_logits = self.classifier(features)
        if self.do_interpolate:
            segm_logits = K.bilinear_resize(segm_logits, input_image.shape[2:])
        if self.num_classes == 1:
            segm_logits = segm_logits[:, 0]
        return segm_logits


class _ObjectAttentionBlock(nn.Module):
    """ObjectAttentionBlock at scale 1 (torchok/models/heads/segmentation/ocr.py:49-102); module names as upstream."""

    def __init__(self, in_channels, key_channels):
        super().__init__()
        self.key_channels = key_channels

        def two(a, b):
            return nn.Sequential(ConvBnAct(a, b, kernel_size=1), ConvBnAct(b, b, kernel_size=1))
        self.f_pixel, self.f_object, self.f_down = (two(in_channels, key_channels) for _ in range(3))
        self.f_up = ConvBnAct(key_channels, in_channels, kernel_size=1)

    def forward(self, x, proxy):
        query = self.f_pixel(x)                              # (B, Kc, H, W)
        key, value = self.f_object(proxy), self.f_down(proxy)    # (B, Kc, K, 1): BatchNorm statistics over B*K "pixels"
        context = K.object_attention(query, key, value, self.key_channels ** -.5)
        return self.f_up(context)


class _SpatialOCR(nn.Module):
    """SpatialOCR (ocr.py:105-130): object attention, concat with the pixel features, 1x1 ConvBnReLU, Dropout2d."""

    def __init__(self, in_channels, key_channels, out_channels, dropout):
        super().__init__()
        self.object_context_block = _ObjectAttentionBlock(in_channels, key_channels)
        self.conv_bn_dropout = nn.Sequential(ConvBnAct(2 * in_channels, out_channels, kernel_size=1),
                                             nn.Dropout2d(dropout))
        self.conv_bn_dropout[0].conv.set_input_layout([in_channels, in_channels])

    def forward(self, feats, proxy):
        context = self.object_context_block(feats, proxy)
        cat = K.bilinear_cat([context, feats], (feats.size(2), feats.size(3)))     # same size: an exact copy + concat
        out = self.conv_bn_dropout[0](cat)
        return K.dropout2d(out, self.conv_bn_dropout[1].p, self.training)


@HEADS.register_class
class OCRSegmentationHead(BaseModel):
    """HRNet-OCR head (torchok/models/heads/segmentation/ocr.py:133-192): auxiliary class maps -> soft object regions
    (tok_spatial_gather) -> object attention (tok_object_attn) -> fused 1x1 units -> classifier; returns
    (out, out_aux) in training mode and `out` in eval mode, resized to the input image."""

    def __init__(self, in_channels, num_classes, do_interpolate=True, ocr_mid_channels=128, ocr_key_channels=64):
        super().__init__(in_channels, num_classes)
        self.do_interpolate, self.num_classes = do_interpolate, num_classes
        self.conv3x3_ocr = ConvBnAct(in_channels, ocr_mid_channels, kernel_size=3, padding=1)
        self.ocr_distri_head = _SpatialOCR(ocr_mid_channels, ocr_key_channels, ocr_mid_channels, 0.05)
        self.last_reduction = ConvBnAct(ocr_mid_channels, ocr_mid_channels // 16, kernel_size=1, stride=1, padding=0)
        self.aux_head = nn.Sequential(ConvBnAct(in_channels, in_channels, kernel_size=1, stride=1, padding=0),
                                      Conv2d(in_channels, num_classes, kernel_size=1, stride=1, padding=0, bias=True))
        self.classifier = Conv2d(ocr_mid_channels // 16, num_classes, kernel_size=1)
        self.init_weights()

    def forward(self, x):
        input_image, feats = x
        out_aux = self.aux_head(feats)
        feats = self.conv3x3_ocr(feats)
        context = K.spatial_gather(feats, out_aux)           # (B, mid, K, 1)
        feats = self.ocr_distri_head(feats, context)
        out = self.classifier(self.last_reduction(feats))
        if self.do_interpolate:
            out = K.bilinear_resize(out, input_image.shape[2:])
            out_aux = K.bilinear_resize(out_aux, input_image.shape[2:])
        if self.num_classes == 1:
            out, out_aux = out[:, 0], out_aux[:, 0]
        return (out, out_aux) if self.training else out
