"""SegmentationHead (torchok/models/heads/segmentation/base.py:11-41): 1x1 classifier conv (with bias) on the neck
features, bilinear resize (align_corners=False) to the input image size, channel squeeze for a single class."""
import torch.nn as nn

from ... import kernels as K
from ...constructor import HEADS
from ..base import BaseModel
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
