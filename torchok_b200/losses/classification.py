"""CrossEntropyLoss under the reference's registry name (torchok/losses/__init__.py:26 registers torch.nn's): fused
softmax + NLL + gradient in tok_softmax_xent.  Accepts (B, C) logits or (B, C, H, W) NHWC-backed logits with (B, H, W)
targets."""
import torch.nn as nn

from .. import kernels as K
from ..constructor import LOSSES


@LOSSES.register_class
class CrossEntropyLoss(nn.Module):
    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction='mean',
                 label_smoothing=0.0):
        super().__init__()
        if weight is not None or reduction != 'mean' or label_smoothing != 0.0:
            raise NotImplementedError('CrossEntropyLoss: class weights, reduction != mean and label smoothing are '
                                      'outside the hot-path scope')
        self.ignore_index = ignore_index
        self.reduction = reduction

    def forward(self, input, target):
        if input.dim() == 4:
            return K.softmax_xent_nhwc(input, target, self.ignore_index)
        return K.softmax_xent(input, target, self.ignore_index)
