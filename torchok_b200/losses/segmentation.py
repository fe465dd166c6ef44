"""Segmentation losses under the reference's registry names.

DiceLoss — torchok/losses/segmentation/dice.py:85-188 (adapted there from pytorch-toolbelt): the `multiclass` mode on raw
logits, which is what the reference's segmentation example uses next to CrossEntropyLoss
(examples/configs/segmentation_sweet_pepper.yaml:21-27), runs on the sm_100a kernels (tok_dice_stats / _finalize / _bwd).
The binary / multilabel modes, `from_logits=False` and a `classes` subset are outside the hot-path scope and raise.
"""
import torch.nn as nn

from .. import kernels as K
from ..constructor import LOSSES

BINARY_MODE, MULTICLASS_MODE, MULTILABEL_MODE = 'binary', 'multiclass', 'multilabel'


@LOSSES.register_class
class DiceLoss(nn.Module):
    def __init__(self, mode, classes=None, log_loss=False, from_logits=True, smooth=0, eps=1e-7):
        super().__init__()
        if mode not in {BINARY_MODE, MULTILABEL_MODE, MULTICLASS_MODE}:
            raise ValueError(f'DiceLoss initialize. Mode {mode} does not supper. Please choose one of from'
                             f'{[BINARY_MODE, MULTILABEL_MODE, MULTICLASS_MODE]}.')
        if classes is not None and mode == BINARY_MODE:
            raise ValueError('DiceLoss initialize. Masking classes is not supported with mode=binary')
        if mode != MULTICLASS_MODE or classes is not None or not from_logits:
            raise NotImplementedError('DiceLoss: the kernels cover mode="multiclass", from_logits=True, all classes')
        self.mode, self.classes, self.from_logits = mode, classes, from_logits
        self.log_loss, self.smooth, self.eps = log_loss, smooth, eps

    def forward(self, input, target):
        if input[:, 0].shape != target.shape:
            raise ValueError(f"Shapes of input {input.shape} and target {target.shape} tensors don't match!")
        return K.dice_multiclass(input, target, self.smooth, self.eps, self.log_loss)
