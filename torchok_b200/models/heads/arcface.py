"""ArcFaceHead (torchok/models/heads/classification/arcface_head.py:12-131).

Same constructor defaults (scale = (C-1)/C * ln((C-1) p/(1-p)) + 1 with p = .999; margin = .5 C/(C-1), or
.9 - cos(2 pi / C) for 2-D embeddings), same `weight` parameter (C x D, xavier-uniform) and same forward contract:
eval -> un-normalised `x @ W^T`; training -> `scale * where(onehot(target), phi(cos), cos)` and ValueError without a
target.  The cosine GEMM runs on tok_linear_* (tcgen05), the margin on tok_arcface_margin_*.

Deviation: the reference's `dynamic_margin=True` reads `self.__step`, an attribute that never exists
(the buffer is registered as `step`), so it raises AttributeError on the first training step.  Here the intended
behaviour is implemented: the margin grows linearly from `min_margin` to `margin` over `num_warmup_steps` steps.
"""
import math

import torch
import torch.nn as nn

from ... import kernels as K
from ...constructor import HEADS
from ..base import BaseModel


@HEADS.register_class
class ArcFaceHead(BaseModel):
    def __init__(self, in_channels, num_classes, scale=None, margin=None, easy_margin=False, dynamic_margin=False,
                 num_warmup_steps=None, min_margin=None):
        super().__init__(in_channels, out_channels=num_classes)
        if scale is None:
            p = .999
            c_1 = (num_classes - 1)
            scale = c_1 / num_classes * math.log(c_1 * p / (1 - p)) + 1
        if margin is None:
            if in_channels == 2:
                margin = .9 - math.cos(2 * math.pi / num_classes)
            else:
                margin = .5 * num_classes / (num_classes - 1)
        self.dynamic_margin = dynamic_margin
        if self.dynamic_margin:
            if num_warmup_steps is None or not isinstance(num_warmup_steps, int):
                raise ValueError('`num_warmup_steps` must be positive int when `dynamic_margin` is True')
            if min_margin is None:
                raise ValueError('`min_margin` must be float when `dynamic_margin` is True')
            self.num_warmup_steps = num_warmup_steps
            self.min_margin = min_margin
            self.max_margin = margin
            self.margin = min_margin
            self.register_buffer('step', torch.tensor(0))
        else:
            self.margin = margin
        self.scale = scale
        self.easy_margin = easy_margin
        self.weight = nn.Parameter(torch.zeros(num_classes, in_channels), requires_grad=True)
        nn.init.xavier_uniform_(self.weight)

    def _update_margin(self):
        if self.dynamic_margin and self.step.is_cuda and torch.cuda.is_current_stream_capturing():
            # the schedule is host arithmetic on a device counter: captured, the margin would be frozen into the graph
            # while `step` kept advancing on every replay
            raise NotImplementedError('ArcFaceHead(dynamic_margin=True) cannot run inside a captured step: build the '
                                      'loop with StreamLoop(use_graph=False) (or TOK_NO_GRAPH=1)')
        if self.dynamic_margin and int(self.step) <= self.num_warmup_steps:
            frac = int(self.step) / self.num_warmup_steps
            self.margin = self.min_margin + frac * (self.max_margin - self.min_margin)
            self.step += 1

    def forward(self, input, target=None):
        if not self.training:
            return K.linear(input, self.weight, None)
        elif target is None:
            raise ValueError('Target is None in training mode.')
        output = K.arcface(input, self.weight, target, self.scale, self.margin, self.easy_margin)
        self._update_margin()
        return output
