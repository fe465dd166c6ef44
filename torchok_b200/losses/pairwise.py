"""ContrastiveLoss and its bases (torchok/losses/representation/pairwise.py:9-136): same class hierarchy, arguments,
regularisation / reduction semantics and error messages; `calc_loss` runs on tok_contrastive_fwd / _bwd (fp32
pairwise distances by direct differences, per-row loss, analytic gradient)."""
import torch
from torch.nn import Module

from .. import kernels as K
from ..constructor import LOSSES


class BasePairwiseLoss(Module):
    def __init__(self, reg=None, reduction='mean', eps=1e-3):
        super().__init__()
        self.reg = reg
        self.reduction = reduction
        self.eps = eps

    def regularize(self, L, emb):
        if self.reg is None:
            return L
        elif self.reg == 'L1':
            return L + self.eps * emb.float().abs().sum(1)
        elif self.reg == 'L2':
            return L + self.eps * torch.norm(emb.float(), p=None, dim=1)
        else:
            raise ValueError(f'Unknown regularization type: {self.reg}')

    def apply_reduction(self, L):
        if self.reduction == 'mean':
            L = L.mean()
        elif self.reduction == 'sum':
            L = L.sum()
        else:
            raise ValueError(f'Unknown reduction type: {self.reduction}')
        return L


class GeneralPairWeightingLoss(BasePairwiseLoss):
    def __init__(self, margin, reg=None, reduction='mean', eps=1e-3):
        super().__init__(reg=reg, reduction=reduction, eps=eps)
        self.margin = margin

    def forward(self, emb1, emb2, R):
        L = self.calc_loss(emb1, emb2, R)
        L = self.regularize(L, emb1)
        L = self.apply_reduction(L)
        return L

    def calc_loss(self, emb1, emb2, R):
        raise NotImplementedError()


@LOSSES.register_class
class ContrastiveLoss(GeneralPairWeightingLoss):
    def calc_loss(self, emb1, emb2, R):
        return K.contrastive_rows(emb1, emb2, R, self.margin)
