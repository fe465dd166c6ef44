"""Plain-torch CPU restatement of the reference's model path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED by reference goldens (shape-only tests upstream); cross-checked against torchvision.

Follows:
  ResNet ............ torchok/models/backbones/resnet.py:408-563 (+ make_blocks :363-405) with timm 0.6.13
                      BasicBlock / Bottleneck / downsample_conv semantics
  Pooling(Linear) ... torchok/models/poolings/classification/pooling.py:7-12, linear.py:8-25
  LinearHead ........ torchok/models/heads/representation/linear_head.py:10-36
  ClassificationHead  torchok/models/heads/classification/classification_head.py:9-40
  ArcFaceHead ....... torchok/models/heads/classification/arcface_head.py:12-131
  ConvBnAct ......... torchok/models/modules/bricks/convbnact.py:8-53
  ClassificationTask  torchok/tasks/classification.py:45-119
"""
import contextlib
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# ---- precision policy ------------------------------------------------------------------------------------------
# Default: plain fp32, exactly the torch.nn graph the reference builds.  Inside `amp_bf16()` the same graph is
# evaluated with the storage rounding of the reference's mixed-precision mode (`trainer.precision: 16`,
# examples/configs/classification_imagenet.yaml:120 -> torch autocast): conv / linear operands and every stored
# activation are rounded to bf16, accumulation and BatchNorm statistics stay fp32.  The CUDA path stores bf16
# activations at the same points, so this mode isolates kernel correctness from the precision policy.
_AMP = False


@contextlib.contextmanager
def amp_bf16(enabled=True):
    global _AMP
    prev, _AMP = _AMP, enabled
    try:
        yield
    finally:
        _AMP = prev


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)  # gradients are stored in bf16 as well


def q(t):
    return _RoundBF16.apply(t) if _AMP else t


def qw(t):
    """operand rounding without gradient rounding (fp32 master weights, bf16 compute copy)"""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach() if _AMP else t


def conv(m, x):
    return q(F.conv2d(x, qw(m.weight), m.bias, m.stride, m.padding, m.dilation, m.groups))


def linear(m, x):
    return q(F.linear(x, qw(m.weight), m.bias))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        shortcut = x
        x = q(F.relu(self.bn1(conv(self.conv1, x))))
        x = self.bn2(conv(self.conv2, x))
        if self.downsample is not None:
            shortcut = q(self.downsample[1](conv(self.downsample[0], shortcut)))
        return q(F.relu(x + shortcut))


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, base_width=64):
        super().__init__()
        width = int(math.floor(planes * (base_width / 64)))
        self.conv1 = nn.Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)  # stride on the 3x3 (v1.5)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        shortcut = x
        x = q(F.relu(self.bn1(conv(self.conv1, x))))
        x = q(F.relu(self.bn2(conv(self.conv2, x))))
        x = self.bn3(conv(self.conv3, x))
        if self.downsample is not None:
            shortcut = q(self.downsample[1](conv(self.downsample[0], shortcut)))
        return q(F.relu(x + shortcut))


class ResNet(nn.Module):
    def __init__(self, block, layers, in_channels=3, base_width=64, zero_init_last=True):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        inplanes = 64
        kw = dict(base_width=base_width) if block is Bottleneck else {}
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            ds = None
            if stride != 1 or inplanes != planes * block.expansion:
                ds = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, 1, stride, bias=False),
                                   nn.BatchNorm2d(planes * block.expansion))
            blocks = [block(inplanes, planes, stride, ds, **kw)]
            inplanes = planes * block.expansion
            blocks += [block(inplanes, planes, **kw) for _ in range(1, n)]
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
        self.out_channels = inplanes
        self.out_encoder_channels = (64,) + tuple(c * block.expansion for c in (64, 128, 256, 512))
        for m in self.modules():  # resnet.py:529-539
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        if zero_init_last:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.zeros_(m.bn3.weight)
                elif isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def forward_features(self, x):
        feats = [x]
        x = q(F.relu(self.bn1(conv(self.conv1, q(x)))))
        feats.append(x)
        x = self.maxpool(x)
        for i in range(4):
            x = getattr(self, f'layer{i + 1}')(x)
            feats.append(x)
        return feats

    def forward(self, x):
        return self.forward_features(x)[-1]


RESNETS = {
    'resnet18': (BasicBlock, [2, 2, 2, 2]), 'resnet34': (BasicBlock, [3, 4, 6, 3]),
    'resnet26': (Bottleneck, [2, 2, 2, 2]), 'resnet50': (Bottleneck, [3, 4, 6, 3]),
    'resnet101': (Bottleneck, [3, 4, 23, 3]), 'resnet152': (Bottleneck, [3, 8, 36, 3]),
}


def resnet(name, **kw):
    block, layers = RESNETS[name]
    return ResNet(block, layers, **kw)


class Pooling(nn.Module):
    def __init__(self, in_channels, pooling_type='avg'):
        super().__init__()
        self.pooling_type = pooling_type
        self.out_channels = in_channels * (2 if pooling_type == 'catavgmax' else 1)

    def forward(self, x):
        avg, mx = x.mean((2, 3)), x.amax((2, 3))
        return q({'avg': avg, 'max': mx, 'avgmax': 0.5 * (avg + mx), 'catavgmax': torch.cat([avg, mx], 1)}[
            self.pooling_type])


class PoolingLinear(Pooling):
    def __init__(self, in_channels, out_channels, pooling_type='avg', bias=True):
        super().__init__(in_channels, pooling_type)
        self.fc = nn.Linear(self.out_channels, out_channels, bias=bias)
        self.out_channels = out_channels
        nn.init.normal_(self.fc.weight, 0, 0.01)
        if bias:
            nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return linear(self.fc, super().forward(x))


class LinearHead(nn.Module):
    def __init__(self, in_channels, out_channels, drop_rate=0.0, bias=True, normalize=False):
        super().__init__()
        self.drop_rate, self.normalize, self.out_channels = drop_rate, normalize, out_channels
        self.fc = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, target=None):
        if self.drop_rate > 0.:
            x = F.dropout(x, self.drop_rate, self.training)
        x = linear(self.fc, x)
        if self.normalize:
            x = q(F.normalize(x, p=2, dim=-1))
        return x[..., 0] if getattr(self, 'squeeze_single', False) and self.out_channels == 1 else x


class ClassificationHead(LinearHead):
    squeeze_single = True

    def __init__(self, in_channels, num_classes, drop_rate=0.0, bias=True):
        super().__init__(in_channels, num_classes, drop_rate, bias)


class ArcFaceHead(nn.Module):
    """arcface_head.py:47-56 (default scale / margin), :95-108 (margin), :120-131 (forward)."""

    def __init__(self, in_channels, num_classes, scale=None, margin=None, easy_margin=False):
        super().__init__()
        if scale is None:
            c1 = num_classes - 1
            scale = c1 / num_classes * math.log(c1 * .999 / (1 - .999)) + 1
        if margin is None:
            margin = .9 - math.cos(2 * math.pi / num_classes) if in_channels == 2 else \
                .5 * num_classes / (num_classes - 1)
        self.scale, self.margin, self.easy_margin = scale, margin, easy_margin
        self.weight = nn.Parameter(torch.zeros(num_classes, in_channels))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x, target=None):
        if not self.training:
            return F.linear(x, self.weight)
        if target is None:
            raise ValueError('Target is None in training mode.')
        cosine = F.linear(F.normalize(x), F.normalize(self.weight))
        sine = torch.sqrt((1.0 - cosine ** 2).clamp(0, 1))
        m = self.margin
        phi = cosine * math.cos(m) - sine * math.sin(m)
        if self.easy_margin:
            phi = torch.where(cosine > 0, phi, cosine)
        else:
            phi = torch.where(cosine > math.cos(math.pi - m), phi, cosine - math.sin(math.pi - m) * m)
        one_hot = torch.zeros_like(cosine).scatter_(1, target.view(-1, 1).long(), 1)
        return torch.where(one_hot == 1, phi, cosine) * self.scale


class ConvBnAct(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, bias=False, use_batchnorm=True,
                 act=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        self.act = nn.ReLU() if act else nn.Identity()

    def forward(self, x):
        return q(self.act(self.bn(conv(self.conv, x))))


class ClassificationTask(nn.Module):
    def __init__(self, backbone, pooling=None, head=None, neck=None):
        super().__init__()
        self.backbone, self.neck = backbone, neck or nn.Identity()
        self.pooling, self.head = pooling or nn.Identity(), head or nn.Identity()

    def forward_with_gt(self, batch):
        emb = self.pooling(self.neck(self.backbone(batch['image'])))
        pred = self.head(emb, batch.get('target')) if not isinstance(self.head, nn.Identity) else emb
        return {'embeddings': emb, 'prediction': pred, 'target': batch.get('target')}


def dedegenerate_(model, seed=0):
    """zero-init-last makes fresh residual branches contribute 0; randomise BN affine + running stats so
    that parity tests exercise every branch.  gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return model


# ---- pairwise path (torchok/losses/representation/pairwise.py:9-136, torchok/tasks/pairwise_task.py:87-107) --------
class ContrastiveLoss(nn.Module):
    """S = cdist(emb1, emb2); L_i = sum_j (1-R) relu(mu-S)^2 + R S^2; optional L1/L2 regulariser; mean/sum."""

    def __init__(self, margin, reg=None, reduction='mean', eps=1e-3):
        super().__init__()
        self.margin, self.reg, self.reduction, self.eps = margin, reg, reduction, eps

    def forward(self, emb1, emb2, R):
        S = torch.cdist(emb1, emb2, p=2, compute_mode='donot_use_mm_for_euclid_dist')
        L = ((1. - R) * F.relu(self.margin - S).pow(2) + R * S.pow(2)).sum(1)
        if self.reg == 'L1':
            L = L + self.eps * emb1.abs().sum(1)
        elif self.reg == 'L2':
            L = L + self.eps * torch.norm(emb1, p=None, dim=1)
        elif self.reg is not None:
            raise ValueError(f'Unknown regularization type: {self.reg}')
        if self.reduction == 'mean':
            return L.mean()
        if self.reduction == 'sum':
            return L.sum()
        raise ValueError(f'Unknown reduction type: {self.reduction}')


def calc_relevance_matrix(y, num_classes):
    """pairwise_task.py:87-107: one-hot -> y y^T > 0."""
    if y.ndim == 1:
        y = torch.zeros(y.shape[0], num_classes).scatter_(1, y[:, None], 1)
    return torch.where(torch.matmul(y, y.transpose(1, 0)) > 0, 1., 0.)
