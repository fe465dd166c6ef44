"""Plain-torch CPU restatement of the reference's model path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinning: see oracle/__init__.py — the torch-only reference files (ConvBnAct, heads, seg neck/head, losses) are PINNED by
reference-executed goldens; the timm / mmdet networks are PARITY UNPINNED by reference goldens (shape-only tests
upstream) and cross-checked against torchvision's independent ResNet / FPN.

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
        # timm SelectAdaptivePool2d: F.adaptive_avg_pool2d / F.adaptive_max_pool2d (the first maximum takes the gradient)
        avg, mx = F.adaptive_avg_pool2d(x, 1).flatten(1), F.adaptive_max_pool2d(x, 1).flatten(1)
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


def dice_loss_multiclass(logits, target, smooth=0.0, eps=1e-7, log_loss=False):
    """torchok/losses/segmentation/dice.py:85-188, DiceLoss(mode='multiclass', from_logits=True).forward with
    soft_dice_score (dice.py:23-56): log_softmax(dim=1).exp(), one-hot targets, per-class sums over (batch, pixels),
    score = (2 I + smooth) / (clamp_min(card, eps) + smooth), classes without a true pixel masked, mean over classes."""
    bs, nc = logits.shape[:2]
    p = logits.log_softmax(dim=1).exp().view(bs, nc, -1)
    t = F.one_hot(target.view(bs, -1), nc).permute(0, 2, 1).type_as(p)
    inter = (p * t).sum(dim=(0, 2))
    card = (p + t).sum(dim=(0, 2))
    score = (2.0 * inter + smooth) / (card.clamp_min(eps) + smooth)
    loss = -torch.log(score.clamp_min(eps)) if log_loss else 1 - score
    loss = loss * (t.sum(dim=(0, 2)) > 0).to(loss.dtype)
    return loss.mean()


def calc_relevance_matrix(y, num_classes):
    """pairwise_task.py:87-107: one-hot -> y y^T > 0."""
    if y.ndim == 1:
        y = torch.zeros(y.shape[0], num_classes).scatter_(1, y[:, None], 1)
    return torch.where(torch.matmul(y, y.transpose(1, 0)) > 0, 1., 0.)


# ---- HRNet (torchok/models/backbones/hrnet.py:49-255 + timm 0.6.13 HighResolutionModule / cfg_cls, Appendix A.2) --------
def _hr_stage(modules, branches, block, blocks, channels):
    return dict(NUM_MODULES=modules, NUM_BRANCHES=branches, BLOCK=block, NUM_BLOCKS=tuple(blocks),
                NUM_CHANNELS=tuple(channels))


def _hr_cfg(c, s1_blocks=4, s1_ch=64, blocks=4, mods=(1, 4, 3)):
    return dict(STEM_WIDTH=64, STAGE1=_hr_stage(1, 1, 'BOTTLENECK', (s1_blocks,), (s1_ch,)),
                STAGE2=_hr_stage(mods[0], 2, 'BASIC', (blocks,) * 2, (c, 2 * c)),
                STAGE3=_hr_stage(mods[1], 3, 'BASIC', (blocks,) * 3, (c, 2 * c, 4 * c)),
                STAGE4=_hr_stage(mods[2], 4, 'BASIC', (blocks,) * 4, (c, 2 * c, 4 * c, 8 * c)))


HRNET_CFGS = dict(hrnet_w18_small=_hr_cfg(16, 1, 32, 2, (1, 1, 1)), hrnet_w18_small_v2=_hr_cfg(18, 2, 64, 2, (1, 3, 2)),
                  hrnet_w18=_hr_cfg(18), hrnet_w32=_hr_cfg(32))
_HR_BLOCKS = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}


class _CB(nn.Sequential):
    """Conv2d -> BatchNorm2d [-> ReLU | Upsample] evaluated with the precision policy of this file."""

    def forward(self, x):
        x = self[1](conv(self[0], x))
        for m in list(self)[2:]:
            x = m(x)
        return q(x)


def _cb(cin, cout, k, s, p, tail=None):
    mods = [nn.Conv2d(cin, cout, k, s, p, bias=False), nn.BatchNorm2d(cout)]
    if tail is not None:
        mods.append(tail)
    return _CB(*mods)


def _hr_layer(block, cin, planes, n, stride=1):
    ds = None
    if stride != 1 or cin != planes * block.expansion:
        ds = nn.Sequential(nn.Conv2d(cin, planes * block.expansion, 1, stride, bias=False),
                           nn.BatchNorm2d(planes * block.expansion))
    layers = [block(cin, planes, stride, ds)]
    layers += [block(planes * block.expansion, planes) for _ in range(1, n)]
    return nn.Sequential(*layers)


class HighResolutionModule(nn.Module):
    def __init__(self, num_branches, block, num_blocks, num_inchannels, num_channels, multi_scale_output=True):
        super().__init__()
        self.num_branches, self.num_inchannels = num_branches, list(num_inchannels)
        branches = []
        for i in range(num_branches):
            branches.append(_hr_layer(block, self.num_inchannels[i], num_channels[i], num_blocks[i]))
            self.num_inchannels[i] = num_channels[i] * block.expansion
        self.branches = nn.ModuleList(branches)
        ch = self.num_inchannels
        if num_branches == 1:
            self.fuse_layers = nn.Identity()
        else:
            rows = []
            for i in range(num_branches if multi_scale_output else 1):
                row = []
                for j in range(num_branches):
                    if j > i:
                        row.append(_cb(ch[j], ch[i], 1, 1, 0, nn.Upsample(scale_factor=2 ** (j - i), mode='nearest')))
                    elif j == i:
                        row.append(nn.Identity())
                    else:
                        chain = []
                        for k in range(i - j):
                            last = k == i - j - 1
                            chain.append(_cb(ch[j], ch[i] if last else ch[j], 3, 2, 1, None if last else nn.ReLU()))
                        row.append(nn.Sequential(*chain))
                rows.append(nn.ModuleList(row))
            self.fuse_layers = nn.ModuleList(rows)

    def forward(self, x):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        x = [b(x[i]) for i, b in enumerate(self.branches)]
        out = []
        for i, row in enumerate(self.fuse_layers):
            y = x[0] if i == 0 else row[0](x[0])
            for j in range(1, self.num_branches):
                y = y + (x[j] if i == j else row[j](x[j]))
            out.append(q(F.relu(y)))
        return out


class HighResolutionNet(nn.Module):
    def __init__(self, cfg, in_channels=3):
        super().__init__()
        self.out_encoder_channels = tuple(cfg['STAGE4']['NUM_CHANNELS'])
        self.conv1 = nn.Conv2d(in_channels, cfg['STEM_WIDTH'], 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cfg['STEM_WIDTH'])
        self.conv2 = nn.Conv2d(cfg['STEM_WIDTH'], 64, 3, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        s1 = cfg['STAGE1']
        block = _HR_BLOCKS[s1['BLOCK']]
        self.layer1 = _hr_layer(block, 64, s1['NUM_CHANNELS'][0], s1['NUM_BLOCKS'][0])
        pre = [block.expansion * s1['NUM_CHANNELS'][0]]
        for idx in (2, 3, 4):
            sc = cfg[f'STAGE{idx}']
            block = _HR_BLOCKS[sc['BLOCK']]
            cur = [c * block.expansion for c in sc['NUM_CHANNELS']]
            trans = []
            for i in range(len(cur)):
                if i < len(pre):
                    trans.append(_cb(pre[i], cur[i], 3, 1, 1, nn.ReLU()) if cur[i] != pre[i] else nn.Identity())
                else:
                    chain = []
                    for j in range(i + 1 - len(pre)):
                        cout = cur[i] if j == i - len(pre) else pre[-1]
                        chain.append(_cb(pre[-1], cout, 3, 2, 1, nn.ReLU()))
                    trans.append(nn.Sequential(*chain))
            setattr(self, f'transition{idx - 1}', nn.ModuleList(trans))
            mods, inch = [], cur
            for _ in range(sc['NUM_MODULES']):
                mods.append(HighResolutionModule(sc['NUM_BRANCHES'], block, sc['NUM_BLOCKS'], inch, sc['NUM_CHANNELS']))
                inch = mods[-1].num_inchannels
            setattr(self, f'stage{idx}', nn.Sequential(*mods))
            pre = inch
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def forward(self, x):
        x = q(F.relu(self.bn1(conv(self.conv1, q(x)))))
        x = q(F.relu(self.bn2(conv(self.conv2, x))))
        x = self.layer1(x)
        xl = [t(x) for t in self.transition1]
        for m in self.stage2:
            xl = m(xl)
        xl = [t(xl[-1]) if not isinstance(t, nn.Identity) else xl[i] for i, t in enumerate(self.transition2)]
        for m in self.stage3:
            xl = m(xl)
        xl = [t(xl[-1]) if not isinstance(t, nn.Identity) else xl[i] for i, t in enumerate(self.transition3)]
        for m in self.stage4:
            xl = m(xl)
        return xl

    def forward_features(self, x):
        return [x] + self.forward(x)


def hrnet(name, in_channels=3):
    return HighResolutionNet(HRNET_CFGS[name], in_channels)


class HRNetSegmentationNeck(nn.Module):
    """necks/segmentation/hrnet.py:16-42"""

    def __init__(self, in_channels):
        super().__init__()
        self.out_channels = sum(in_channels)
        self.convbnact = ConvBnAct(self.out_channels, self.out_channels, 1)

    def forward(self, features):
        image, x0, x1, x2, x3 = features
        size = x0.shape[2:]
        up = [x0] + [q(F.interpolate(t, size=size, mode='bilinear', align_corners=False)) for t in (x1, x2, x3)]
        return [image, self.convbnact(torch.cat(up, 1))]


class HRNetClassificationNeck(nn.Module):
    """necks/classification/hrnet.py:12-85, overwrite quirk included."""

    def __init__(self, in_channels):
        super().__init__()
        hc = [32, 64, 128, 256]
        self.out_channels = 2048

        def layer(cin, planes):
            ds = ConvBnAct(cin, planes * 4, 1, act=False) if cin != planes * 4 else None
            return nn.Sequential(_NeckBottleneck(cin, planes, 1, ds))
        self.incre_modules = nn.ModuleList([layer(c, hc[i]) for i, c in enumerate(in_channels)])
        self.downsamp_modules = nn.ModuleList([ConvBnAct(hc[i] * 4, hc[i + 1] * 4, 3, padding=1, stride=2)
                                               for i in range(len(in_channels) - 1)])
        self.final_layer = ConvBnAct(hc[3] * 4, 2048, 1)

    def forward(self, x):
        y = self.incre_modules[0](x[0])
        for i in range(len(self.downsamp_modules)):
            y = self.downsamp_modules[i](y)
            if i + 1 < len(x):
                y = self.incre_modules[i + 1](x[i + 1])
        return self.final_layer(y)


class _NeckBottleneck(Bottleneck):
    """timm Bottleneck whose shortcut is a ConvBnAct(act_layer=None) module (necks/classification/hrnet.py:60-68)."""

    def forward(self, x):
        shortcut = x
        x = q(F.relu(self.bn1(conv(self.conv1, x))))
        x = q(F.relu(self.bn2(conv(self.conv2, x))))
        x = self.bn3(conv(self.conv3, x))
        if self.downsample is not None:
            shortcut = self.downsample(shortcut)
        return q(F.relu(x + shortcut))


class SegmentationHead(nn.Module):
    """heads/segmentation/base.py:11-41"""

    def __init__(self, in_channels, num_classes, do_interpolate=True):
        super().__init__()
        self.num_classes, self.do_interpolate = num_classes, do_interpolate
        self.classifier = nn.Conv2d(in_channels, num_classes, 1)

    def forward(self, x):
        image, feats = x
        logits = conv(self.classifier, feats)
        if self.do_interpolate:
            logits = q(F.interpolate(logits, size=image.shape[2:], mode='bilinear'))
        return logits[:, 0] if self.num_classes == 1 else logits


class SegmentationTask(nn.Module):
    """tasks/segmentation.py:67-94"""

    def __init__(self, backbone, neck, head):
        super().__init__()
        self.backbone, self.neck, self.head = backbone, neck, head

    def forward_with_gt(self, batch):
        pred = self.head(self.neck(self.backbone.forward_features(batch['image'])))
        return {'prediction': pred, 'target': batch.get('target')}


class FPN(nn.Module):
    """mmdet 3.0.0 necks.FPN with torchok's reversed in_channels (necks/detection/fpn.py:61-117); ConvModule defaults
    (no norm, no activation) => plain biased convs under `.conv`.  Parity unpinned by reference goldens (mmdet not
    vendored, no upstream test); cross-checked against torchvision.ops.FeaturePyramidNetwork in tests/test_oracle_models.py."""

    class _CM(nn.Module):
        def __init__(self, cin, cout, k, stride=1, padding=0):
            super().__init__()
            self.conv = nn.Conv2d(cin, cout, k, stride, padding)

        def forward(self, x):
            return conv(self.conv, x)

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False):
        super().__init__()
        self.in_channels = list(in_channels[::-1])
        n = len(self.in_channels)
        self.num_outs, self.start_level, self.relu_before_extra_convs = num_outs, start_level, relu_before_extra_convs
        self.backbone_end_level = n if end_level in (-1, n - 1) else end_level + 1
        self.add_extra_convs = 'on_input' if add_extra_convs is True else add_extra_convs
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(start_level, self.backbone_end_level):
            self.lateral_convs.append(FPN._CM(self.in_channels[i], out_channels, 1))
            self.fpn_convs.append(FPN._CM(out_channels, out_channels, 3, padding=1))
        extra = num_outs - self.backbone_end_level + start_level
        if self.add_extra_convs and extra >= 1:
            for i in range(extra):
                cin = self.in_channels[self.backbone_end_level - 1] if (i == 0 and self.add_extra_convs == 'on_input') \
                    else out_channels
                self.fpn_convs.append(FPN._CM(cin, out_channels, 3, 2, 1))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.normal_(m.bias, std=0.1)

    def forward(self, inputs):
        lat = [c(inputs[i + self.start_level]) for i, c in enumerate(self.lateral_convs)]
        used = len(lat)
        for i in range(used - 1, 0, -1):
            lat[i - 1] = q(lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest'))
        outs = [self.fpn_convs[i](lat[i]) for i in range(used)]
        if self.num_outs > len(outs):
            if not self.add_extra_convs:
                for _ in range(self.num_outs - used):
                    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
            else:
                src = {'on_input': inputs[self.backbone_end_level - 1], 'on_lateral': lat[-1], 'on_output': outs[-1]}[
                    self.add_extra_convs]
                outs.append(self.fpn_convs[used](src))
                for i in range(used + 1, self.num_outs):
                    outs.append(self.fpn_convs[i](F.relu(outs[-1]) if self.relu_before_extra_convs else outs[-1]))
        return tuple(outs)


# ------------------------------------------------------------------------------------------- N4: OCR head, UNet neck
# Restated ahead of the CUDA path (no kernels yet): pinned against the reference's own files executed by
# path (tests/golden/make_reference_goldens.py), so that the kernels have a checked target when they are written.
class OCRSegmentationHead(nn.Module):
    """torchok/models/heads/segmentation/ocr.py:133-192 with SpatialGather_Module (:24-46), ObjectAttentionBlock
    (:49-102) at scale 1 and SpatialOCR (:105-130).  State-dict keys follow the reference's module names."""

    class _Attention(nn.Module):
        def __init__(self, cin, key):
            super().__init__()
            self.key_channels = key
            two = lambda a, b: nn.Sequential(ConvBnAct(a, b, 1), ConvBnAct(b, b, 1))  # noqa: E731
            self.f_pixel, self.f_object, self.f_down = two(cin, key), two(cin, key), two(cin, key)
            self.f_up = ConvBnAct(key, cin, 1)

        def forward(self, x, proxy):
            b, _, h, w = x.shape
            query = self.f_pixel(x).view(b, self.key_channels, -1).permute(0, 2, 1)
            key = self.f_object(proxy).view(b, self.key_channels, -1)
            value = self.f_down(proxy).view(b, self.key_channels, -1).permute(0, 2, 1)
            sim = q(F.softmax(q((self.key_channels ** -.5) * torch.matmul(query, key)), dim=-1))
            ctx = q(torch.matmul(sim, value)).permute(0, 2, 1).contiguous().view(b, self.key_channels, h, w)
            return self.f_up(ctx)

    class _SpatialOCR(nn.Module):
        def __init__(self, cin, key, cout, dropout):
            super().__init__()
            self.object_context_block = OCRSegmentationHead._Attention(cin, key)
            self.conv_bn_dropout = nn.Sequential(ConvBnAct(2 * cin, cout, 1), nn.Dropout2d(dropout))

        def forward(self, feats, proxy):
            return self.conv_bn_dropout(torch.cat([self.object_context_block(feats, proxy), feats], 1))

    def __init__(self, in_channels, num_classes, do_interpolate=True, ocr_mid_channels=128, ocr_key_channels=64):
        super().__init__()
        self.num_classes, self.do_interpolate = num_classes, do_interpolate
        self.conv3x3_ocr = ConvBnAct(in_channels, ocr_mid_channels, 3, padding=1)
        self.ocr_distri_head = OCRSegmentationHead._SpatialOCR(ocr_mid_channels, ocr_key_channels, ocr_mid_channels, 0.05)
        self.last_reduction = ConvBnAct(ocr_mid_channels, ocr_mid_channels // 16, 1)
        self.aux_head = nn.Sequential(ConvBnAct(in_channels, in_channels, 1), nn.Conv2d(in_channels, num_classes, 1))
        self.classifier = nn.Conv2d(ocr_mid_channels // 16, num_classes, 1)

    def forward(self, x):
        image, feats = x
        out_aux = conv(self.aux_head[1], self.aux_head[0](feats))
        feats = self.conv3x3_ocr(feats)
        b, k = out_aux.shape[:2]
        probs = q(F.softmax(out_aux.view(b, k, -1), dim=2))                         # soft object regions, scale 1
        context = q(torch.matmul(probs, feats.view(b, feats.size(1), -1).permute(0, 2, 1)))
        context = context.permute(0, 2, 1).unsqueeze(3)                             # b x c x k x 1
        out = conv(self.classifier, self.last_reduction(self.ocr_distri_head(feats, context)))
        if self.do_interpolate:
            out = q(F.interpolate(out, size=image.shape[2:], mode='bilinear', align_corners=False))
            out_aux = q(F.interpolate(out_aux, size=image.shape[2:], mode='bilinear', align_corners=False))
        if self.num_classes == 1:
            out, out_aux = out[:, 0], out_aux[:, 0]
        return (out, out_aux) if self.training else out


class UnetNeck(nn.Module):
    """torchok/models/necks/segmentation/unet.py:77-131 (use_attention=False): optional centre block, then decoder
    blocks `nearest x2 -> cat(skip) -> ConvBnAct 3x3 -> ConvBnAct 3x3`, deepest feature first."""

    class _Block(nn.Module):
        def __init__(self, cin, skip, cout, use_batchnorm):
            super().__init__()
            self.conv1 = ConvBnAct(cin + skip, cout, 3, padding=1, use_batchnorm=use_batchnorm)
            self.conv2 = ConvBnAct(cout, cout, 3, padding=1, use_batchnorm=use_batchnorm)

        def forward(self, x, skip=None):
            x = F.interpolate(x, scale_factor=2, mode='nearest')
            if skip is not None:
                if skip.size(2) != x.size(2):
                    skip = F.interpolate(skip, size=x.shape[2:], mode='nearest')
                x = torch.cat([x, skip], dim=1)
            return self.conv2(self.conv1(x))

    def __init__(self, in_channels, decoder_channels=(512, 256, 128, 64, 64), use_batchnorm=True, center=True):
        super().__init__()
        enc = list(in_channels)[::-1]
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0] * (len(decoder_channels) - len(enc) + 1)
        self.out_channels = decoder_channels[-1]
        self.center = nn.Sequential(ConvBnAct(enc[0], enc[0], 3, padding=1, use_batchnorm=use_batchnorm),
                                    ConvBnAct(enc[0], enc[0], 3, padding=1, use_batchnorm=use_batchnorm)) \
            if center else nn.Identity()
        self.blocks = nn.ModuleList(UnetNeck._Block(i, s, o, use_batchnorm)
                                    for i, s, o in zip(ins, skips, decoder_channels))

    def forward(self, features):
        head, *skips, image = features[::-1]
        x = self.center(head)
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return [image, x]
