"""Plain-torch CPU restatement of Swin Transformer V2.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED by reference goldens (upstream tests are shape-only: test_backbone.py:161-182); the block arithmetic
and the whole network are cross-checked against torchvision's independent SwinTransformerBlockV2 / PatchMergingV2 /
SwinTransformer (V2) in tests/test_oracle_models.py (outputs and input gradient, copied weights).

Follows torchok/models/backbones/swin.py:71-275 (BasicLayer returning (downsampled, pre-downsample), feature_norms,
BCHW outputs) and timm 0.6.13 swin_transformer_v2: PatchEmbed conv4x4 s4 + LN; res-post-norm
blocks; WindowAttention with cosine similarity, logit_scale clamp ln(100), 16*sigmoid(cpb_mlp) bias, -100 shift mask;
PatchMerging 2x2 gather -> Linear(4C, 2C, bias=False) -> LN; window = min(res, window), shift = 0 when res <= window.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .models import linear as qlinear
from .models import q, qw


def window_partition(x, ws):
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def window_reverse(windows, ws, h, w):
    b = int(windows.shape[0] / (h * w / ws / ws))
    x = windows.view(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(b, h, w, -1)


class WindowAttention(nn.Module):
    def __init__(self, dim, ws, heads):
        super().__init__()
        self.dim, self.ws, self.heads = dim, ws, heads
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((heads, 1, 1))))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, heads, bias=False))
        rel = torch.arange(-(ws - 1), ws, dtype=torch.float32)
        table = torch.stack(torch.meshgrid([rel, rel], indexing='ij')).permute(1, 2, 0).contiguous().unsqueeze(0)
        table = table / max(ws - 1, 1) * 8
        table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
        self.register_buffer('relative_coords_table', table, persistent=False)
        coords = torch.flatten(torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij')), 1)
        r = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        r[:, :, 0] += ws - 1
        r[:, :, 1] += ws - 1
        r[:, :, 0] *= 2 * ws - 1
        self.register_buffer('relative_position_index', r.sum(-1), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.register_buffer('k_bias', torch.zeros(dim), persistent=False)
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, mask=None):
        b_, n, c = x.shape
        bias = torch.cat((self.q_bias, self.k_bias, self.v_bias))
        qkv = q(F.linear(x, qw(self.qkv.weight), bias))
        qkv = qkv.reshape(b_, n, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        qq, k, v = qkv.unbind(0)
        attn = F.normalize(qq, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        attn = attn * torch.clamp(self.logit_scale, max=math.log(1. / 0.01)).exp()
        t = self.cpb_mlp(self.relative_coords_table).view(-1, self.heads)
        t = t[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
        attn = attn + (16 * torch.sigmoid(t)).unsqueeze(0)
        if mask is not None:
            nw = mask.shape[0]
            attn = attn.view(b_ // nw, nw, self.heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.heads, n, n)
        attn = attn.softmax(dim=-1)
        x = q((attn @ v).transpose(1, 2).reshape(b_, n, c))
        return qlinear(self.proj, x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return qlinear(self.fc2, q(self.act(qlinear(self.fc1, x))))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, res, heads, window_size=7, shift_size=0, mlp_ratio=4.):
        super().__init__()
        self.res = res
        self.ws = min(res[0], window_size)
        self.shift = 0 if res[0] <= self.ws else shift_size
        self.attn = WindowAttention(dim, self.ws, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = nn.LayerNorm(dim)
        if self.shift > 0:
            h, w = res
            img = torch.zeros((1, h, w, 1))
            cnt = 0
            for hs in (slice(0, -self.ws), slice(-self.ws, -self.shift), slice(-self.shift, None)):
                for wsl in (slice(0, -self.ws), slice(-self.ws, -self.shift), slice(-self.shift, None)):
                    img[:, hs, wsl, :] = cnt
                    cnt += 1
            mw = window_partition(img, self.ws).view(-1, self.ws * self.ws)
            m = mw.unsqueeze(1) - mw.unsqueeze(2)
            self.register_buffer('attn_mask', m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0)))
        else:
            self.attn_mask = None

    def _attn(self, x):
        h, w = self.res
        b, l, c = x.shape
        x = x.view(b, h, w, c)
        if self.shift > 0:
            x = torch.roll(x, shifts=(-self.shift, -self.shift), dims=(1, 2))
        xw = window_partition(x, self.ws).view(-1, self.ws * self.ws, c)
        aw = self.attn(xw, mask=self.attn_mask).view(-1, self.ws, self.ws, c)
        x = window_reverse(aw, self.ws, h, w)
        if self.shift > 0:
            x = torch.roll(x, shifts=(self.shift, self.shift), dims=(1, 2))
        return x.view(b, l, c)

    def forward(self, x):
        x = q(x + self.norm1(self._attn(x)))
        return q(x + self.norm2(self.mlp(x)))


class PatchMerging(nn.Module):
    def __init__(self, res, dim):
        super().__init__()
        self.res, self.dim = res, dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(2 * dim)

    def forward(self, x):
        h, w = self.res
        b, l, c = x.shape
        x = x.view(b, h, w, c)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
        return q(self.norm(qlinear(self.reduction, x.view(b, -1, 4 * c))))


class BasicLayer(nn.Module):
    def __init__(self, dim, res, depth, heads, window_size, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([SwinTransformerBlock(dim, res, heads, window_size,
                                                          0 if i % 2 == 0 else window_size // 2) for i in range(depth)])
        self.downsample = PatchMerging(res, dim) if downsample else None

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return (self.downsample(x) if self.downsample is not None else x), x


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, cin, dim):
        super().__init__()
        self.grid_size = (img_size // patch, img_size // patch)
        self.proj = nn.Conv2d(cin, dim, patch, patch)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        y = q(F.conv2d(q(x), qw(self.proj.weight), self.proj.bias, stride=self.proj.stride))
        return q(self.norm(y.flatten(2).transpose(1, 2)))


class SwinTransformerV2(nn.Module):
    def __init__(self, img_size=256, in_channels=3, embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24),
                 window_size=7):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, 4, in_channels, embed_dim)
        gs = self.patch_embed.grid_size
        self.res = [(gs[0] // 2 ** i, gs[1] // 2 ** i) for i in range(len(depths))]
        self.chs = [embed_dim * 2 ** i for i in range(len(depths))]
        self.out_encoder_channels = tuple(self.chs)
        self.layers = nn.ModuleList([BasicLayer(self.chs[i], self.res[i], depths[i], num_heads[i], window_size,
                                                i < len(depths) - 1) for i in range(len(depths))])
        self.feature_norms = nn.ModuleList([nn.LayerNorm(c) for c in self.chs])
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _bchw(self, x, i):
        x = q(self.feature_norms[i](x))
        h, w = self.res[i]
        return x.view(-1, h, w, self.chs[i]).permute(0, 3, 1, 2).contiguous()

    def forward_features(self, x):
        feats = [x]
        t = self.patch_embed(x)
        for i, layer in enumerate(self.layers):
            t, a = layer(t)
            feats.append(self._bchw(a, i))
        return feats

    def forward(self, x):
        t = self.patch_embed(x)
        for layer in self.layers:
            t, _ = layer(t)
        return self._bchw(t, -1)


def dedegenerate_ln_(model, seed=0):
    """res-post-norm init zeroes norm1/norm2, so fresh blocks are identities; randomise every LayerNorm
    affine (and the attention biases / logit scales) so that parity tests exercise the blocks."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, m in model.named_modules():
            if isinstance(m, nn.LayerNorm):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if isinstance(m, WindowAttention) or type(m).__name__ == 'WindowAttention':
                m.q_bias.copy_(torch.randn(m.q_bias.shape, generator=g) * 0.2)
                m.v_bias.copy_(torch.randn(m.v_bias.shape, generator=g) * 0.2)
                m.logit_scale.add_(torch.randn(m.logit_scale.shape, generator=g) * 0.3)
    return model
