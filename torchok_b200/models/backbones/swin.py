"""Swin Transformer V2 backbones on the sm_100a kernels.

Mirror of torchok/models/backbones/swin.py:84-275 (SwinTransformerV2: patch embedding, four BasicLayers that return
`(downsampled, pre-downsample)`, per-stage `feature_norms`, BCHW outputs) with timm 0.6.13's
swin_transformer_v2 blocks restated (timm is not vendored): res-post-norm blocks
`x + drop_path(norm(f(x)))`, cosine window attention with a learned logit scale clamped at ln(100), continuous
relative position bias `16 * sigmoid(cpb_mlp(log-spaced coords))`, PatchMerging = 2x2 gather -> Linear(4C, 2C) ->
LayerNorm.  Parameter names follow timm (`layers.0.blocks.1.attn.cpb_mlp.0.weight`, `...attn.q_bias`,
`layers.0.downsample.reduction.weight`, `patch_embed.proj.weight`, `feature_norms.2.bias`).

Execution: tokens live as one (B*H*W, C) bf16 matrix (= the NHWC grid).  Linear layers run on the tcgen05 GEMM
(tok_linear_*), LayerNorm + residual (+ stochastic depth scale) is one pass (tok_layernorm_*), GELU one pass, and the
whole attention core of a block — normalise q/k, logits, bias, shift mask, softmax, PV, with the cyclic shift and the
window partition folded into its addressing — is one kernel (tok_window_attn_*: forward on tcgen05 with S and O in
TMEM, backward on CUDA cores; window <= 8).
The (2 ws - 1)^2 x 2 -> 512 -> heads cpb MLP, the relative-position gather and the 16 * sigmoid that produce the bias table
are one small kernel chain (tok_cpb_bias_*), forward and backward.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels as K
from ...constructor import BACKBONES
from ..base import BaseBackbone
from ..modules.layers import Conv2d, ConvFn


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class Linear(nn.Linear):
    """nn.Linear parameters executed by tok_linear_* on (rows, in_features) bf16 matrices."""

    def forward(self, x, bias_grad_external=False, res_link=None):
        return K.linear(x, self.weight, self.bias, bias_grad_external, res_link)


class LayerNorm(nn.LayerNorm):
    def forward(self, x, residual=None, rowscale=None, rows_per_sample=1, colsum_param=None, res_link=None):
        return K.layernorm(x, self.weight, self.bias, self.eps, residual, rowscale, rows_per_sample, colsum_param,
                           res_link)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = Linear(hidden_features, in_features)

    def forward(self, x, fc2_bias_external=False, res_link=None):
        # the bias gradients of fc1 / fc2 are the column sums of the GELU / LayerNorm input gradients: those backward
        # kernels accumulate them on the way instead of a separate pass over dy per linear layer
        # ... and the GELU backward itself runs in the epilogue of fc2's data-gradient GEMM (K.gelu_linear)
        hid = self.fc1.out_features
        ext1 = self.fc1.bias is not None and (K.gelu_dgrad_fused(hid, self.fc2.out_features) or K.gelu_fuses_colsum(hid))
        h = self.fc1(x, ext1, res_link)
        return K.gelu_linear(h, self.fc2.weight, self.fc2.bias, self.fc1.bias if ext1 else None, fc2_bias_external)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, pretrained_window_size=(0, 0)):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.pretrained_window_size = pretrained_window_size
        if window_size[0] != window_size[1]:
            raise NotImplementedError('square windows only')
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True),
                                     nn.Linear(512, num_heads, bias=False))
        ws = window_size[0]
        rel_h = torch.arange(-(ws - 1), ws, dtype=torch.float32)
        table = torch.stack(torch.meshgrid([rel_h, rel_h], indexing='ij')).permute(1, 2, 0).contiguous().unsqueeze(0)
        div = (pretrained_window_size[0] - 1) if pretrained_window_size[0] > 0 else (ws - 1)
        table = table / max(div, 1)
        table = table * 8
        table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
        self.register_buffer('relative_coords_table', table, persistent=False)
        coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij'))
        flat = torch.flatten(coords, 1)
        rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer('relative_position_index', rel.sum(-1), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.register_buffer('k_bias', torch.zeros(dim), persistent=False)
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = self.k_bias = self.v_bias = None
        self.proj = Linear(dim, dim)

    def bias_table(self):
        # cpb MLP + relative_position_index gather + 16 * sigmoid as one kernel chain (tok_cpb_bias_*); the gradients of
        # the three cpb parameters are accumulated by its backward
        mlp = self.cpb_mlp
        w1, b1, w2 = mlp[0].weight, mlp[0].bias, mlp[2].weight
        if w1.is_cuda and w1.is_contiguous() and w2.is_contiguous() and w2.shape[1] == 512:
            return K.cpb_bias(self.relative_coords_table, w1, b1, w2, self.window_size[0])
        n = self.window_size[0] * self.window_size[1]
        t = mlp(self.relative_coords_table).view(-1, self.num_heads)
        t = t[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
        return 16 * torch.sigmoid(t)

    def forward(self, x, geom, proj_bias_external=False, res_link=None):
        # q/v bias gradients ride in the tcgen05 attention backward (windows up to 8x8); the large-window kernel leaves
        # them to the qkv linear layer
        ext = (self.q_bias is not None and K.attn_fuses_qv_bias_grad()
               and self.window_size[0] * self.window_size[1] <= 64)
        qkv = K.qkv_linear(x, self.qkv.weight, self.q_bias, self.v_bias, ext, res_link)
        out = K.window_attention(qkv, self.bias_table(), self.logit_scale, geom,
                                 (self.q_bias, self.v_bias) if ext else None)
        return self.proj(out, proj_bias_external)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 drop_path=0., pretrained_window_size=0):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        ws = [min(r, w) for r, w in zip(input_resolution, to_2tuple(window_size))]
        ss = [0 if r <= w else s for r, w, s in zip(input_resolution, ws, to_2tuple(shift_size))]
        self.window_size, self.shift_size = tuple(ws), tuple(ss)
        self.attn = WindowAttention(dim, self.window_size, num_heads, qkv_bias, to_2tuple(pretrained_window_size))
        self.norm1 = LayerNorm(dim)
        self.drop_path_rate = float(drop_path)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = LayerNorm(dim)
        if any(self.shift_size):
            self.register_buffer('attn_mask', self._make_mask())  # kept for state-dict parity; the kernel derives it
        else:
            self.attn_mask = None

    def _make_mask(self):
        h, w = self.input_resolution
        img = torch.zeros((1, h, w, 1))
        cnt = 0
        for hs in (slice(0, -self.window_size[0]), slice(-self.window_size[0], -self.shift_size[0]),
                   slice(-self.shift_size[0], None)):
            for wsl in (slice(0, -self.window_size[1]), slice(-self.window_size[1], -self.shift_size[1]),
                        slice(-self.shift_size[1], None)):
                img[:, hs, wsl, :] = cnt
                cnt += 1
        ws = self.window_size
        win = img.view(1, h // ws[0], ws[0], w // ws[1], ws[1], 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws[0] * ws[1])
        mask = win.unsqueeze(1) - win.unsqueeze(2)
        return mask.masked_fill(mask != 0, float(-100.0)).masked_fill(mask == 0, float(0.0))

    def _rowscale(self, batch, device):
        if self.drop_path_rate == 0. or not self.training:
            return None
        keep = 1 - self.drop_path_rate
        return torch.empty(batch, device=device).bernoulli_(keep).div_(keep)

    def forward(self, x, batch):
        h, w = self.input_resolution
        geom = (batch, h, w, self.dim, self.num_heads, self.window_size[0], self.shift_size[0])
        rps = h * w
        ext = K.layernorm_fuses_colsum(self.dim)   # proj / fc2 bias gradients come out of the LayerNorm backward
        # skip-connection gradients join the branch gradients inside the qkv / fc1 dgrad GEMMs (K.ResidualLink); only
        # when x itself needs a gradient, otherwise the branch's dgrad (and with it the link) would never run
        link1 = K.ResidualLink() if x.requires_grad else None
        x = self.norm1(self.attn(x, geom, ext, link1), residual=x, rowscale=self._rowscale(batch, x.device),
                       rows_per_sample=rps, colsum_param=self.attn.proj.bias if ext else None, res_link=link1)
        link2 = K.ResidualLink() if x.requires_grad else None
        x = self.norm2(self.mlp(x, ext, link2), residual=x, rowscale=self._rowscale(batch, x.device), rows_per_sample=rps,
                       colsum_param=self.mlp.fc2.bias if ext else None, res_link=link2)
        return x


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = LayerNorm(2 * dim)

    def forward(self, x, batch):
        h, w = self.input_resolution
        x = K.patch_merge(x.reshape(-1, self.dim), batch, h, w)   # the 2x2 gather of timm's PatchMerging, one kernel
        return self.norm(K.linear(x, self.reduction.weight, None))


class BasicLayer(nn.Module):
    """timm BasicLayer as adapted by torchok/models/backbones/swin.py:71-81: returns (downsampled, pre-downsample)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True,
                 drop_path=0., downsample=None, pretrained_window_size=0):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.grad_checkpointing = False
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size,
                                 shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio,
                                 qkv_bias=qkv_bias, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 pretrained_window_size=pretrained_window_size) for i in range(depth)])
        self.downsample = downsample(input_resolution, dim) if downsample is not None else None

    def forward(self, x, batch):
        for blk in self.blocks:
            x = blk(x, batch)
        return (self.downsample(x, batch) if self.downsample is not None else x), x

    def _init_respostnorm(self):
        for blk in self.blocks:
            nn.init.constant_(blk.norm1.bias, 0)
            nn.init.constant_(blk.norm1.weight, 0)
            nn.init.constant_(blk.norm2.bias, 0)
            nn.init.constant_(blk.norm2.weight, 0)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = LayerNorm(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        b, c, h, w = x.shape
        if (h, w) != self.img_size:
            raise AssertionError(f"Input image size ({h}*{w}) doesn't match model ({self.img_size[0]}*{self.img_size[1]}).")
        ph, pw = self.patch_size
        if ph == pw and self.proj.bias is not None and K.patch_embed_supported(c, ph, h, w, self.proj.out_channels):
            tokens = K.patch_embed(x, self.proj.weight, self.proj.bias)   # direct kernel on the NCHW fp32 image
        else:
            y = ConvFn.apply(x, self.proj, False, self.proj.weight, self.proj.bias)   # (B, C, H/4, W/4), NHWC memory
            tokens = y.permute(0, 2, 3, 1).reshape(-1, y.shape[1])
        return self.norm(tokens) if not isinstance(self.norm, nn.Identity) else tokens


class SwinTransformerV2(BaseBackbone):
    def __init__(self, img_size=256, patch_size=4, in_channels=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4., qkv_bias=True, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 pretrained_window_sizes=(0, 0, 0, 0), load_attn_mask=True):
        super().__init__(in_channels=in_channels)
        if drop_rate != 0. or attn_drop_rate != 0.:
            raise NotImplementedError('SwinTransformerV2: dropout inside attention / MLP is outside the hot-path scope')
        if embed_dim // num_heads[0] != 32:
            raise NotImplementedError('the window-attention kernel is specialised for head_dim 32')
        self.img_size = to_2tuple(img_size)
        self.num_layers = len(depths)
        self.embed_dim, self.ape, self.patch_norm = embed_dim, ape, patch_norm
        self.encoder_channels = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        self._out_channels = self.encoder_channels[-1]
        self._out_encoder_channels = self.encoder_channels
        self.load_attn_mask = load_attn_mask
        self.patch_embed = PatchEmbed(img_size, patch_size, in_channels, embed_dim, norm_layer if patch_norm else None)
        gs = self.patch_embed.grid_size
        self.input_resolutions = [(gs[0] // (2 ** i), gs[1] // (2 ** i)) for i in range(self.num_layers)]
        self.patches_resolution = gs
        if ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=.02)
        else:
            self.absolute_pos_embed = None
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i), input_resolution=self.input_resolutions[i], depth=depths[i],
                num_heads=num_heads[i], window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                downsample=PatchMerging if (i < self.num_layers - 1) else None,
                pretrained_window_size=pretrained_window_sizes[i]))
        self.feature_norms = nn.ModuleList([LayerNorm(c) for c in self.encoder_channels])
        self.init_weights()

    @torch.no_grad()
    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for bly in self.layers:
            bly._init_respostnorm()

    def no_weight_decay(self):
        nod = ['absolute_pos_embed']
        for n, _ in self.named_modules():
            if any(kw in n for kw in ('cpb_mlp', 'logit_scale', 'relative_position_bias_table')):
                nod.append(n)
        return nod

    def _forward_patch_emb(self, x):
        x = self.patch_embed(x)
        if self.ape:
            b = x.shape[0] // self.patch_embed.num_patches
            x = (x.view(b, -1, self.embed_dim) + self.absolute_pos_embed.to(x.dtype)).view(-1, self.embed_dim)
        return x

    def _normalize_with_bhwc_reshape(self, x, layer_number, batch, normalize=True):
        if normalize:
            x = self.feature_norms[layer_number](x)
        h, w = self.input_resolutions[layer_number]
        c = self.encoder_channels[layer_number]
        return x.view(batch, h, w, c).permute(0, 3, 1, 2)   # logical BCHW on NHWC memory

    def _forward_collect(self, x):
        batch = x.shape[0]
        tokens = self._forward_patch_emb(x)
        feats = []
        for i, layer in enumerate(self.layers):
            tokens, attn = layer(tokens, batch)
            feats.append((attn, i))
        return feats, batch

    def forward_features(self, x):
        feats, batch = self._forward_collect(x)
        return [x] + [self._normalize_with_bhwc_reshape(a, i, batch) for a, i in feats]

    def forward(self, x):
        batch = x.shape[0]
        tokens = self._forward_patch_emb(x)
        for layer in self.layers:
            tokens, _ = layer(tokens, batch)
        return self._normalize_with_bhwc_reshape(tokens, -1, batch)

    def load_state_dict(self, state_dict, strict=True):
        if not self.load_attn_mask:
            state_dict = {k: v for k, v in state_dict.items() if 'attn_mask' not in k}
        return super().load_state_dict(state_dict, strict)

    def get_stages(self, stage):
        return nn.ModuleList([self.patch_embed, self.pos_drop] + list(self.layers[:stage]))


def _create(variant, pretrained=False, **kwargs):
    for k in ('num_classes', 'global_pool', 'in_chans'):
        kwargs.pop(k, None)
    model = SwinTransformerV2(**kwargs)
    if pretrained:
        from ...constructor.load import load_pretrained
        load_pretrained(model, variant)
    return model


def _register(name, **fixed):
    def factory(pretrained=False, **kwargs):
        return _create(name, pretrained, **dict(fixed, **kwargs))
    factory.__name__ = factory.__qualname__ = name
    factory.__module__ = __name__
    factory.__doc__ = f'Swin-V2 {name} {fixed}'
    globals()[name] = factory
    return BACKBONES.register_class(factory)


# swin.py:285-405 of the reference (windows up to 8x8: tcgen05 attention kernels; window 12 / 16 / 24 variants: the
# large-window CUDA-core attention kernels, csrc/tok_swin.cu::window_attn_big_*)
_register('swinv2_custom')
_T, _S, _B = dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24)), \
    dict(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24)), \
    dict(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32))
_register('swinv2_tiny_window16_256', window_size=16, **_T)
_register('swinv2_tiny_window8_256', window_size=8, **_T)
_register('swinv2_small_window16_256', window_size=16, **_S)
_register('swinv2_small_window8_256', window_size=8, **_S)
_register('swinv2_base_window16_256', window_size=16, **_B)
_register('swinv2_base_window8_256', window_size=8, **_B)
