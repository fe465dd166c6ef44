#!/usr/bin/env python
"""Golden vectors for the in-tree WIRING of the three backbones, produced by the REFERENCE's own files executed by path:

    torchok/models/backbones/resnet.py   (make_blocks :363-405, ResNet :408-563, resnet18 / resnet50 factories)
    torchok/models/backbones/hrnet.py    (transition / layer / stage builders :111-192, forward_stages :194-210)
    torchok/models/backbones/swin.py     (BasicLayer :71-81, SwinTransformerV2 :84-275, swinv2_* factories)
    torchok/models/backbones/base_backbone.py, torchok/models/base.py

timm is not installed in this image, so the names those files import from it are provided by stand-ins:

    timm.models.resnet.{BasicBlock,Bottleneck}   torchvision's independent blocks behind timm's constructor signature
                                                  (same state-dict keys, same v1.5 arithmetic) + zero_init_last()
    timm.models.resnet.downsample_conv            Conv2d(1x1, stride) + norm — timm 0.6.13 semantics for the plain variants
    timm.models.hrnet.HighResolutionModule        oracle/models.py::HighResolutionModule behind timm's signature, built
                                                  from the torchvision blocks above; cfg_cls = timm's published tables
    timm.models.swin_transformer_v2.*             oracle/swin.py blocks behind timm's signatures (those blocks are
                                                  cross-checked against torchvision's SwinTransformerBlockV2 elsewhere)
    timm.models.features.FeatureHooks, helpers.build_model_with_cfg, layers.*   minimal restatements

What this pins is everything the reference itself contributes to these networks: stem, stage / transition / downsample
construction, strides and channel arithmetic, the stage loop, feature hooks, forward_features ordering, feature norms and
the BCHW reshape, state-dict key names, out_channels / out_encoder_channels, get_stages().  The parameters are not stored:
both sides rebuild them with `seeded_state` (make_n4_goldens.py).  Replayed by tests/test_oracle_backbone_goldens.py
(oracle networks, CPU, 1e-5) and tests/test_reference_goldens_gpu.py (kernels, bf16 bar).

    python tests/golden/make_backbone_goldens.py      # needs /root/reference; writes tests/golden/backbone_goldens.pt
"""
import collections
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_reference_goldens import bf, install_stub_tree, load  # noqa: E402

OUT = os.path.join(HERE, 'backbone_goldens.pt')

# timm 0.6.13 hrnet.cfg_cls rows used here (published architecture tables of the HRNet paper)
def _stage(mods, branches, block, blocks, channels):
    return dict(NUM_MODULES=mods, NUM_BRANCHES=branches, BLOCK=block, NUM_BLOCKS=tuple(blocks),
                NUM_CHANNELS=tuple(channels), FUSE_METHOD='SUM')


def _hr(c, s1_blocks=4, s1_ch=64, blocks=4, mods=(1, 4, 3)):
    return dict(STEM_WIDTH=64, STAGE1=_stage(1, 1, 'BOTTLENECK', (s1_blocks,), (s1_ch,)),
                STAGE2=_stage(mods[0], 2, 'BASIC', (blocks,) * 2, (c, 2 * c)),
                STAGE3=_stage(mods[1], 3, 'BASIC', (blocks,) * 3, (c, 2 * c, 4 * c)),
                STAGE4=_stage(mods[2], 4, 'BASIC', (blocks,) * 4, (c, 2 * c, 4 * c, 8 * c)))


CFG_CLS = dict(hrnet_w18_small=_hr(16, 1, 32, 2, (1, 1, 1)), hrnet_w18_small_v2=_hr(18, 2, 64, 2, (1, 3, 2)),
               hrnet_w18=_hr(18), hrnet_w30=_hr(30), hrnet_w32=_hr(32), hrnet_w40=_hr(40), hrnet_w44=_hr(44),
               hrnet_w48=_hr(48), hrnet_w64=_hr(64))


def seeded_state(module, seed):
    """Same rule as make_n4_goldens.seeded_state, plus: derived integer / mask buffers are left alone and 1-D
    'logit_scale'-like multi-dim parameters stay small."""
    gen = torch.Generator().manual_seed(seed)
    state = {}
    for k, v in module.state_dict().items():
        if k.endswith('num_batches_tracked') or 'attn_mask' in k or 'relative_' in k or not v.is_floating_point():
            state[k] = v.clone()
        elif k.endswith('running_var'):
            state[k] = torch.rand(v.shape, generator=gen) + 0.5
        elif k.endswith('running_mean'):
            state[k] = torch.randn(v.shape, generator=gen) * 0.1
        elif k.endswith('logit_scale'):
            state[k] = v.clone() + torch.randn(v.shape, generator=gen) * 0.3
        elif v.dim() > 1:
            state[k] = bf(torch.randn(v.shape, generator=gen) / (v[0].numel() ** 0.5))
        elif k.endswith('weight'):
            state[k] = torch.rand(v.shape, generator=gen) + 0.5
        else:
            state[k] = torch.randn(v.shape, generator=gen) * 0.1
    return state


def install_timm_stubs():
    import torchvision.models.resnet as tvr
    from oracle import models as om
    from oracle import swin as osw

    def mod(name, pkg=False):
        m = types.ModuleType(name)
        if pkg:
            m.__path__ = []
        sys.modules[name] = m
        return m

    timm = mod('timm', True)
    data = mod('timm.data')
    data.IMAGENET_DEFAULT_MEAN, data.IMAGENET_DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    models = mod('timm.models', True)
    timm.data, timm.models = data, models

    helpers = mod('timm.models.helpers')

    def build_model_with_cfg(model_cls, variant, pretrained, model_cfg=None, pretrained_strict=True,
                             pretrained_filter_fn=None, kwargs_filter=None, **kwargs):
        assert not pretrained
        for k in (kwargs_filter or ()):
            kwargs.pop(k, None)
        return model_cls(**kwargs) if model_cfg is None else model_cls(cfg=model_cfg, **kwargs)
    helpers.build_model_with_cfg = build_model_with_cfg

    layers = mod('timm.models.layers')
    layers.BlurPool2d = layers.GroupNorm = type('Unused', (), {})
    layers.DropPath = nn.Identity
    layers.get_attn = lambda *a, **k: None
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    feats = mod('timm.models.features')

    class FeatureHooks:   # timm 0.6.13 features.py: forward hooks on the named modules, outputs handed out once
        def __init__(self, hooks, named_modules, out_map=None, default_hook_type='forward'):
            modules = {k: v for k, v in named_modules}
            self._out = collections.defaultdict(collections.OrderedDict)
            for i, h in enumerate(hooks):
                name = h['module']
                assert h.get('hook_type', default_hook_type) == 'forward'
                modules[name].register_forward_hook(lambda m, inp, out, key=(out_map[i] if out_map else name):
                                                    self._out[out.device].__setitem__(key, out))

        def get_output(self, device):
            out = self._out[device]
            self._out[device] = collections.OrderedDict()
            return out
    feats.FeatureHooks = FeatureHooks

    res = mod('timm.models.resnet')

    class BasicBlock(tvr.BasicBlock):
        def __init__(self, inplanes, planes, stride=1, downsample=None, cardinality=1, base_width=64, reduce_first=1,
                     dilation=1, first_dilation=None, act_layer=nn.ReLU, norm_layer=nn.BatchNorm2d, attn_layer=None,
                     aa_layer=None, drop_block=None, drop_path=None):
            assert cardinality == 1 and base_width == 64 and reduce_first == 1 and dilation == 1
            assert attn_layer is None and aa_layer is None and drop_block is None and drop_path is None
            super().__init__(inplanes, planes, stride, downsample, norm_layer=norm_layer or nn.BatchNorm2d)

        def zero_init_last(self):
            nn.init.zeros_(self.bn2.weight)

    class Bottleneck(tvr.Bottleneck):
        def __init__(self, inplanes, planes, stride=1, downsample=None, cardinality=1, base_width=64, reduce_first=1,
                     dilation=1, first_dilation=None, act_layer=nn.ReLU, norm_layer=nn.BatchNorm2d, attn_layer=None,
                     aa_layer=None, drop_block=None, drop_path=None):
            assert cardinality == 1 and reduce_first == 1 and dilation == 1
            assert attn_layer is None and aa_layer is None and drop_block is None and drop_path is None
            super().__init__(inplanes, planes, stride, downsample, 1, base_width, 1, norm_layer or nn.BatchNorm2d)

        def zero_init_last(self):
            nn.init.zeros_(self.bn3.weight)

    def downsample_conv(in_channels, out_channels, kernel_size, stride=1, dilation=1, first_dilation=None,
                        norm_layer=None):
        assert kernel_size == 1 and dilation == 1
        return nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, stride=stride, padding=0, bias=False),
                             (norm_layer or nn.BatchNorm2d)(out_channels))

    res.BasicBlock, res.Bottleneck, res.downsample_conv = BasicBlock, Bottleneck, downsample_conv
    res.downsample_avg = res.create_aa = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    res.drop_blocks = lambda rate=0.: [None, None, None, None]

    hr = mod('timm.models.hrnet')
    hr._BN_MOMENTUM = 0.1
    hr.BasicBlock, hr.Bottleneck = BasicBlock, Bottleneck
    hr.blocks_dict = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}
    hr.cfg_cls = CFG_CLS

    class HighResolutionModule(om.HighResolutionModule):
        def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                     multi_scale_output=True):
            assert fuse_method == 'SUM'
            super().__init__(num_branches, blocks, num_blocks, num_inchannels, num_channels, multi_scale_output)

        def get_num_in_chs(self):
            return self.num_inchannels
    hr.HighResolutionModule = HighResolutionModule

    sw = mod('timm.models.swin_transformer_v2')

    class PatchEmbed(osw.PatchEmbed):
        def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
            assert norm_layer is nn.LayerNorm
            super().__init__(img_size, patch_size, in_chans, embed_dim)
            self.num_patches = self.grid_size[0] * self.grid_size[1]

    class PatchMerging(osw.PatchMerging):
        def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
            super().__init__(input_resolution, dim)

    class BasicLayer(nn.Module):   # timm 0.6.13 BasicLayer.__init__; forward is overridden by the reference (swin.py:75-81)
        def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, drop=0.,
                     attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, pretrained_window_size=0):
            super().__init__()
            assert qkv_bias and drop == 0. and attn_drop == 0. and pretrained_window_size == 0
            self.grad_checkpointing = False
            self.blocks = nn.ModuleList([osw.SwinTransformerBlock(dim, input_resolution, num_heads, window_size,
                                                                  0 if i % 2 == 0 else window_size // 2, mlp_ratio)
                                         for i in range(depth)])
            self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None \
                else nn.Identity()

        def _init_respostnorm(self):
            for blk in self.blocks:
                for n in (blk.norm1, blk.norm2):
                    nn.init.constant_(n.bias, 0)
                    nn.init.constant_(n.weight, 0)
    sw.BasicLayer, sw.PatchEmbed, sw.PatchMerging = BasicLayer, PatchEmbed, PatchMerging
    sw.checkpoint_filter_fn = lambda sd, model: sd


def run_case(model, seed, x, train):
    state = seeded_state(model, seed)
    model.load_state_dict(state)
    model.train(train)
    x = x.clone().requires_grad_(True)
    feats = model.forward_features(x)
    gen = torch.Generator().manual_seed(seed + 1)
    rs = [bf(torch.randn(f.shape, generator=gen)) for f in feats[1:]]
    sum((f * r).sum() for f, r in zip(feats[1:], rs)).backward()
    last = model(x.detach())
    last = last if isinstance(last, (list, tuple)) else [last]
    # every parameter gradient as (norm, sum) — full tensors only for the small ones
    gnorm, keep = {}, {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            gnorm[n] = (float(p.grad.norm()), float(p.grad.double().sum()))
            if p.grad.numel() <= 2048:
                keep[n] = p.grad.detach().clone()
    names = list(keep)
    keep = {n: keep[n] for n in names[::max(1, len(names) // 12)]}
    return dict(keys=list(state.keys()), feats=[f.detach().clone() for f in feats[1:]], dx=x.grad.clone(),
                forward=[t.detach().clone() for t in last], grads=keep, gnorm=gnorm,
                running={k: v.clone() for k, v in model.state_dict().items() if 'running_' in k and train},
                out_channels=model.out_channels, out_encoder_channels=tuple(model.out_encoder_channels),
                stages=[len(model.get_stages(i)) for i in range(5)])


def main():
    install_stub_tree()
    install_timm_stubs()
    pkg = types.ModuleType('torchok.models.backbones')
    pkg.__path__ = []
    sys.modules['torchok.models.backbones'] = pkg
    load('torchok.models.base')
    bb = load('torchok.models.backbones.base_backbone')
    pkg.BaseBackbone, pkg.BackboneWrapper = bb.BaseBackbone, bb.BackboneWrapper
    resnet = load('torchok.models.backbones.resnet')
    hrnet = load('torchok.models.backbones.hrnet')
    swin = load('torchok.models.backbones.swin')
    g = torch.Generator().manual_seed(77)
    out = {}
    cases = [
        ('resnet18', lambda: resnet.resnet18(pretrained=False, in_channels=3), {}, (2, 3, 64, 64), True),
        ('resnet18', lambda: resnet.resnet18(pretrained=False, in_channels=3), {}, (2, 3, 32, 32), False),
        ('resnet50', lambda: resnet.resnet50(pretrained=False, in_channels=3), {}, (2, 3, 32, 32), True),
        ('hrnet_w18_small', lambda: hrnet.hrnet_w18_small(pretrained=False, in_channels=3), {}, (2, 3, 64, 64), True),
        ('hrnet_w18', lambda: hrnet.hrnet_w18(pretrained=False, in_channels=3), {}, (2, 3, 64, 64), False),
        ('swinv2_custom', lambda: swin.swinv2_custom(pretrained=False, img_size=64, window_size=4, depths=(2, 2, 2, 2)),
         dict(img_size=64, window_size=4, depths=(2, 2, 2, 2)), (2, 3, 64, 64), True),
        ('swinv2_tiny_window8_256', lambda: swin.swinv2_tiny_window8_256(pretrained=False, img_size=64),
         dict(img_size=64), (2, 3, 64, 64), False),
    ]
    for i, (name, make, kwargs, shape, train) in enumerate(cases):
        torch.manual_seed(i)
        m = make()
        x = bf(torch.randn(shape, generator=g))
        c = run_case(m, 1000 + i, x, train)
        c.update(name=name, kwargs=kwargs, x=x, train=train, seed=1000 + i)
        out.setdefault(name, []).append(c)
        print(name, 'train' if train else 'eval', [tuple(f.shape) for f in c['feats']], len(c['keys']), 'keys')
    torch.save(out, OUT)
    print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB')


if __name__ == '__main__':
    main()
