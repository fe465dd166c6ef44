"""The model-path oracle (oracle/models.py) against independent implementations and the reference's shape contracts.

The reference's own model tests are shape-only (tests/additional_tests/models/backbones/test_backbone.py:140-158) and
the arithmetic lives in timm 0.6.13 (not importable here) => "parity unpinned" by reference goldens; the restatement is
pinned instead against torchvision's independent ResNet (identical state_dict keys: bit-identical outputs)."""
import pytest
import torch
import torchvision

from oracle import models as om


@pytest.mark.parametrize('name', ['resnet18', 'resnet50'])
def test_oracle_resnet_is_bit_identical_to_torchvision(name):
    torch.manual_seed(0)
    o = om.resnet(name)
    om.dedegenerate_(o, 0)
    tv = getattr(torchvision.models, name)()
    missing = tv.load_state_dict(o.state_dict(), strict=False)
    assert set(missing.missing_keys) <= {'fc.weight', 'fc.bias'} and not missing.unexpected_keys
    tv.fc = torch.nn.Identity()
    tv.avgpool = torch.nn.Identity()
    x = torch.randn(2, 3, 64, 64)
    for mode in ('eval', 'train'):
        getattr(o, mode)()
        getattr(tv, mode)()
        with torch.no_grad():
            a = o(x)
            b = tv(x).reshape(a.shape)
        assert torch.equal(a, b), mode


def test_forward_features_shape_contract():
    """test_backbone.py:149-154 of the reference: resnet18 @64x64 -> 6 tensors with these shapes."""
    o = om.resnet('resnet18')
    feats = o.forward_features(torch.randn(2, 3, 64, 64))
    shapes = [tuple(f.shape) for f in feats]
    assert shapes == [(2, 3, 64, 64), (2, 64, 32, 32), (2, 64, 16, 16), (2, 128, 8, 8), (2, 256, 4, 4), (2, 512, 2, 2)]


def test_amp_mode_rounds_storage_only():
    torch.manual_seed(1)
    o = om.resnet('resnet18')
    om.dedegenerate_(o, 1)
    o.eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        a = o(x)
        with om.amp_bf16():
            b = o(x)
    rel = ((a - b).abs().max() / a.abs().max()).item()
    assert 0 < rel < 5e-2


def test_oracle_hrnet_shape_contracts():
    """Reference shape tests: hrnet_w18_small @64 (test_backbone.py:76-88), hrnet_w18 necks @224 (test_hrnet.py:15-26),
    segmentation head -> (B, classes, H, W)."""
    o = om.hrnet('hrnet_w18_small').eval()
    with torch.no_grad():
        feats = o.forward_features(torch.randn(2, 3, 64, 64))
    assert [tuple(f.shape) for f in feats] == [(2, 3, 64, 64), (2, 16, 16, 16), (2, 32, 8, 8), (2, 64, 4, 4), (2, 128, 2, 2)]
    neck = om.HRNetSegmentationNeck(o.out_encoder_channels).eval()
    head = om.SegmentationHead(neck.out_channels, 10).eval()
    with torch.no_grad():
        assert tuple(head(neck(feats)).shape) == (2, 10, 64, 64)
    cls = om.HRNetClassificationNeck(o.out_encoder_channels).eval()
    with torch.no_grad():
        assert tuple(cls(feats[1:]).shape) == (2, 2048, 2, 2)


def test_oracle_swinv2_block_matches_torchvision():
    """Independent cross-check of the Swin-V2 restatement (oracle/swin.py) against torchvision's
    SwinTransformerBlockV2 (same Microsoft lineage, different code): shifted and unshifted blocks, 8x8 grid, window 4."""
    from torchvision.models.swin_transformer import PatchMergingV2, SwinTransformerBlockV2

    from oracle import swin as osw
    torch.manual_seed(0)
    for shift in (0, 2):
        o = osw.SwinTransformerBlock(64, (8, 8), 2, window_size=4, shift_size=shift)
        osw.dedegenerate_ln_(o, 1)
        tv = SwinTransformerBlockV2(64, 2, [4, 4], [shift, shift], mlp_ratio=4.0, stochastic_depth_prob=0.0)
        sd = o.state_dict()
        with torch.no_grad():
            tv.norm1.load_state_dict({'weight': sd['norm1.weight'], 'bias': sd['norm1.bias']})
            tv.norm2.load_state_dict({'weight': sd['norm2.weight'], 'bias': sd['norm2.bias']})
            tv.attn.qkv.weight.copy_(sd['attn.qkv.weight'])
            tv.attn.qkv.bias.copy_(torch.cat([sd['attn.q_bias'], torch.zeros(64), sd['attn.v_bias']]))
            tv.attn.proj.weight.copy_(sd['attn.proj.weight'])
            tv.attn.proj.bias.copy_(sd['attn.proj.bias'])
            tv.attn.logit_scale.copy_(sd['attn.logit_scale'])
            tv.attn.cpb_mlp[0].weight.copy_(sd['attn.cpb_mlp.0.weight'])
            tv.attn.cpb_mlp[0].bias.copy_(sd['attn.cpb_mlp.0.bias'])
            tv.attn.cpb_mlp[2].weight.copy_(sd['attn.cpb_mlp.2.weight'])
            tv.mlp[0].weight.copy_(sd['mlp.fc1.weight'])
            tv.mlp[0].bias.copy_(sd['mlp.fc1.bias'])
            tv.mlp[3].weight.copy_(sd['mlp.fc2.weight'])
            tv.mlp[3].bias.copy_(sd['mlp.fc2.bias'])
        x = torch.randn(2, 8, 8, 64)
        o.eval(), tv.eval()
        with torch.no_grad():
            a = o(x.view(2, 64, 64)).view(2, 8, 8, 64)
            b = tv(x)
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), (shift, (a - b).abs().max())
    pm = osw.PatchMerging((8, 8), 64)
    tvm = PatchMergingV2(64)
    with torch.no_grad():
        tvm.reduction.weight.copy_(pm.reduction.weight)
        tvm.norm.weight.copy_(pm.norm.weight)
        tvm.norm.bias.copy_(pm.norm.bias)
        x = torch.randn(2, 8, 8, 64)
        assert torch.allclose(pm(x.view(2, 64, 64)).view(2, 4, 4, 128), tvm(x), atol=1e-5)


def test_oracle_swinv2_shape_contract():
    from oracle import swin as osw
    o = osw.SwinTransformerV2(img_size=64, window_size=8, depths=(1, 1, 1, 1)).eval()
    with torch.no_grad():
        feats = o.forward_features(torch.randn(2, 3, 64, 64))
    assert [tuple(f.shape) for f in feats] == [(2, 3, 64, 64), (2, 96, 16, 16), (2, 192, 8, 8), (2, 384, 4, 4), (2, 768, 2, 2)]


@pytest.mark.parametrize('extra', ['maxpool', 'p6p7'])
def test_oracle_fpn_matches_torchvision_fpn(extra):
    """mmdet 3.0.0's FPN (what torchok/models/necks/detection/fpn.py:61-117 wraps) is not in this image; torchvision's
    FeaturePyramidNetwork is an independent implementation of the same pyramid: 1x1 laterals, nearest top-down add,
    3x3 output convs, and either a stride-2 max-pool level (mmdet: add_extra_convs=False) or P6/P7 convs on C5 with a
    ReLU in between (mmdet: add_extra_convs='on_input', relu_before_extra_convs=True).  Same weights => same outputs
    and input gradients."""
    from collections import OrderedDict

    from torchvision.ops import FeaturePyramidNetwork
    from torchvision.ops.feature_pyramid_network import LastLevelMaxPool, LastLevelP6P7
    torch.manual_seed(11)
    chans, out_c = [24, 40, 64], 16
    blocks = LastLevelMaxPool() if extra == 'maxpool' else LastLevelP6P7(chans[-1], out_c)
    tv = FeaturePyramidNetwork(chans, out_c, extra_blocks=blocks)
    kw = dict(num_outs=4) if extra == 'maxpool' else \
        dict(num_outs=5, add_extra_convs='on_input', relu_before_extra_convs=True)
    fpn = om.FPN(chans[::-1], out_c, **kw)               # torchok hands the channels over reversed (fpn.py:66)
    with torch.no_grad():
        for i in range(3):
            fpn.lateral_convs[i].conv.load_state_dict(tv.inner_blocks[i][0].state_dict())
            fpn.fpn_convs[i].conv.load_state_dict(tv.layer_blocks[i][0].state_dict())
        if extra == 'p6p7':
            fpn.fpn_convs[3].conv.load_state_dict(tv.extra_blocks.p6.state_dict())
            fpn.fpn_convs[4].conv.load_state_dict(tv.extra_blocks.p7.state_dict())
    feats = [torch.randn(2, c, s, s) for c, s in zip(chans, (20, 10, 5))]
    a = [f.clone().requires_grad_(True) for f in feats]
    b = [f.clone().requires_grad_(True) for f in feats]
    want = list(tv(OrderedDict((str(i), f) for i, f in enumerate(a))).values())
    got = list(fpn(b))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and torch.allclose(g, w, rtol=1e-5, atol=1e-6), float((g - w).abs().max())
    sum((w * w).sum() for w in want).backward()
    sum((g * g).sum() for g in got).backward()
    for x, y in zip(a, b):
        assert torch.allclose(x.grad, y.grad, rtol=1e-4, atol=1e-5)


def _copy_swin_block(o, tv, dim):
    sd = o.state_dict()
    with torch.no_grad():
        tv.norm1.load_state_dict({'weight': sd['norm1.weight'], 'bias': sd['norm1.bias']})
        tv.norm2.load_state_dict({'weight': sd['norm2.weight'], 'bias': sd['norm2.bias']})
        tv.attn.qkv.weight.copy_(sd['attn.qkv.weight'])
        tv.attn.qkv.bias.copy_(torch.cat([sd['attn.q_bias'], torch.zeros(dim), sd['attn.v_bias']]))
        tv.attn.proj.load_state_dict({'weight': sd['attn.proj.weight'], 'bias': sd['attn.proj.bias']})
        tv.attn.logit_scale.copy_(sd['attn.logit_scale'])
        tv.attn.cpb_mlp[0].load_state_dict({'weight': sd['attn.cpb_mlp.0.weight'], 'bias': sd['attn.cpb_mlp.0.bias']})
        tv.attn.cpb_mlp[2].weight.copy_(sd['attn.cpb_mlp.2.weight'])
        tv.mlp[0].load_state_dict({'weight': sd['mlp.fc1.weight'], 'bias': sd['mlp.fc1.bias']})
        tv.mlp[3].load_state_dict({'weight': sd['mlp.fc2.weight'], 'bias': sd['mlp.fc2.bias']})


def test_oracle_swinv2_network_matches_torchvision():
    """The WIRING of the Swin-V2 restatement (patch embedding -> stages of alternately shifted blocks -> PatchMerging
    between stages -> final LayerNorm -> B x C x H x W, torchok/models/backbones/swin.py:204-256 over timm 0.6.13) against
    torchvision's independent SwinTransformer(block=SwinTransformerBlockV2, downsample_layer=PatchMergingV2): same
    weights, 128x128 input, window 4 (every stage grid >= the window, where timm shrinks the window and torchvision
    pads instead), outputs and the input gradient."""
    from functools import partial

    from torchvision.models.swin_transformer import PatchMergingV2, SwinTransformer, SwinTransformerBlockV2

    from oracle import swin as osw
    torch.manual_seed(5)
    depths, heads, dim = (2, 2, 2, 2), (1, 2, 4, 8), 32
    o = osw.SwinTransformerV2(img_size=128, embed_dim=dim, depths=depths, num_heads=heads, window_size=4)
    osw.dedegenerate_ln_(o, 3)
    tv = SwinTransformer(patch_size=[4, 4], embed_dim=dim, depths=list(depths), num_heads=list(heads), window_size=[4, 4],
                         stochastic_depth_prob=0.0, num_classes=3, block=SwinTransformerBlockV2,
                         downsample_layer=PatchMergingV2, norm_layer=partial(torch.nn.LayerNorm, eps=1e-5))
    with torch.no_grad():
        tv.features[0][0].load_state_dict(o.patch_embed.proj.state_dict())
        tv.features[0][2].load_state_dict(o.patch_embed.norm.state_dict())
        for i, layer in enumerate(o.layers):
            for j, blk in enumerate(layer.blocks):
                _copy_swin_block(blk, tv.features[1 + 2 * i][j], dim * 2 ** i)
            if layer.downsample is not None:
                tv.features[2 + 2 * i].reduction.weight.copy_(layer.downsample.reduction.weight)
                tv.features[2 + 2 * i].norm.load_state_dict(layer.downsample.norm.state_dict())
        tv.norm.load_state_dict(o.feature_norms[-1].state_dict())
    o.eval(), tv.eval()
    x = torch.randn(2, 3, 128, 128)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    a = o(xa)
    b = tv.permute(tv.norm(tv.features(xb)))
    assert a.shape == b.shape == (2, dim * 8, 4, 4)
    assert torch.allclose(a, b, atol=5e-5, rtol=1e-4), float((a - b).abs().max())
    r = torch.randn_like(a)
    (a * r).sum().backward()
    (b * r).sum().backward()
    assert torch.allclose(xa.grad, xb.grad, atol=5e-5, rtol=1e-3), float((xa.grad - xb.grad).abs().max())
