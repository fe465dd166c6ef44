"""GPU parity of the Swin-V2 path against oracle/swin.py (itself cross-checked against
torchvision's SwinTransformerBlockV2).  Tolerance 1e-2 of the tensor maximum for bf16 tensors (north_star), with the
whole-network bar "GPU error vs fp32 oracle <= 1.5 x oracle-bf16-AMP error + 5e-3" used for the other backbones."""
import copy

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('rows,c,res', [(64, 96, False), (50, 768, True), (33, 200, True)])
def test_layernorm_residual(rows, c, res):
    from torchok_b200 import kernels as K
    torch.manual_seed(c)
    x, r = _bf(torch.randn(rows, c) * 2 + 0.5), _bf(torch.randn(rows, c))
    w, b = torch.rand(c) + 0.5, torch.randn(c) * 0.1
    xo, ro, wo, bo = (t.clone().requires_grad_(True) for t in (x, r, w, b))
    y = F.layer_norm(xo, (c,), wo, bo)
    if res:
        y = ro + y
    g = _bf(torch.randn_like(y))
    (y * g).sum().backward()
    xm, rm = x.cuda().requires_grad_(True), r.cuda().requires_grad_(True)
    wm, bm = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    ym = K.layernorm(xm.to(torch.bfloat16), wm, bm, 1e-5, rm.to(torch.bfloat16) if res else None)
    (ym.float() * g.cuda()).sum().backward()
    assert rel_err(ym, y) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 1e-2
    assert rel_err(wm.grad, wo.grad) < 1e-2 and rel_err(bm.grad, bo.grad) < 1e-2
    if res:
        assert rel_err(rm.grad, ro.grad) < 1e-2


@pytest.mark.parametrize('rows,c', [(128, 96), (100, 256), (37, 384)])
def test_gelu(rows, c):
    """exact-erf GELU; for C % 128 == 0 the backward also returns the column sums of dx (Mlp.fc1's bias gradient)."""
    from torchok_b200 import kernels as K
    torch.manual_seed(c)
    x = _bf(torch.randn(rows, c) * 2)
    xo, xm = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    g = _bf(torch.randn_like(x))
    (F.gelu(xo) * g).sum().backward()
    fused = K.gelu_fuses_colsum(c)
    bias = torch.zeros(c, device='cuda', requires_grad=True)
    ym = K.gelu(xm.to(torch.bfloat16), bias if fused else None)
    (ym.float() * g.cuda()).sum().backward()
    assert rel_err(ym, F.gelu(x)) < 1e-2 and rel_err(xm.grad, xo.grad) < 1e-2
    if fused:
        assert rel_err(bias.grad, xo.grad.sum(0)) < 1e-2


@pytest.mark.parametrize('m,hidden,out', [(300, 384, 96), (1000, 768, 192), (77, 40, 24), (4096, 3072, 768)])
def test_mlp_tail_fused_gelu_backward(m, hidden, out):
    """fc2(GELU(h)) (timm Mlp, built by torchok/models/backbones/swin.py:71-81): the opt-in fused node — gelu' in the
    epilogue of fc2's data-gradient GEMM, tok_linear_dgrad_gelu (measured slower than the two passes, so off by default:
    kernels.py) — against fp32 autograd (1e-2 of the maximum) and against the two-launch sequence it replaces,
    tok_linear_dgrad + tok_gelu_bwd (same rounding points: bit-identical dh)."""
    from torchok_b200 import kernels as K
    from torchok_b200._lib import lib
    from torchok_b200.kernels import _p, _st
    torch.manual_seed(hidden + out)
    h = _bf(torch.randn(m, hidden) * 1.5)
    w = _bf(torch.randn(out, hidden) / hidden ** 0.5)
    b = torch.randn(out) * 0.1
    g = _bf(torch.randn(m, out))
    ho, wo, bo = (t.clone().requires_grad_(True) for t in (h, w, b))
    (F.linear(F.gelu(ho), wo, bo) * g).sum().backward()
    hm = h.cuda().to(torch.bfloat16).requires_grad_(True)
    wm, bm = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    fc1_bias = torch.zeros(hidden, device='cuda', requires_grad=True)
    y = K.GeluLinearFn.apply(hm, wm, bm, fc1_bias, False)   # the opt-in node (TOK_GELU_DGRAD=1), called directly
    (y.float() * g.cuda()).sum().backward()
    assert rel_err(y, F.linear(F.gelu(h), w, b)) < 1e-2
    assert rel_err(hm.grad, ho.grad) < 1e-2
    assert rel_err(wm.grad, wo.grad) < 1e-2 and rel_err(bm.grad, bo.grad) < 1e-2
    assert rel_err(fc1_bias.grad, ho.grad.sum(0)) < 1e-2
    # the two launches it replaces
    L = lib()
    gb, wb, hb = g.cuda().to(torch.bfloat16), w.cuda().to(torch.bfloat16), h.cuda().to(torch.bfloat16)
    da = torch.empty(m, hidden, device='cuda', dtype=torch.bfloat16)
    L.tok_linear_dgrad(m, out, hidden, _p(gb), _p(wb), _p(da), _st())
    dh2 = torch.empty_like(da)
    L.tok_gelu_bwd(da.numel(), hidden, _p(hb), _p(da), _p(dh2), None, _st())
    dh1 = torch.full_like(da, float('nan'))
    col = torch.zeros(hidden, device='cuda')
    L.tok_linear_dgrad_gelu(m, out, hidden, _p(gb), _p(wb), _p(hb), _p(dh1), _p(col), _st())
    torch.cuda.synchronize()
    assert torch.equal(dh1, dh2)
    ref = dh2.float().sum(0)
    assert float((col - ref).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-4


@pytest.mark.parametrize('rows,c', [(70, 96), (33, 384), (9, 1024)])
def test_layernorm_colsum(rows, c):
    """LayerNorm backward also accumulates the column sums of dx into the bias of the linear layer feeding it."""
    from torchok_b200 import kernels as K
    assert K.layernorm_fuses_colsum(c)
    torch.manual_seed(rows)
    x = _bf(torch.randn(rows, c) * 3 - 1)
    w, b = torch.rand(c) + 0.5, torch.randn(c) * 0.1
    xo = x.clone().requires_grad_(True)
    g = _bf(torch.randn(rows, c))
    (F.layer_norm(xo, (c,), w, b) * g).sum().backward()
    xm = x.cuda().requires_grad_(True)
    wm, bm = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    lin_bias = torch.zeros(c, device='cuda', requires_grad=True)
    ym = K.layernorm(xm.to(torch.bfloat16), wm, bm, 1e-5, colsum_param=lin_bias)
    (ym.float() * g.cuda()).sum().backward()
    assert rel_err(xm.grad, xo.grad) < 1e-2
    assert rel_err(lin_bias.grad, xo.grad.sum(0)) < 1e-2


@pytest.mark.parametrize('b,hw,e', [(3, 32, 96), (2, 224, 96), (2, 64, 128)])
def test_patch_embed(b, hw, e):
    """timm PatchEmbed.proj (Conv2d(3, E, 4, 4) on the NCHW image) + flatten/transpose; weights in channels_last memory
    as the parameter arena keeps them."""
    from torchok_b200 import kernels as K
    assert K.patch_embed_supported(3, 4, hw, hw, e)
    torch.manual_seed(e + hw)
    # bf16-representable image and weights: the tensor-core path (r3: tok_patchify + tok_linear_*) multiplies bf16 operands
    # like the reference's precision-16 mode does, so on these inputs it is as exact as the fp32 CUDA-core kernels
    x = _bf(torch.randn(b, 3, hw, hw))
    w = _bf(torch.randn(e, 3, 4, 4) * 0.2)
    bias = torch.randn(e) * 0.1
    wo, bo = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = F.conv2d(x, wo, bo, stride=4).flatten(2).transpose(1, 2).reshape(-1, e)
    g = _bf(torch.randn_like(ref))
    (ref * g).sum().backward()
    for fmt in (torch.contiguous_format, torch.channels_last):
        wm = w.cuda().contiguous(memory_format=fmt).requires_grad_(True)
        bm = bias.cuda().requires_grad_(True)
        out = K.patch_embed(x.cuda(), wm, bm)
        (out.float() * g.cuda()).sum().backward()
        assert rel_err(out, ref) < 1e-2
        assert rel_err(wm.grad, wo.grad) < 1e-3 and rel_err(bm.grad, bo.grad) < 1e-3


def test_patch_merge_is_the_reference_cat():
    """timm PatchMerging gather: bit-exact against torch.cat of the four strided slices, forward and backward."""
    from torchok_b200 import kernels as K
    b, h, w, c = 3, 8, 6, 24
    x = _bf(torch.randn(b, h, w, c))
    xo = x.clone().requires_grad_(True)
    ref = torch.cat([xo[:, 0::2, 0::2, :], xo[:, 1::2, 0::2, :], xo[:, 0::2, 1::2, :], xo[:, 1::2, 1::2, :]], -1)
    ref = ref.reshape(-1, 4 * c)
    g = _bf(torch.randn_like(ref))
    (ref * g).sum().backward()
    xm = x.cuda().to(torch.bfloat16).view(-1, c).requires_grad_(True)
    out = K.patch_merge(xm, b, h, w)
    (out.float() * g.cuda()).sum().backward()
    assert torch.equal(out.float().cpu(), ref.detach())
    assert torch.equal(xm.grad.float().cpu().view(b, h, w, c), xo.grad)


@pytest.mark.parametrize('dim,heads,res,ws,shift', [(96, 3, 8, 4, 0), (96, 3, 8, 4, 2), (64, 2, 14, 7, 3), (192, 6, 8, 8, 0),
                                                    (128, 4, 16, 8, 4),
                                                    # windows above 8x8 (large-window kernels): 12 / 16 / 24, with and
                                                    # without the cyclic shift
                                                    (96, 3, 24, 12, 6), (96, 3, 16, 16, 0), (64, 2, 32, 16, 8),
                                                    (64, 2, 24, 24, 0), (32, 1, 48, 24, 12)])
def test_swin_block_forward_backward(dim, heads, res, ws, shift):
    """timm SwinTransformerBlock: attention (bias, logit scale, shift mask), res-post-norm, MLP — every gradient."""
    from oracle import swin as osw
    from torchok_b200.models.backbones import swin as psw
    torch.manual_seed(dim + res + shift)
    o = osw.SwinTransformerBlock(dim, (res, res), heads, window_size=ws, shift_size=shift)
    osw.dedegenerate_ln_(o, 2)
    with torch.no_grad():
        for n_, p in o.named_parameters():
            if p.dim() == 2 and 'cpb' not in n_:
                p.copy_(_bf(p * 5))   # weights with visible magnitude, bf16-representable
    m = psw.SwinTransformerBlock(dim, (res, res), heads, window_size=ws, shift_size=shift)
    m.load_state_dict(o.state_dict())
    m.cuda().train()
    o.train()
    b = 3
    x = _bf(torch.randn(b, res * res, dim))
    xo = x.clone().requires_grad_(True)
    xm = x.cuda().requires_grad_(True)
    yo = o(xo)
    g = _bf(torch.randn_like(yo))
    (yo * g).sum().backward()
    ym = m(xm.view(-1, dim).to(torch.bfloat16), b)
    (ym.float().view(b, -1, dim) * g.cuda()).sum().backward()
    assert rel_err(ym.view(b, -1, dim), yo) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 2e-2
    po = dict(o.named_parameters())
    # calibration for the cancellation-prone gradients: the oracle's own bf16-AMP evaluation of the same block
    from oracle import models as om
    o16 = copy.deepcopy(o)
    o16.zero_grad()
    with om.amp_bf16():
        (o16(x.clone()) * g).sum().backward()
    pa = dict(o16.named_parameters())
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        e, e_amp = rel_l2(p.grad, po[k].grad), rel_l2(pa[k].grad, po[k].grad)
        # (logit_scale is a sum over every (query, key) pair of dS * cos with sum_j dS_ij = 0: a heavily cancelling sum —
        # one scalar per head — that only meets 3e-2 at windows <= 8x8 because the kernels form the cosine from split-bf16
        # operands; at 576-token windows even the oracle's AMP evaluation moves it by tens of percent, hence the
        # calibrated alternative)
        assert e < max(3e-2, 1.5 * e_amp + 1e-2), (k, e, e_amp)


def test_swinv2_network_forward_features_and_backward():
    import torchok_b200 as tb
    from oracle import models as om
    from oracle import swin as osw
    torch.manual_seed(0)
    kw = dict(img_size=64, window_size=8, depths=(2, 2, 2, 2))
    o = osw.SwinTransformerV2(**kw)
    osw.dedegenerate_ln_(o, 0)
    m = tb.BACKBONES.get('swinv2_custom')(pretrained=False, drop_path_rate=0.0, **kw)
    m.load_state_dict(o.state_dict())
    m.cuda().train()
    o16 = copy.deepcopy(o)
    x = torch.randn(4, 3, 64, 64)
    with torch.no_grad():
        fo = o.forward_features(x)
        with om.amp_bf16():
            fa = o16.forward_features(x)
    fm = m.forward_features(x.cuda())
    assert [tuple(f.shape) for f in fm] == [(4, 3, 64, 64), (4, 96, 16, 16), (4, 192, 8, 8), (4, 384, 4, 4), (4, 768, 2, 2)]
    for i, (a, b_, c) in enumerate(zip(fm[1:], fo[1:], fa[1:])):
        e, e_amp = rel_err(a, b_), rel_err(c, b_)
        print(f'swinv2 stage {i}: gpu-vs-fp32 {e:.4f} | oracle-amp-vs-fp32 {e_amp:.4f}')
        assert e < 1.5 * e_amp + 5e-3, (i, e, e_amp)
    grads = {}
    r = None
    for mode in ('amp', 'fp32'):
        o.zero_grad()
        with om.amp_bf16(mode == 'amp'):
            y = o(x)
            r = torch.randn_like(y) if r is None else r
            (y * r).sum().backward()
        grads[mode] = {k: p.grad.clone() for k, p in o.named_parameters() if p.grad is not None}
    ym = m(x.cuda())
    assert tuple(ym.shape) == (4, 768, 2, 2)
    (ym.float() * r.cuda()).sum().backward()
    worst = worst_amp = 0.0
    for k, p in m.named_parameters():
        if k not in grads['fp32']:
            continue
        assert p.grad is not None, k
        e = rel_l2(p.grad, grads['fp32'][k])
        e_amp = rel_l2(grads['amp'][k], grads['fp32'][k])
        worst, worst_amp = max(worst, e), max(worst_amp, e_amp)
        assert e < 1.5 * e_amp + 2e-2, (k, e, e_amp)
    print(f'swinv2 backward: worst rel_l2 gpu-vs-fp32 {worst:.4f} | oracle-amp-vs-fp32 {worst_amp:.4f}')


def test_swin_classification_task_step():
    """BASELINE config 3 in miniature: ClassificationTask(swinv2_custom) + Pooling + ClassificationHead + CE."""
    import torchok_b200 as tb
    cfg = tb.load_config({
        'task': {'name': 'ClassificationTask', 'params': {
            'backbone_name': 'swinv2_custom',
            'backbone_params': {'pretrained': False, 'img_size': 64, 'window_size': 8, 'depths': (1, 1, 1, 1)},
            'pooling_name': 'Pooling', 'head_name': 'ClassificationHead', 'head_params': {'num_classes': 10}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]},
    })
    task = tb.TASKS.get('ClassificationTask')(cfg, **cfg.task.params).cuda().train()
    out = task.training_step({'image': torch.randn(8, 3, 64, 64).cuda(), 'target': torch.randint(0, 10, (8,)).cuda()})
    out['loss'].backward()
    assert torch.isfinite(out['loss'])
    # feature_norms 0-2 only serve forward_features (swin.py:240-249 of the reference); everything else must train
    missing = [n for n, p in task.named_parameters() if p.grad is None and 'feature_norms' not in n]
    assert not missing, missing
    assert task.backbone.feature_norms[3].weight.grad is not None


def test_swinv2_tiny_window16_256_is_the_references_own_test_model():
    """tests/additional_tests/models/backbones/test_backbone.py:161-182 of the reference builds
    `swinv2_tiny_window16_256` at 256 x 256 and checks the output / feature shapes; here additionally parity of the four
    feature maps with the oracle (windows of 16 x 16 = 256 tokens in stages 1-2, 16 and 8 in stages 3-4)."""
    import torchok_b200 as tb
    from oracle import models as om
    from oracle import swin as osw
    torch.manual_seed(0)
    o = osw.SwinTransformerV2(img_size=256, window_size=16)
    osw.dedegenerate_ln_(o, 0)
    m = tb.BACKBONES.get('swinv2_tiny_window16_256')(pretrained=False, drop_path_rate=0.0)
    m.load_state_dict(o.state_dict())
    m.cuda().eval()
    o.eval()
    o16 = copy.deepcopy(o)
    x = torch.randn(2, 3, 256, 256)
    with torch.no_grad():
        fo = o.forward_features(x)
        with om.amp_bf16():
            fa = o16.forward_features(x)
        fm = m.forward_features(x.cuda())
        last = m(x.cuda())
    assert tuple(last.shape) == (2, 768, 8, 8)
    assert [tuple(f.shape) for f in fm] == [(2, 3, 256, 256), (2, 96, 64, 64), (2, 192, 32, 32), (2, 384, 16, 16),
                                            (2, 768, 8, 8)]
    for i, (a, b_, c) in enumerate(zip(fm[1:], fo[1:], fa[1:])):
        e, e_amp = rel_err(a, b_), rel_err(c, b_)
        print(f'swinv2_tiny_window16_256 stage {i}: gpu-vs-fp32 {e:.4f} | oracle-amp-vs-fp32 {e_amp:.4f}')
        assert e < 1.5 * e_amp + 5e-3, (i, e, e_amp)
