"""GPU parity of the SURVEY 8f N4 parts — OCRSegmentationHead and UnetNeck — against vectors produced by the REFERENCE's
own files (tests/golden/n4_goldens.pt: torchok/models/heads/segmentation/ocr.py and necks/segmentation/unet.py executed
by path): strict state-dict load, outputs, input gradients, parameter gradients, BatchNorm running statistics.

Bars: the goldens are fp32, the kernels store bf16 activations, and these are stacks of 6-10 fused units evaluated on
2-sample batches, so every quantity is compared with what the ORACLE's bf16-AMP evaluation of the same case loses against
the fp32 golden (printed beside it): gpu error <= 1.5 x oracle-AMP error + 1e-2 (max norm, outputs and running
statistics) / + 2e-2 (relative L2, gradients: ReLU-mask flips of borderline bf16 pre-activations move single elements by
their full magnitude, see tests/test_reference_goldens_gpu.py).  Plus kernel-level tests of the new products against
torch on bf16-representable inputs at north_star's 1e-2."""
import os

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

G = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'n4_goldens.pt'), weights_only=False)


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('b,c,k,hw', [(2, 32, 5, 8), (3, 128, 19, 16), (1, 512, 3, 12), (2, 64, 1, 6)])
def test_spatial_gather_kernel(b, c, k, hw):
    from torchok_b200 import kernels as K
    torch.manual_seed(c + k)
    feats, logits = _bf(torch.randn(b, c, hw, hw)), _bf(torch.randn(b, k, hw, hw) * 2)
    fo, lo = feats.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    p = F.softmax(lo.view(b, k, -1), dim=2)
    ref = torch.matmul(p, fo.view(b, c, -1).permute(0, 2, 1)).permute(0, 2, 1).unsqueeze(3)
    r = _bf(torch.randn_like(ref))
    (ref * r).sum().backward()
    fm, lm = feats.cuda().requires_grad_(True), logits.cuda().requires_grad_(True)
    out = K.spatial_gather(fm, lm)
    assert tuple(out.shape) == (b, c, k, 1)
    (out.float() * r.cuda()).sum().backward()
    assert rel_err(out, ref) < 1e-2
    assert rel_err(fm.grad, fo.grad) < 1e-2
    assert rel_err(lm.grad, lo.grad) < 1.5e-2


@pytest.mark.parametrize('b,kc,k,hw', [(2, 16, 5, 8), (2, 64, 19, 16), (1, 8, 1, 6), (2, 256, 7, 5)])
def test_object_attention_kernel(b, kc, k, hw):
    from torchok_b200 import kernels as K
    torch.manual_seed(kc + k)
    q, key, val = _bf(torch.randn(b, kc, hw, hw)), _bf(torch.randn(b, kc, k, 1)), _bf(torch.randn(b, kc, k, 1))
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q, key, val))
    sim = F.softmax(kc ** -.5 * torch.matmul(qo.view(b, kc, -1).permute(0, 2, 1), ko.view(b, kc, -1)), dim=-1)
    ref = torch.matmul(sim, vo.view(b, kc, -1).permute(0, 2, 1)).permute(0, 2, 1).reshape(b, kc, hw, hw)
    r = _bf(torch.randn_like(ref))
    (ref * r).sum().backward()
    qm, km, vm = (t.cuda().requires_grad_(True) for t in (q, key, val))
    out = K.object_attention(qm, km, vm, kc ** -.5)
    (out.float() * r.cuda()).sum().backward()
    assert rel_err(out, ref) < 1e-2
    assert rel_err(qm.grad, qo.grad) < 1.5e-2
    assert rel_err(km.grad, ko.grad) < 1.5e-2
    assert rel_err(vm.grad, vo.grad) < 1.5e-2


@pytest.mark.parametrize('chans,sizes,out', [((8, 16), (4, 8), 8), ((12, 8), (5, 10), 10), ((24,), (3,), 6), ((8, 8), (4, 7), 8)])
def test_nearest_cat_kernel(chans, sizes, out):
    from torchok_b200 import kernels as K
    torch.manual_seed(sum(chans))
    xs = [_bf(torch.randn(2, c, s, s)) for c, s in zip(chans, sizes)]
    ref_in = [x.clone().requires_grad_(True) for x in xs]
    y = torch.cat([F.interpolate(x, size=(out, out), mode='nearest') for x in ref_in], 1)
    r = _bf(torch.randn_like(y))
    (y * r).sum().backward()
    gin = [x.cuda().requires_grad_(True) for x in xs]
    cat = K.nearest_cat(gin, (out, out))
    segs, off = [], 0
    for c in chans:
        segs.append(cat[:, off:off + c])
        off += K.ceil8(c)
    got = torch.cat(segs, 1)
    assert torch.equal(got.float().cpu(), y.detach())
    (got.float() * r.cuda()).sum().backward()
    for a, b_ in zip(gin, ref_in):
        assert rel_err(a.grad, b_.grad) < 1e-2


def _oracle_amp(build, case, run):
    """The oracle evaluated in its bf16-AMP mode on the golden's state and inputs: what the reference's own precision-16
    mode costs against the fp32 golden — the calibration of the bars below."""
    from oracle import models as om
    o = build()
    o.load_state_dict(case['state'])
    for mod in o.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    o.train(case['train'])
    with om.amp_bf16():
        outs, ins = run(o)
    return o, outs, ins


def _compare(tag, got, amp, ref, l2=False, slack=1e-2):
    f = rel_l2 if l2 else rel_err
    e, e_amp = f(got, ref), f(amp, ref)
    print(f'  {tag}: gpu {e:.4f} | oracle-amp {e_amp:.4f}')
    assert e < 1.5 * e_amp + slack, (tag, e, e_amp)


def _check_params(m, o, case):
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    ogr = {n: p.grad for n, p in o.named_parameters() if p.grad is not None}
    assert set(grads) == set(case['grads'])
    for n, g in case['grads'].items():
        if float(g.norm()) > 1e-6:
            # gradients with a handful of elements (the 2-channel BatchNorm of last_reduction at mid = 32) are sums over a
            # ReLU mask that a single flipped pixel moves by percents: wider absolute slack for them
            _compare(f'grad {n}', grads[n], ogr[n], g, l2=True, slack=2e-2 if g.numel() > 8 else 1e-1)
    now, onow = m.state_dict(), o.state_dict()
    for k, v in case['state_after'].items():
        _compare(k, now[k], onow[k], v)


@pytest.mark.parametrize('case', G['OCRSegmentationHead'], ids=lambda c: f"{c['args']}-train{int(c['train'])}")
def test_ocr_segmentation_head_replays_reference(case):
    import torchok_b200 as tb
    from oracle import models as om
    cin, ncls, mid, key = case['args']
    m = tb.HEADS.get('OCRSegmentationHead')(in_channels=cin, num_classes=ncls, ocr_mid_channels=mid, ocr_key_channels=key)
    m.load_state_dict(case['state'], strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m.cuda().train(case['train'])

    def run(o):
        fo = case['f'].clone().requires_grad_(True)
        yo = o([case['image'], fo])
        yo = yo if isinstance(yo, tuple) else (yo,)
        sum((t * r).sum() for t, r in zip(yo, case['rs'])).backward()
        return yo, fo
    o, yo, fo = _oracle_amp(lambda: om.OCRSegmentationHead(cin, ncls, ocr_mid_channels=mid, ocr_key_channels=key), case, run)
    f = case['f'].cuda().requires_grad_(True)
    y = m([case['image'].cuda(), f])
    ys = y if isinstance(y, tuple) else (y,)
    assert len(ys) == len(case['ys']) == (2 if case['train'] else 1)
    print(f'OCR head {case["args"]} train={case["train"]}')
    for i, (t, a, ref) in enumerate(zip(ys, yo, case['ys'])):
        assert tuple(t.shape) == tuple(ref.shape)
        _compare(f'output {i}', t, a, ref)
    sum((t.float() * r.cuda()).sum() for t, r in zip(ys, case['rs'])).backward()
    _compare('df', f.grad, fo.grad, case['df'], l2=True, slack=2e-2)
    _check_params(m, o, case)


@pytest.mark.parametrize('case', G['UnetNeck'], ids=lambda c: f"{c['args']}-train{int(c['train'])}")
def test_unet_neck_replays_reference(case):
    import torchok_b200 as tb
    from oracle import models as om
    chans, dec, center, use_bn = case['args']
    m = tb.NECKS.get('UnetNeck')(in_channels=list(chans), decoder_channels=dec, use_batchnorm=use_bn, center=center)
    m.load_state_dict(case['state'], strict=True)
    m.cuda().train(case['train'])

    def run(o):
        fs = [case['feats'][0]] + [t.clone().requires_grad_(True) for t in case['feats'][1:]]
        _, yo = o(fs)
        (yo * case['r']).sum().backward()
        return yo, fs
    o, yo, fso = _oracle_amp(lambda: om.UnetNeck(list(chans), decoder_channels=dec, use_batchnorm=use_bn, center=center),
                             case, run)
    feats = [case['feats'][0].cuda()] + [f.cuda().requires_grad_(True) for f in case['feats'][1:]]
    image, y = m(feats)
    assert image is feats[0] and m.out_channels == dec[-1]
    print(f'UnetNeck {case["args"]} train={case["train"]}')
    _compare('output', y, yo, case['y'])
    (y.float() * case['r'].cuda()).sum().backward()
    for i, (f, fo, d) in enumerate(zip(feats[1:], fso[1:], case['dfeats'])):
        _compare(f'dfeat {i}', f.grad, fo.grad, d, l2=True, slack=2e-2)
    _check_params(m, o, case)
