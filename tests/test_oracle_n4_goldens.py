"""The CPU oracle of the N4 parts (OCRSegmentationHead, UnetNeck — SURVEY §8f; their kernels are not written yet)
replayed against vectors produced by the REFERENCE's own files (tests/golden/make_n4_goldens.py executes
torchok/models/heads/segmentation/ocr.py and torchok/models/necks/segmentation/unet.py by path; fixture
tests/golden/n4_goldens.pt): same state_dict, same seeded inputs, outputs / every gradient / BatchNorm running
statistics equal to fp32 round-off.  This pins the target the CUDA path will be held to."""
import os

import pytest
import torch

from oracle import models as om

G = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'n4_goldens.pt'), weights_only=False)


def close(a, b, rtol=2e-5, atol=2e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


def check_params(m, case):
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(grads) == set(case['grads'])
    for n, gref in case['grads'].items():
        close(grads[n], gref, atol=5e-5)
    now = m.state_dict()
    for k, v in case['state_after'].items():
        close(now[k], v)


@pytest.mark.parametrize('case', G['OCRSegmentationHead'], ids=lambda c: f"{c['args']}-train{int(c['train'])}")
def test_ocr_segmentation_head(case):
    cin, ncls, mid, key = case['args']
    m = om.OCRSegmentationHead(cin, ncls, ocr_mid_channels=mid, ocr_key_channels=key)
    m.load_state_dict(case['state'])
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m.train(case['train'])
    f = case['f'].clone().requires_grad_(True)
    y = m([case['image'], f])
    ys = y if isinstance(y, tuple) else (y,)
    assert len(ys) == len(case['ys']) == (2 if case['train'] else 1)
    for t, ref in zip(ys, case['ys']):
        close(t, ref)
    sum((t * r).sum() for t, r in zip(ys, case['rs'])).backward()
    close(f.grad, case['df'], atol=5e-5)
    check_params(m, case)


@pytest.mark.parametrize('case', G['UnetNeck'], ids=lambda c: f"{c['args']}-train{int(c['train'])}")
def test_unet_neck(case):
    chans, dec, center, use_bn = case['args']
    m = om.UnetNeck(list(chans), decoder_channels=dec, use_batchnorm=use_bn, center=center)
    m.load_state_dict(case['state'])
    m.train(case['train'])
    feats = [case['feats'][0]] + [f.clone().requires_grad_(True) for f in case['feats'][1:]]
    image, y = m(feats)
    assert image is feats[0] and m.out_channels == dec[-1]
    close(y, case['y'])
    (y * case['r']).sum().backward()
    for f, d in zip(feats[1:], case['dfeats']):
        close(f.grad, d, atol=5e-5)
    check_params(m, case)


@pytest.mark.parametrize('case', G['HRNetClassificationNeck'], ids=lambda c: f"train{int(c['train'])}")
def test_hrnet_classification_neck(case):
    """necks/classification/hrnet.py:12-85 executed by path with torchvision's independent Bottleneck standing in for
    timm's (same constructor order, arithmetic and state-dict keys): layer construction, key names, the S7 overwrite
    quirk (only the last branch reaches the output, so only its input receives a gradient), running statistics."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('_mk_n4', os.path.join(os.path.dirname(__file__), 'golden', 'make_n4_goldens.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    m = om.HRNetClassificationNeck(list(case['chans']))
    m.load_state_dict(mk.seeded_state(m, case['seed']))          # same keys and shapes as the reference's module
    m.train(case['train'])
    feats = [f.clone().requires_grad_(True) for f in case['feats']]
    y = m(feats)
    close(y, case['y'], rtol=1e-4, atol=1e-4)
    (y * case['r']).sum().backward()
    for f, d in zip(feats, case['dfeats']):
        if d is None:
            assert f.grad is None or float(f.grad.abs().max()) == 0.0
        else:
            close(f.grad, d, rtol=1e-4, atol=1e-4)
    norms = {n: float(p.grad.norm()) for n, p in m.named_parameters() if p.grad is not None}
    assert set(norms) == set(case['grad_norms'])
    for n, v in case['grad_norms'].items():
        assert norms[n] == pytest.approx(v, rel=2e-3, abs=1e-5), n
    now = m.state_dict()
    for k, v in case['state_after'].items():
        close(now[k], v, rtol=1e-4, atol=1e-5)
