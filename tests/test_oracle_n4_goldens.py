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
