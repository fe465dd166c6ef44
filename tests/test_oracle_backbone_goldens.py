"""Backbone WIRING pinned against the reference's own resnet.py / hrnet.py / swin.py executed by path
(tests/golden/make_backbone_goldens.py; timm's blocks replaced by torchvision's / the oracle's, see that file's header).

The oracle networks (oracle/models.py::ResNet / HighResolutionNet, oracle/swin.py::SwinTransformerV2) must accept the
reference's state dict strictly (identical key names and shapes) and reproduce forward_features, forward, the input
gradient, every parameter-gradient norm and the updated BatchNorm running statistics to fp32 round-off.
"""
import os
import sys

import pytest
import torch

from oracle import models as om
from oracle import swin as osw

sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
from make_backbone_goldens import seeded_state  # noqa: E402

G = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'backbone_goldens.pt'))
CASES = [(name, i) for name, cs in G.items() for i in range(len(cs))]


def build_oracle(name, kwargs):
    if name.startswith('resnet'):
        return om.resnet(name)
    if name.startswith('hrnet'):
        return om.hrnet(name)
    kw = dict(img_size=kwargs.get('img_size', 256), window_size=kwargs.get('window_size', 7),
              depths=kwargs.get('depths', (2, 2, 6, 2)))
    if 'window8' in name:
        kw['window_size'] = 8
    if 'window16' in name:
        kw['window_size'] = 16
    return osw.SwinTransformerV2(**kw)


def close(a, b, tol=2e-5):
    scale = max(float(b.abs().max()), 1e-6)
    assert float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


@pytest.mark.parametrize('name,idx', CASES)
def test_oracle_network_matches_reference_wiring(name, idx):
    c = G[name][idx]
    m = build_oracle(name, c['kwargs'])
    state = seeded_state(m, c['seed'])
    # strict key equality, in order: the reference's parameter names are the state-dict contract
    assert [k for k in state.keys()] == c['keys']
    m.load_state_dict(state, strict=True)
    m.train(c['train'])
    x = c['x'].clone().requires_grad_(True)
    feats = m.forward_features(x)[1:]
    assert len(feats) == len(c['feats'])
    for f, g in zip(feats, c['feats']):
        assert f.shape == g.shape
        close(f, g)
    gen = torch.Generator().manual_seed(c['seed'] + 1)
    rs = [t.to(torch.bfloat16).float() for t in (torch.randn(f.shape, generator=gen) for f in feats)]
    sum((f * r).sum() for f, r in zip(feats, rs)).backward()
    close(x.grad, c['dx'], 1e-4)
    grads = dict(m.named_parameters())
    for n, (norm, total) in c['gnorm'].items():
        assert abs(float(grads[n].grad.norm()) - norm) <= 2e-4 * max(norm, 1e-3), n
    for n, g in c['grads'].items():
        close(grads[n].grad, g, 2e-4)
    last = m(c['x'])
    last = last if isinstance(last, (list, tuple)) else [last]
    for a, b in zip(last, c['forward']):
        close(a, b)
    assert tuple(m.out_encoder_channels) == tuple(c['out_encoder_channels'])
    if c['train']:
        sd = m.state_dict()
        # forward_features + forward = two training-mode passes, like the generator
        for k, v in c['running'].items():
            close(sd[k], v, 1e-5)
