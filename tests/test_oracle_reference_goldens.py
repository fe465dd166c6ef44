"""The CPU oracle (oracle/models.py) replayed against vectors produced by the REFERENCE's own code
(tests/golden/make_reference_goldens.py loads the reference files by path under stub registries; the fixture is
tests/golden/reference_goldens.pt).  This is what pins the oracle for SURVEY §8 rows a3, a8, a12-a15, a17 and the N4
DiceLoss: same state_dict, same seeded inputs, outputs and every gradient equal to fp32 round-off."""
import os

import pytest
import torch

from oracle import models as om

G = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens.pt'), weights_only=False)
TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    tol = dict(TOL, **kw)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, **tol), float((a - b).abs().max())


@pytest.mark.parametrize('case', G['ConvBnAct'], ids=lambda c: str(c['args']))
def test_conv_bn_act(case):
    cin, cout, k, stride, pad, act, train = case['args']
    m = om.ConvBnAct(cin, cout, k, padding=pad, stride=stride, act=act)
    m.load_state_dict(case['state'])
    m.train(train)
    x = case['x'].clone().requires_grad_(True)
    y = m(x)
    (y * case['r']).sum().backward()
    close(y, case['y'])
    close(x.grad, case['dx'], atol=2e-5)
    close(m.conv.weight.grad, case['dw'], atol=2e-5)
    close(m.bn.weight.grad, case['dgamma'], atol=2e-5)
    close(m.bn.bias.grad, case['dbeta'], atol=2e-5)
    for key in ('bn.running_mean', 'bn.running_var'):
        close(m.state_dict()[key], case['state_after'][key])


@pytest.mark.parametrize('case', G['LinearHead'], ids=lambda c: f"{c['kind']}{c['args']}")
def test_linear_and_classification_head(case):
    cin, cout, norm = case['args']
    m = om.LinearHead(cin, cout, normalize=norm) if case['kind'] == 'LinearHead' else om.ClassificationHead(cin, cout)
    m.load_state_dict(case['state'])
    x = case['x'].clone().requires_grad_(True)
    y = m(x)
    (y * case['r']).sum().backward()
    close(y, case['y'])
    close(x.grad, case['dx'])
    close(m.fc.weight.grad, case['dw'])
    close(m.fc.bias.grad, case['db'])


@pytest.mark.parametrize('case', G['ArcFaceHead'], ids=lambda c: str(c['args']))
def test_arcface_head(case):
    cin, ncls, kw, train = case['args']
    m = om.ArcFaceHead(cin, ncls, **kw)
    assert m.scale == pytest.approx(case['scale'], rel=1e-12) and m.margin == pytest.approx(case['margin'], rel=1e-12)
    m.load_state_dict(case['state'])
    m.train(train)
    x = case['x'].clone().requires_grad_(True)
    y = m(x, case['target'])
    (y * case['r']).sum().backward()
    close(y, case['y'], atol=2e-5)
    close(x.grad, case['dx'], atol=2e-5)
    close(m.weight.grad, case['dw'], atol=2e-5)


@pytest.mark.parametrize('case', G['SegmentationHead'], ids=lambda c: str(c['args']))
def test_segmentation_head(case):
    cin, ncls, size = case['args']
    m = om.SegmentationHead(cin, ncls)
    m.load_state_dict(case['state'])
    f = case['f'].clone().requires_grad_(True)
    y = m([torch.zeros(f.shape[0], 3, size, size), f])
    (y * case['r']).sum().backward()
    close(y, case['y'])
    close(f.grad, case['df'])


def test_hrnet_segmentation_neck():
    case = G['HRNetSegmentationNeck'][0]
    m = om.HRNetSegmentationNeck(case['chans'])
    m.load_state_dict(case['state'])
    m.eval()
    feats = [case['feats'][0]] + [f.clone().requires_grad_(True) for f in case['feats'][1:]]
    y = m(feats)[-1]
    (y * case['r']).sum().backward()
    close(y, case['y'], atol=2e-5)
    for f, d in zip(feats[1:], case['dfeats']):
        close(f.grad, d, atol=2e-5)


@pytest.mark.parametrize('case', G['ContrastiveLoss'], ids=lambda c: str(c['args']))
def test_contrastive_loss(case):
    margin, reg, red = case['args']
    e1, e2 = case['e1'].clone().requires_grad_(True), case['e2'].clone().requires_grad_(True)
    val = om.ContrastiveLoss(margin, reg, red)(e1, e2, case['R'])
    val.backward()
    close(val, case['loss'], atol=2e-5)
    close(e1.grad, case['d1'], atol=2e-5)
    close(e2.grad, case['d2'], atol=2e-5)


@pytest.mark.parametrize('case', G['calc_relevance_matrix'], ids=lambda c: str(c['num_classes']))
def test_calc_relevance_matrix(case):
    assert torch.equal(om.calc_relevance_matrix(case['y'], case['num_classes']), case['R'])


@pytest.mark.parametrize('case', G['DiceLoss'], ids=lambda c: str(c['kwargs']))
def test_dice_loss(case):
    x = case['logits'].clone().requires_grad_(True)
    val = om.dice_loss_multiclass(x, case['target'], **case['kwargs'])
    val.backward()
    close(val, case['loss'])
    close(x.grad, case['grad'])


# ------------------------------------------------------------------ host logic pinned by the reference's own code
@pytest.mark.parametrize('case', G['JointLoss'], ids=lambda c: f"{c['weights']}-{c['normalize']}")
def test_joint_loss_matches_reference(case):
    """torchok_b200.losses.JointLoss (host logic, shipped) vs the reference's losses/base.py executed by path."""
    from torchok_b200.losses.base import JointLoss
    j = JointLoss([torch.nn.MSELoss(), torch.nn.L1Loss()],
                  [{'input': 'pred_a', 'target': 'gt'}, {'input': 'pred_b', 'target': 'gt'}],
                  tags=['mse', 'l1'], weights=list(case['weights']), normalize_weights=case['normalize'])
    total, tagged = j(**case['inputs'])
    close(total, case['total'])
    assert set(tagged) == set(case['tagged'])
    for k, v in case['tagged'].items():
        close(tagged[k], v)


@pytest.mark.parametrize('case', G['paramwise_cfg'], ids=lambda c: ','.join(sorted(c['cfg'])))
def test_paramwise_cfg_matches_reference(case):
    """constructor/paramwise.py vs the param groups built by the reference's Constructor.add_params
    (constructor/constructor.py:162-251) on the same module tree."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        '_mk_goldens', os.path.join(os.path.dirname(__file__), 'golden', 'make_reference_goldens.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    from torchok_b200.constructor.paramwise import paramwise_multipliers
    module = mk.build_from_spec(case['spec'])
    for name in case['frozen']:
        dict(module.named_parameters())[name].requires_grad_(False)
    mult = paramwise_multipliers(module, case['cfg'])
    got = {n: mult[p] for n, p in module.named_parameters()}
    assert set(got) == set(case['multipliers'])
    for n, (lr, wd) in case['multipliers'].items():
        assert got[n] == pytest.approx((lr, wd), rel=1e-12, abs=1e-12), (n, got[n], (lr, wd))
