"""GPU parity of the embedding heads, the pairwise loss and PairwiseLearnTask against the CPU oracle.

Tolerances: the kernels take bf16 GEMM operands (fp32 accumulate), so logits / gradients are held to 1e-2 of the
tensor's max magnitude (BASELINE.json north_star: 1e-2 bf16); the contrastive loss is fp32 end to end: 1e-4."""
import math

import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('b,d,c,easy', [(64, 512, 1000, False), (37, 128, 11318, False), (16, 64, 10, True)])
def test_arcface_head_train_and_eval(b, d, c, easy):
    """arcface_head.py:110-131: training branch with margin on the target column, eval = plain linear."""
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(b + c)
    o = om.ArcFaceHead(d, c, easy_margin=easy)
    m = tb.HEADS.get('ArcFaceHead')(in_channels=d, num_classes=c, easy_margin=easy)
    m.load_state_dict(o.state_dict())
    m.cuda()
    assert abs(m.scale - o.scale) < 1e-9 and abs(m.margin - o.margin) < 1e-9
    x = _bf(torch.randn(b, d))
    y = torch.randint(0, c, (b,))
    xo, xm = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    o.train(), m.train()
    lo = o(xo, y)
    r = torch.randn_like(lo)
    (lo * r).sum().backward()
    lm = m(xm, y.cuda())
    assert tuple(lm.shape) == (b, c)
    (lm.float() * r.cuda()).sum().backward()
    assert rel_err(lm, lo) < 1e-2
    # the margin column itself (where the two implementations could differ most)
    idx = torch.arange(b)
    assert rel_err(lm[idx.cuda(), y.cuda()], lo[idx, y]) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 1.5e-2
    assert rel_err(m.weight.grad, o.weight.grad) < 1.5e-2
    with pytest.raises(ValueError, match='Target is None'):
        m(xm)
    o.eval(), m.eval()
    with torch.no_grad():
        assert rel_err(m(x.cuda()), o(x)) < 1e-2


def test_arcface_defaults_match_reference_formulas():
    import torchok_b200 as tb
    h = tb.HEADS.get('ArcFaceHead')(in_channels=512, num_classes=11318)
    c = 11318
    assert h.scale == pytest.approx((c - 1) / c * math.log((c - 1) * .999 / .001) + 1)
    assert h.margin == pytest.approx(.5 * c / (c - 1))
    assert tuple(h.weight.shape) == (c, 512)
    h2 = tb.HEADS.get('ArcFaceHead')(in_channels=2, num_classes=10)
    assert h2.margin == pytest.approx(.9 - math.cos(2 * math.pi / 10))
    with pytest.raises(ValueError):
        tb.HEADS.get('ArcFaceHead')(in_channels=8, num_classes=10, dynamic_margin=True)


@pytest.mark.parametrize('b,d', [(32, 512), (7, 40)])
def test_linear_head_normalize(b, d):
    """LinearHead(normalize=True) (linear_head.py:27-36): FC then F.normalize, forward and backward."""
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(b)
    o = om.LinearHead(d, 64, normalize=True)
    m = tb.HEADS.get('LinearHead')(in_channels=d, out_channels=64, normalize=True)
    with torch.no_grad():
        o.fc.weight.copy_(_bf(o.fc.weight))
    m.load_state_dict(o.state_dict())
    m.cuda()
    x = _bf(torch.randn(b, d))
    xo, xm = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    yo = o(xo)
    r = torch.randn_like(yo)
    (yo * r).sum().backward()
    ym = m(xm)
    (ym.float() * r.cuda()).sum().backward()
    assert rel_err(ym, yo) < 1e-2
    assert rel_err(ym.float().norm(dim=1), torch.ones(b)) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 2e-2
    assert rel_err(m.fc.weight.grad, o.fc.weight.grad) < 2e-2


@pytest.mark.parametrize('b,m_,d,margin,reg,reduction', [(64, 64, 512, 0.5, None, 'mean'), (33, 50, 128, 1.0, 'L2', 'sum'),
                                                          (8, 8, 16, 0.2, 'L1', 'mean')])
def test_contrastive_loss(b, m_, d, margin, reg, reduction):
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(b + d)
    e1 = torch.nn.functional.normalize(torch.randn(b, d))
    e2 = torch.nn.functional.normalize(torch.randn(m_, d))
    R = (torch.rand(b, m_) < 0.2).float()
    lo_fn = om.ContrastiveLoss(margin, reg, reduction)
    lm_fn = tb.LOSSES.get('ContrastiveLoss')(margin=margin, reg=reg, reduction=reduction)
    a1, a2 = e1.clone().requires_grad_(True), e2.clone().requires_grad_(True)
    g1, g2 = e1.cuda().requires_grad_(True), e2.cuda().requires_grad_(True)
    lo = lo_fn(a1, a2, R)
    lo.backward()
    lm = lm_fn(g1, g2, R.cuda())
    lm.backward()
    assert abs(float(lm) - float(lo)) / abs(float(lo)) < 1e-4
    assert rel_err(g1.grad, a1.grad) < 1e-4 and rel_err(g2.grad, a2.grad) < 1e-4
    with pytest.raises(ValueError, match='Unknown reduction type'):
        tb.LOSSES.get('ContrastiveLoss')(margin=1.0, reduction='median')(g1, g2, R.cuda())


def test_contrastive_loss_shared_embedding_tensor():
    """PairwiseLearnTask feeds the SAME tensor as emb1 and emb2 (pairwise_task.py:79): both gradient paths add up,
    and the zero self-distances on the diagonal must not produce NaNs."""
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(3)
    e = torch.nn.functional.normalize(torch.randn(48, 64))
    y = torch.randint(0, 6, (48,))
    R = om.calc_relevance_matrix(y, 6)
    a = e.clone().requires_grad_(True)
    g = e.cuda().requires_grad_(True)
    om.ContrastiveLoss(0.7)(a, a, R).backward()
    loss = tb.LOSSES.get('ContrastiveLoss')(margin=0.7)(g, g, R.cuda())
    loss.backward()
    assert torch.isfinite(g.grad).all()
    assert rel_err(g.grad, a.grad) < 1e-4


def test_pairwise_task_forward_with_gt_and_relevance_matrix():
    """The reference's positional-argument bug is fixed: the pairwise_sop.yaml task block builds."""
    import torchok_b200 as tb
    from oracle import models as om
    cfg = tb.load_config({
        'task': {'name': 'PairwiseLearnTask', 'params': {
            'num_classes': 10, 'backbone_name': 'resnet18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
            'pooling_name': 'Pooling', 'head_name': 'LinearHead',
            'head_params': {'out_channels': 32, 'normalize': True}}},
        'joint_loss': {'losses': [{'name': 'ContrastiveLoss', 'params': {'margin': 0.5},
                                   'mapping': {'emb1': 'emb1', 'emb2': 'emb2', 'R': 'R'}}]},
    })
    task = tb.TASKS.get('PairwiseLearnTask')(cfg, **cfg.task.params).cuda().train()
    x = torch.randn(16, 3, 32, 32).cuda()
    y = torch.randint(0, 10, (16,)).cuda()
    out = task.forward_with_gt({'image': x, 'target': y})
    assert set(out) == {'emb1', 'emb2', 'R', 'target'} and out['emb1'] is out['emb2']
    assert tuple(out['emb1'].shape) == (16, 32)
    assert torch.equal(out['R'].cpu(), om.calc_relevance_matrix(y.cpu(), 10))
    step = task.training_step({'image': x, 'target': y})
    step['loss'].backward()
    assert torch.isfinite(step['loss']) and task.head.fc.weight.grad is not None
    yy = torch.zeros(4, 5).cuda()
    yy[0, 1] = yy[1, 1] = yy[2, 3] = 1
    assert torch.equal(task.calc_relevance_matrix(yy).cpu(), om.calc_relevance_matrix(yy.cpu(), 5))
