"""GPU parity of Pooling / PoolingLinear for every pooling_type the reference accepts
(torchok/models/poolings/classification/pooling.py:7-12 = timm SelectAdaptivePool2d(flatten=True); linear.py:8-25):
forward and backward against the oracle (F.adaptive_avg_pool2d / F.adaptive_max_pool2d) on bf16-representable inputs.
Bar: 1e-2 of the tensor maximum (north_star, bf16)."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('ptype', ['avg', 'max', 'avgmax', 'catavgmax'])
@pytest.mark.parametrize('n,c,hw', [(4, 64, 7), (3, 2048, 7), (2, 18, 5), (5, 512, 1)])
def test_pooling_types_fwd_bwd(ptype, n, c, hw):
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(n * c + hw)
    x = _bf(torch.randn(n, c, hw, hw))
    x[0, 0] = 0.0                          # an all-equal channel: the first position must take the max-path gradient
    m = tb.POOLINGS.get('Pooling')(in_channels=c, pooling_type=ptype)
    o = om.Pooling(c, ptype)
    assert m.out_channels == o.out_channels == (2 * c if ptype == 'catavgmax' else c)
    xo = x.clone().requires_grad_(True)
    yo = o(xo)
    r = _bf(torch.randn_like(yo))
    (yo * r).sum().backward()
    xm = x.cuda().requires_grad_(True)
    ym = m(xm)
    assert tuple(ym.shape) == tuple(yo.shape)
    (ym.float() * r.cuda()).sum().backward()
    assert rel_err(ym, yo) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 1e-2
    if ptype != 'avg':
        # exactly one position per (sample, channel) carries the max-path gradient: same sparsity pattern as the oracle
        nz_o = (xo.grad[0, 0] != xo.grad[0, 0].flatten()[-1]).sum() if hw > 1 else 0
        nz_m = (xm.grad[0, 0].cpu().float() != xm.grad[0, 0].cpu().float().flatten()[-1]).sum() if hw > 1 else 0
        assert int(nz_o) == int(nz_m)


@pytest.mark.parametrize('ptype', ['avg', 'max', 'avgmax', 'catavgmax'])
def test_pooling_linear(ptype):
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(3)
    c, k = 256, 128
    m = tb.POOLINGS.get('PoolingLinear')(in_channels=c, out_channels=k, pooling_type=ptype)
    o = om.PoolingLinear(c, k, ptype)
    with torch.no_grad():
        o.fc.weight.copy_(_bf(torch.randn_like(o.fc.weight) * 0.05))
        o.fc.bias.copy_(torch.randn_like(o.fc.bias) * 0.1)
    m.load_state_dict(o.state_dict(), strict=True)
    m.cuda()
    assert m.out_channels == k
    x = _bf(torch.randn(8, c, 7, 7))
    xo = x.clone().requires_grad_(True)
    with om.amp_bf16():
        yo = o(xo)
        r = _bf(torch.randn_like(yo))
        (yo * r).sum().backward()
    xm = x.cuda().requires_grad_(True)
    ym = m(xm)
    (ym.float() * r.cuda()).sum().backward()
    assert rel_err(ym, yo) < 1e-2
    assert rel_err(xm.grad, xo.grad) < 1e-2
    assert rel_err(m.fc.weight.grad, o.fc.weight.grad) < 1e-2
    assert rel_err(m.fc.bias.grad, o.fc.bias.grad) < 1e-2
