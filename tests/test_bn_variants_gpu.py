"""The opt-in BatchNorm kernel variants (kept off because they measured slower: torchok_b200/kernels.py notes) must still
be CORRECT: each is compared with the default kernel sequence it replaces, on the same operands, including a channel
count below the pitch (nn.BatchNorm2d(18) at pitch 24: the *_cv rule — pad lanes get scale = shift = 0 and the parameter
buffers are never touched past c_valid) — and the default *_cv kernels themselves against a plain fp32 restatement of
torch.nn.BatchNorm2d(training) + ReLU and its autograd (torchok/models/modules/bricks/convbnact.py:44-53).

  tok_bn_apply_chain / tok_bn_apply_bits_chain   == tok_bn_finalize_train_cv + tok_bn_apply / tok_bn_apply_bits
  tok_bn_bwd_fused_cv                            == tok_bn_bwd_reduce2_finalize_cv + tok_bn_bwd_apply2
Bars: bit-identical bf16 outputs (same fp32 arithmetic, same rounding), 1e-6 relative on the fp32 side results."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(rows, c, cv, seed=0):
    from torchok_b200.kernels import _p  # noqa: F401
    g = torch.Generator(device='cuda').manual_seed(seed)
    dev = 'cuda'
    y = torch.randn(rows, c, device=dev, generator=g).bfloat16()
    y[:, cv:] = 0                     # pad lanes of a real layer are exactly zero
    res = torch.randn(rows, c, device=dev, generator=g).bfloat16()
    res[:, cv:] = 0
    gamma = (torch.rand(cv, device=dev, generator=g) + 0.5)
    beta = torch.randn(cv, device=dev, generator=g) * 0.1
    return y, res, gamma, beta


def _sums(y):
    yf = y.float()
    return yf.sum(0).contiguous(), (yf * yf).sum(0).contiguous()


@pytest.mark.parametrize('rows,c,cv', [(4096, 64, 64), (3000, 24, 18), (777, 40, 36), (50176, 256, 256)])
@pytest.mark.parametrize('with_res', [False, True])
def test_forward_finalize_apply_variants(rows, c, cv, with_res):
    from torchok_b200._lib import lib
    from torchok_b200.kernels import _p, _st
    L = lib()
    y, res, gamma, beta = _setup(rows, c, cv)
    st = _st()
    eps, mom = 1e-5, 0.1

    def run(variant):
        s, q = _sums(y)
        rm, rv = torch.zeros(cv, device='cuda'), torch.ones(cv, device='cuda')
        small = torch.zeros(4, c, device='cuda')
        out = torch.empty_like(y)
        bits = torch.zeros(rows * c // 8, dtype=torch.uint8, device='cuda')
        r = res if with_res else None
        if variant == 'default':
            L.tok_bn_finalize_train_cv(c, cv, float(rows), _p(s), _p(q), _p(gamma), _p(beta), eps, mom, _p(rm), _p(rv),
                                       _p(small[0]), _p(small[1]), _p(small[2]), _p(small[3]), st)
            if with_res:
                L.tok_bn_apply_bits(rows, c, _p(y), _p(small[0]), _p(small[1]), _p(r), _p(out), _p(bits), st)
            else:
                L.tok_bn_apply(rows, c, _p(y), _p(small[0]), _p(small[1]), None, 1, _p(out), st)
        else:
            if not with_res and not L.tok_bn_apply_train_supported(rows, c):
                pytest.skip('grid stride not a multiple of the channel-vector count for this shape')
            stale = torch.full((2, 16), 7.0, device='cuda')     # the accumulators a previous chain launch left behind
            if with_res:
                L.tok_bn_apply_bits_chain(rows, c, cv, _p(y), _p(s), _p(q), _p(gamma), _p(beta), eps, mom, _p(rm), _p(rv),
                                          _p(small[0]), _p(small[1]), _p(small[2]), _p(small[3]), _p(stale), 32, _p(r),
                                          _p(out), _p(bits), st)
            else:
                L.tok_bn_apply_chain(rows, c, cv, _p(y), _p(s), _p(q), _p(gamma), _p(beta), eps, mom, _p(rm), _p(rv),
                                     _p(small[0]), _p(small[1]), _p(small[2]), _p(small[3]), _p(stale), 32, None, 1,
                                     _p(out), st)
            torch.cuda.synchronize()
            assert float(stale.abs().max()) == 0.0      # zeroed by CTA 0 of the chain launch
            assert float(s.abs().max()) > 0.0           # ... while its own sums stay for the next chain launch
        torch.cuda.synchronize()
        return out, bits, small, rm, rv

    o0, b0, sm0, rm0, rv0 = run('default')
    o1, b1, sm1, rm1, rv1 = run('chain')
    assert torch.equal(o0, o1)
    if with_res:
        assert torch.equal(b0, b1)
    for a, b in ((sm0, sm1), (rm0, rm1), (rv0, rv1)):
        assert float((a - b).abs().max()) <= 1e-6 * float(a.abs().max() + 1e-12)
    # against torch.nn.BatchNorm2d(training) semantics in fp32
    yf = y.float()[:, :cv]
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    ref = (yf - mean) * torch.rsqrt(var + eps) * gamma + beta
    if with_res:
        ref = ref + res.float()[:, :cv]
    ref = torch.relu(ref)
    assert float((o0.float()[:, :cv] - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    assert float(o0[:, cv:].float().abs().max()) == 0.0 if cv < c and not with_res else True
    assert float((rm0 - mom * mean).abs().max()) <= 1e-5 * float(mean.abs().max() + 1e-6) + 1e-7
    unb = yf.var(0, unbiased=True)
    assert float((rv0 - ((1 - mom) + mom * unb)).abs().max()) <= 1e-5 * float(unb.abs().max())


@pytest.mark.parametrize('rows,c,cv', [(4096, 64, 64), (3000, 24, 18), (12544, 512, 512), (50176, 256, 256)])
@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('want_dres', [False, True])
def test_backward_fused_equals_reduce_then_apply(rows, c, cv, mode, want_dres):
    from torchok_b200._lib import lib
    from torchok_b200.kernels import _p, _st
    L = lib()
    y, _, gamma, _ = _setup(rows, c, cv, seed=1)
    gen = torch.Generator(device='cuda').manual_seed(2)
    g = torch.randn(rows, c, device='cuda', generator=gen).bfloat16()
    g[:, cv:] = 0
    bits = torch.randint(0, 256, (rows * c // 8,), dtype=torch.uint8, device='cuda', generator=gen)
    yf = y.float()
    mean = torch.zeros(c, device='cuda')
    invstd = torch.ones(c, device='cuda')
    mean[:cv] = yf[:, :cv].mean(0)
    invstd[:cv] = torch.rsqrt(yf[:, :cv].var(0, unbiased=False) + 1e-5)
    scale = torch.zeros(c, device='cuda')
    shift = torch.zeros(c, device='cuda')
    scale[:cv] = gamma * invstd[:cv]
    shift[:cv] = -mean[:cv] * scale[:cv]
    st = _st()

    def run(fused):
        acc = torch.zeros(2, c, device='cuda')
        coefs = torch.zeros(3, c, device='cuda')
        dgam, dbet = torch.zeros(cv, device='cuda'), torch.zeros(cv, device='cuda')
        words = torch.zeros(4, dtype=torch.int32, device='cuda')
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if want_dres else None
        if fused:
            for _ in range(2):     # twice: the release word toggles, the ticket resets, the sums come back zeroed
                dgam.zero_()
                dbet.zero_()
                L.tok_bn_bwd_fused_cv(rows, c, cv, _p(g), None, _p(y), mode, _p(bits), _p(scale), _p(shift), _p(acc[0]),
                                      _p(acc[1]), _p(mean), _p(invstd), _p(gamma), _p(coefs[0]), _p(coefs[1]), _p(coefs[2]),
                                      _p(dgam), _p(dbet), 1, words.data_ptr(), words.data_ptr() + 4, _p(dy), _p(dres), st)
        else:
            L.tok_bn_bwd_reduce2_finalize_cv(rows, c, cv, _p(g), None, _p(y), mode, _p(bits), _p(scale), _p(shift),
                                             _p(acc[0]), _p(acc[1]), _p(mean), _p(invstd), _p(gamma), _p(coefs[0]),
                                             _p(coefs[1]), _p(coefs[2]), _p(dgam), _p(dbet), 1, words.data_ptr(), st)
            L.tok_bn_bwd_apply2(rows, c, _p(g), None, _p(y), mode, _p(bits), _p(scale), _p(shift), _p(coefs[0]),
                                _p(coefs[1]), _p(coefs[2]), _p(dy), _p(dres), st)
        torch.cuda.synchronize()
        assert float(acc.abs().max()) == 0.0          # accumulators handed back zeroed
        return dy, dres, coefs, dgam, dbet

    d0, r0, c0, ga0, be0 = run(False)
    d1, r1, c1, ga1, be1 = run(True)
    # the sums are accumulated with fp32 atomics in launch-dependent order: coefficients agree to fp32 round-off and
    # the bf16 results to one ulp of the tensor maximum
    for a, b in ((c0, c1), (ga0, ga1), (be0, be1)):
        assert float((a - b).abs().max()) <= 2e-5 * float(a.abs().max() + 1e-12)
    assert float((d0.float() - d1.float()).abs().max()) <= 8e-3 * float(d0.float().abs().max())
    if want_dres:
        assert torch.equal(r0, r1)
    # fp32 restatement of BatchNorm + ReLU backward for the valid channels
    gf = g.float()[:, :cv]
    if mode == 1:
        gf = gf * ((yf[:, :cv] * scale[:cv] + shift[:cv]) > 0)
    elif mode == 2:
        b = bits.view(rows, c // 8)
        m = torch.stack([(b >> j) & 1 for j in range(8)], dim=-1).reshape(rows, c)[:, :cv]
        gf = gf * m
    xhat = (yf[:, :cv] - mean[:cv]) * invstd[:cv]
    dbeta = gf.sum(0)
    dgamma = (gf * xhat).sum(0)
    ref = gamma * invstd[:cv] * (gf - dbeta / rows - xhat * dgamma / rows)
    assert float((d0.float()[:, :cv] - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    assert float((ga0 - dgamma).abs().max()) <= 1e-3 * float(dgamma.abs().max())
    assert float((be0 - dbeta).abs().max()) <= 1e-3 * float(dbeta.abs().max() + 1e-6)
