"""Parity of the halo 3x3 kernels (csrc/tok_conv3.cu: forward + BatchNorm sums, data gradient (+addend), weight
gradient) against torch's fp32 convolution on the CPU (small / ragged shapes) and on the GPU (BASELINE-size layers of
ResNet-50 and HRNet-W18), and bit-for-bit against the generic persistent kernel on the shapes both can run.

Covers what the reference's layers ask of a 3x3 / stride 1 / pad 1 Conv2d (timm BasicBlock / Bottleneck convs built by
torchok/models/backbones/resnet.py:363-405, HighResolutionModule branches of torchok/models/backbones/hrnet.py:140-192):
channel counts that are not multiples of 8 (18, 36: unpadded weights through tokConvDesc.wk / wc), widths that do not
divide the 128-row accumulator, single-row images, image heights that do not divide the row block, the widest supported
row (W + 2 = 256), two channel blocks per tap (C = 128).  Bars: 1e-2 of the tensor maximum for bf16 outputs
(north_star), 1e-3 for the fp32 column sums, 2e-3 for the fp32 weight gradient."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _refs(x, w, dy, device):
    """fp32 convolution, input gradient and weight gradient of NCHW operands on `device`."""
    x32, w32 = x.float().to(device), w.float().to(device)
    dy32 = dy.permute(0, 3, 1, 2).float().to(device)
    ref = F.conv2d(x32, w32, padding=1).permute(0, 2, 3, 1)
    refd = torch.nn.grad.conv2d_input(tuple(x.shape), w32, dy32, padding=1).permute(0, 2, 3, 1)
    refw = torch.nn.grad.conv2d_weight(x32, tuple(w.shape), dy32, padding=1).permute(0, 2, 3, 1)
    return ref.cuda(), refd.cuda(), refw.cuda()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _run(n, c, h, w_, k, ref_device, unpadded=False):
    from torchok_b200 import kernels as K
    from torchok_b200._lib import lib
    import ctypes as C
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        dev = torch.device('cuda')
        torch.manual_seed(c * 1000 + k + h)
        cp, kp = K.ceil8(c), K.ceil8(k)
        x = torch.randn(n, c, h, w_, device=dev).to(torch.bfloat16)
        w = (torch.randn(k, c, 3, 3, device=dev) / (c * 9) ** 0.5).to(torch.bfloat16)
        dy = torch.randn(n, h, w_, k, device=dev).to(torch.bfloat16)
        add = torch.randn(n, h, w_, c, device=dev).to(torch.bfloat16)
        # activations carry zero pad lanes up to the pitch; weights are padded unless the descriptor is direct
        xn = torch.zeros(n, h, w_, cp, device=dev, dtype=torch.bfloat16)
        xn[..., :c] = x.permute(0, 2, 3, 1)
        dyn = torch.zeros(n, h, w_, kp, device=dev, dtype=torch.bfloat16)
        dyn[..., :k] = dy
        addn = torch.zeros(n, h, w_, cp, device=dev, dtype=torch.bfloat16)
        addn[..., :c] = add
        if unpadded:
            d, p, q = K.conv_desc(n, h, w_, cp, kp, 3, 3, 1, 1, 1, k, c)
            assert lib().tok_conv_halo_caps(C.byref(d)) == 7
            wk = w.permute(0, 2, 3, 1).contiguous()
            dw = torch.zeros(k, 3, 3, c, device=dev)
        else:
            d, p, q = K.conv_desc(n, h, w_, cp, kp, 3, 3, 1, 1, 1)
            wk = torch.zeros(kp, 3, 3, cp, device=dev, dtype=torch.bfloat16)
            wk[:k, :, :, :c] = w.permute(0, 2, 3, 1)
            dw = torch.zeros(kp, 3, 3, cp, device=dev)
        y = torch.full((n, h, w_, kp), float('nan'), device=dev, dtype=torch.bfloat16)
        stats = torch.zeros(2, kp, device=dev)
        K.conv_fprop(d, xn, wk, y, stats)
        dx = torch.full((n, h, w_, cp), float('nan'), device=dev, dtype=torch.bfloat16)
        K.conv_dgrad(d, dyn, wk, dx)
        dxa = torch.full((n, h, w_, cp), float('nan'), device=dev, dtype=torch.bfloat16)
        K.conv_dgrad(d, dyn, wk, dxa, addn)
        K.conv_wgrad(d, xn, dyn, dw)
        torch.cuda.synchronize()
        ref, refd, refw = _refs(x, w, dy, ref_device)
        assert _rel(y[..., :k].float(), ref) < 1e-2
        assert _rel(dx[..., :c].float(), refd) < 1e-2
        assert _rel(dxa[..., :c].float(), refd + add.float()) < 1e-2
        assert _rel(dw[:k, :, :, :c], refw) < 2e-3
        # pad lanes stay exactly zero (the next layer's TMA reads them)
        if kp != k:
            assert float(y[..., k:].float().abs().max()) == 0.0
        if cp != c:
            assert float(dx[..., c:].float().abs().max()) == 0.0
        col = y.float().sum((0, 1, 2))
        sq = (y.float() ** 2).sum((0, 1, 2))
        assert float((stats[0] - col).abs().max() / col.abs().max().clamp_min(1e-6)) < 1e-3
        assert float((stats[1] - sq).abs().max() / sq.abs().max()) < 1e-3
        return d, xn, wk, dyn, y, dx
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize('n,c,h,w_,k', [
    (2, 64, 8, 8, 64), (3, 24, 9, 13, 24), (2, 40, 17, 30, 40), (2, 128, 7, 7, 128), (1, 8, 1, 1, 8),
    (4, 32, 20, 128, 24), (2, 24, 5, 254, 40), (2, 16, 3, 3, 8), (1, 48, 1, 37, 64), (2, 128, 28, 28, 128),
    (1, 56, 33, 9, 48),
])
def test_halo_conv_small_and_ragged_vs_cpu_fp32(n, c, h, w_, k):
    _run(n, c, h, w_, k, 'cpu')


@pytest.mark.parametrize('n,c,h,w_,k', [(2, 18, 16, 16, 18), (3, 36, 9, 21, 36), (2, 18, 32, 128, 36), (1, 36, 7, 5, 18),
                                        (2, 20, 11, 11, 44)])
def test_halo_conv_unpadded_weights_vs_cpu_fp32(n, c, h, w_, k):
    """HRNet's 18 / 36-channel branches: activations at pitch 24 / 40, weights and weight gradient at their real size."""
    _run(n, c, h, w_, k, 'cpu', unpadded=True)
    _run(n, c, h, w_, k, 'cpu', unpadded=False)


@pytest.mark.parametrize('n,c,h,w_,k', [(256, 64, 56, 56, 64), (256, 128, 28, 28, 128), (32, 18, 128, 128, 18),
                                        (32, 36, 64, 64, 36)])
def test_halo_conv_baseline_size_layers(n, c, h, w_, k):
    """ResNet-50 bs256 layer1 / layer2 3x3s and HRNet-W18 bs32 @512 branch convs, against torch's fp32 GPU convolution."""
    _run(n, c, h, w_, k, 'cuda', unpadded=(c % 8 != 0))


def test_halo_equals_generic_kernel_bit_for_bit():
    """Same operands through the generic persistent kernel (TOK_CONV_HALO=0): the single-k-block layers accumulate the
    same products in fp32 and round once, so outputs are identical, not merely close."""
    from torchok_b200 import kernels as K
    d, xn, wk, dyn, y, dx = _run(4, 64, 24, 24, 64, 'cpu')
    os.environ['TOK_CONV_HALO'] = '0'
    try:
        y0 = torch.empty_like(y)
        K.conv_fprop(d, xn, wk, y0)
        dx0 = torch.empty_like(dx)
        K.conv_dgrad(d, dyn, wk, dx0)
        torch.cuda.synchronize()
    finally:
        del os.environ['TOK_CONV_HALO']
    assert torch.equal(y, y0)
    assert torch.equal(dx, dx0)


@pytest.mark.parametrize('n,c,h,w_,k,r', [(4, 64, 24, 24, 64, 3), (2, 24, 17, 31, 40, 3), (8, 256, 14, 14, 64, 1),
                                           (4, 1024, 7, 7, 256, 1), (4, 256, 14, 14, 256, 3)])
def test_dgrad_masked_addend_equals_explicit_product(n, c, h, w_, k, r):
    """tok_conv_dgrad_masked (dx = dgrad(dy) + addend * [bit]) against tok_conv_dgrad fed the materialised product — the
    residual-gradient path of a block (timm `x += shortcut; x = act(x)`, resnet.py:363-405) on the halo kernel and on the
    generic kernel (1x1, 3x3 with Cin > 128).  Same fp32 sums, same single rounding: bit-identical."""
    import ctypes as C
    from torchok_b200 import kernels as K
    from torchok_b200._lib import lib
    dev = torch.device('cuda')
    torch.manual_seed(n * 100 + c)
    pad = r // 2
    d, p, q = K.conv_desc(n, h, w_, c, k, r, r, 1, pad, 1)
    assert lib().tok_conv_dgrad_masked_supported(C.byref(d)) == 1
    wk = (torch.randn(k, r, r, c, device=dev) / (c * r * r) ** 0.5).to(torch.bfloat16)
    dy = torch.randn(n, h, w_, k, device=dev).to(torch.bfloat16)
    add = torch.randn(n, h, w_, c, device=dev).to(torch.bfloat16)
    bits = torch.randint(0, 256, (n * h * w_ * c // 8,), dtype=torch.uint8, device=dev)
    mask = torch.stack([(bits >> j) & 1 for j in range(8)], dim=-1).reshape(n, h, w_, c).to(torch.bfloat16)
    ref = torch.empty(n, h, w_, c, device=dev, dtype=torch.bfloat16)
    K.conv_dgrad(d, dy, wk, ref, (add * mask).contiguous())
    got = torch.full_like(ref, float('nan'))
    K.conv_dgrad(d, dy, wk, got, add, bits)
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
