"""Large-shape parity of the persistent tcgen05 conv kernel (ResNet-50 layer1 / layer2 shapes at bs128-256): fprop, the
BatchNorm column sums of its epilogue and dgrad against torch's fp32 convolution on bf16-representable operands.  These
sizes reach code paths the small unit-test shapes do not: >= 8 m-tiles per CTA (the weight-resident variant), several
n-tiles per CTA, the 128x256 tile.  (A parity-aliasing race in the weight-resident variant was found by exactly this
check: scripts/check_big_conv.py.)  Bar: 1e-2 of the tensor maximum (north_star, bf16); column sums 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,c,hw,k,r', [(128, 64, 56, 64, 3), (128, 64, 56, 256, 1), (128, 256, 56, 64, 1),
                                        (128, 128, 28, 512, 1), (256, 128, 28, 128, 3), (64, 64, 56, 64, 1),
                                        (256, 256, 14, 1024, 1), (256, 512, 7, 512, 3)])
def test_conv_fprop_dgrad_large_shapes(n, c, hw, k, r):
    from torchok_b200 import kernels as K
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        dev = torch.device('cuda')
        torch.manual_seed(c + k)
        pad = r // 2
        x = torch.randn(n, c, hw, hw, device=dev).to(torch.bfloat16)
        w = (torch.randn(k, c, r, r, device=dev) / (c * r * r) ** 0.5).to(torch.bfloat16)
        d, p, q = K.conv_desc(n, hw, hw, c, k, r, r, 1, pad, 1)
        xn, wk = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
        y = torch.empty(n, p, q, k, device=dev, dtype=torch.bfloat16)
        stats = torch.zeros(2, k, device=dev)
        K.conv_fprop(d, xn, wk, y, stats)
        ref = F.conv2d(x.float(), w.float(), padding=pad).permute(0, 2, 3, 1)
        assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1e-2
        col = y.float().sum((0, 1, 2))
        assert float((stats[0] - col).abs().max() / col.abs().max()) < 1e-3
        sq = (y.float() ** 2).sum((0, 1, 2))
        assert float((stats[1] - sq).abs().max() / sq.abs().max()) < 1e-3
        dy = torch.randn(n, p, q, k, device=dev).to(torch.bfloat16)
        dx = torch.empty(n, hw, hw, c, device=dev, dtype=torch.bfloat16)
        K.conv_dgrad(d, dy, wk, dx)
        refd = torch.nn.grad.conv2d_input((n, c, hw, hw), w.float(), dy.permute(0, 3, 1, 2).float(),
                                          padding=pad).permute(0, 2, 3, 1)
        assert float((dx.float() - refd).abs().max() / refd.abs().max()) < 1e-2
    finally:
        torch.backends.cudnn.allow_tf32 = prev
