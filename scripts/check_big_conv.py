#!/usr/bin/env python
"""Large-shape check of the persistent conv kernel (the weight-resident variant only engages at >= 8 m-tiles per CTA,
which the unit tests' shapes do not reach): fprop / dgrad of ResNet-50's layer1 / layer2 shapes against torch's fp32
convolution on the same bf16-representable operands.  python scripts/check_big_conv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from torchok_b200 import kernels as K  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
ok = True
for (n, c, hw, k, r, stride) in [(128, 64, 56, 64, 3, 1), (128, 64, 56, 256, 1, 1), (128, 256, 56, 64, 1, 1),
                                 (128, 128, 28, 512, 1, 1), (256, 128, 28, 128, 3, 1), (64, 64, 56, 64, 1, 1)]:
    torch.manual_seed(c + k)
    pad = r // 2
    x = torch.randn(n, c, hw, hw, device=dev).to(torch.bfloat16)
    w = (torch.randn(k, c, r, r, device=dev) / (c * r * r) ** 0.5).to(torch.bfloat16)
    d, p, q = K.conv_desc(n, hw, hw, c, k, r, r, stride, pad, 1)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = w.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(n, p, q, k, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, k, device=dev)
    K.conv_fprop(d, xn, wk, y, stats)
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    e_f = float((y.float() - ref).abs().max() / ref.abs().max())
    e_s = float((stats[0] - y.float().sum((0, 1, 2))).abs().max() / y.float().sum((0, 1, 2)).abs().max())
    dy = torch.randn(n, p, q, k, device=dev).to(torch.bfloat16)
    dx = torch.empty(n, hw, hw, c, device=dev, dtype=torch.bfloat16)
    K.conv_dgrad(d, dy, wk, dx)
    refd = torch.nn.grad.conv2d_input((n, c, hw, hw), w.float(), dy.permute(0, 3, 1, 2).float(), stride=stride,
                                      padding=pad).permute(0, 2, 3, 1)
    e_d = float((dx.float() - refd).abs().max() / refd.abs().max())
    good = e_f < 1e-2 and e_d < 1e-2 and e_s < 1e-3
    ok &= good
    print(f'{"PASS" if good else "FAIL"} n{n} {c}x{hw}x{hw}->{k} k{r}: fprop {e_f:.2e} stats {e_s:.2e} dgrad {e_d:.2e}', flush=True)
print('BIG CONV CHECK', 'OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
