#!/usr/bin/env python
"""ncu target: fprop (+stats), dgrad and wgrad of one 3x3 layer, twice.  python scripts/halo_one_w.py n c h w k"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchok_b200 import kernels as K  # noqa: E402

n, c, h, w_, k = [int(v) for v in sys.argv[1:6]]
dev = torch.device('cuda')
x = torch.randn(n, h, w_, c, device=dev).to(torch.bfloat16)
wk = (torch.randn(k, 3, 3, c, device=dev) / (c * 9) ** 0.5).to(torch.bfloat16)
dy = torch.randn(n, h, w_, k, device=dev).to(torch.bfloat16)
d, p, q = K.conv_desc(n, h, w_, c, k, 3, 3, 1, 1, 1)
y = torch.empty(n, h, w_, k, device=dev, dtype=torch.bfloat16)
dx = torch.empty(n, h, w_, c, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(k, 3, 3, c, device=dev)
stats = torch.zeros(2, k, device=dev)
for _ in range(2):
    K.conv_fprop(d, x, wk, y, stats)
    K.conv_dgrad(d, dy, wk, dx)
    K.conv_wgrad(d, x, dy, dw)
torch.cuda.synchronize()
