#!/usr/bin/env python
"""Launches bench.py's `roofline` kernel (tcgen05 implicit-GEMM conv, 3x3 256->256 @14x14, bs256) a few times so that
`ncu --set full -k regex:conv_fwd_persist --launch-skip 3 --launch-count 1` captures one warm launch of it.
The DRAM bytes of that capture are the `roofline.traffic` figure quoted in bench.py / profiles/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchok_b200 import kernels as K  # noqa: E402

n, h, c, k = 256, 14, 256, 256
dev = torch.device('cuda')
d, p, q = K.conv_desc(n, h, h, c, k, 3, 3, 1, 1, 1)
x = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
w = torch.randn(k, 3, 3, c, device=dev).to(torch.bfloat16)
y = torch.empty(n, p, q, k, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(2, k, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.zero_()
    K.conv_fprop(d, x, w, y, stats)
torch.cuda.synchronize()
