#!/usr/bin/env python
"""One halo-kernel launch per direction for a given shape (ncu target): python scripts/halo_one.py n c h w k [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchok_b200 import kernels as K  # noqa: E402

n, c, h, w_, k = [int(v) for v in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dev = torch.device('cuda')
x = torch.randn(n, h, w_, c, device=dev).to(torch.bfloat16)
wk = (torch.randn(k, 3, 3, c, device=dev) / (c * 9) ** 0.5).to(torch.bfloat16)
dy = torch.randn(n, h, w_, k, device=dev).to(torch.bfloat16)
d, p, q = K.conv_desc(n, h, w_, c, k, 3, 3, 1, 1, 1)
y = torch.empty(n, h, w_, k, device=dev, dtype=torch.bfloat16)
dx = torch.empty(n, h, w_, c, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(2, k, device=dev)
for _ in range(reps):
    K.conv_fprop(d, x, wk, y, stats)
    K.conv_dgrad(d, dy, wk, dx)
torch.cuda.synchronize()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record()
for _ in range(reps):
    K.conv_fprop(d, x, wk, y, stats)
e1.record()
for _ in range(reps):
    K.conv_dgrad(d, dy, wk, dx)
e2.record()
torch.cuda.synchronize()
print(f'n{n} {c}x{h}x{w_}->{k}: fprop {e0.elapsed_time(e1) * 1e3 / reps:.1f} us dgrad {e1.elapsed_time(e2) * 1e3 / reps:.1f} us')
