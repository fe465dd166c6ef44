#!/usr/bin/env python
"""FLOPs of one C4 image (HRNet-W18 + HRNetSegmentationNeck + SegmentationHead @512^2, fwd + bwd) with torch's
FlopCounterMode on the CPU oracle (same convention as the other workloads in bench.py: 2 x MAC, the data
gradient of the first convolution excluded).  python scripts/count_flops_hrnet.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.utils.flop_counter import FlopCounterMode  # noqa: E402

from oracle import models as om  # noqa: E402

torch.manual_seed(0)
bb = om.hrnet('hrnet_w18')
task = om.SegmentationTask(bb, om.HRNetSegmentationNeck([18, 36, 72, 144]), om.SegmentationHead(270, 19)).train()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.randn(1, 3, size, size)
y = torch.randint(0, 19, (1, size, size))
with FlopCounterMode(display=False) as fc:
    out = task.forward_with_gt({'image': x, 'target': y})['prediction']
    loss = torch.nn.functional.cross_entropy(out, y)
    loss.backward()
total = fc.get_total_flops()
# x does not require grad, so autograd already skips the first conv's data gradient
print(f'HRNet-W18 seg @{size}: {total / 1e9:.2f} GFLOP per image (fwd + bwd)')
