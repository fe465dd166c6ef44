#!/usr/bin/env python
"""Window-attention microbench / parity aid: times tok_window_attn_fwd / _bwd on one Swin stage geometry with CUDA
events and (with `check`) compares the tcgen05 backward against the CUDA-core backward (TOK_ATTN_BWD_CUDA_CORES=1, read
per call).  Usage: python scripts/attn_bench.py [stage 1-4] [batch] [check]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchok_b200._lib import lib  # noqa: E402
from torchok_b200.kernels import _p, _st  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
check = 'check' in sys.argv
H = 56 >> (stage - 1)
C = 96 << (stage - 1)
heads = 3 << (stage - 1)
ws = 7
shift = 3 if H > ws else 0
N = ws * ws
torch.manual_seed(0)
dev = torch.device('cuda')
qkv = torch.randn(B * H * H, 3 * C, device=dev).to(torch.bfloat16)
bias = (16 * torch.sigmoid(torch.randn(heads, N, N, device=dev))).contiguous()
ls = torch.full((heads,), 2.3026, device=dev)
out = torch.empty(B * H * H, C, device=dev, dtype=torch.bfloat16)
g = torch.randn(B * H * H, C, device=dev).to(torch.bfloat16)
L = lib()


def fwd():
    L.tok_window_attn_fwd(B, H, H, C, heads, ws, shift, _p(qkv), _p(ls), _p(bias), _p(out), _st())


def bwd(mode):
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros_like(bias)
    dls = torch.zeros_like(ls)
    os.environ['TOK_ATTN_BWD_CUDA_CORES'] = '1' if mode else '0'
    L.tok_window_attn_bwd(B, H, H, C, heads, ws, shift, _p(qkv), _p(ls), _p(bias), _p(g), _p(dqkv), _p(dbias), _p(dls),
                          None, _st())
    return dqkv, dbias, dls


def timed(fn, it=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


print(f'stage {stage}: B={B} H=W={H} C={C} heads={heads} ws={ws} shift={shift}')
print(f'  fwd        {timed(fwd):9.1f} us')
print(f'  bwd (tc)   {timed(lambda: bwd(0)):9.1f} us')
if check:
    print(f'  bwd (cuda) {timed(lambda: bwd(1), 2):9.1f} us')
    a, b = bwd(0), bwd(1)
    torch.cuda.synchronize()
    for name, x, y in zip(('dqkv', 'dbias', 'dlogit_scale'), a, b):
        x, y = x.float(), y.float()
        print(f'  {name:13s} rel_l2 = {float((x - y).norm() / y.norm()):.3e}   max|ref| = {float(y.abs().max()):.3e}   '
              f'max|diff| = {float((x - y).abs().max()):.3e}')
    for part, nm in enumerate(('dq', 'dk', 'dv')):
        x, y = a[0].float()[:, part * C:(part + 1) * C], b[0].float()[:, part * C:(part + 1) * C]
        print(f'  {nm:13s} rel_l2 = {float((x - y).norm() / y.norm()):.3e}')
