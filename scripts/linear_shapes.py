"""Swin-T linear layers at bs256 one by one: CUDA-event time of tok_linear_fwd / _dgrad / _dgrad_add / _wgrad per stage
against the HBM floor (bytes / 6.55 TB/s) and the tensor floor (FLOPs / 1.38 PFLOP/s)."""
import torch

from torchok_b200._lib import lib
from torchok_b200.kernels import _p, _st

L = lib()
B = 256


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(n):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    return min(ts)


for hw, c in [(56, 96), (28, 192), (14, 384), (7, 768)]:
    m = B * hw * hw
    for name, k, n in [('qkv', c, 3 * c), ('proj', c, c), ('fc1', c, 4 * c), ('fc2', 4 * c, c)]:
        x = torch.randn(m, k, device='cuda').bfloat16()
        w = (torch.randn(n, k, device='cuda') / k ** 0.5).bfloat16()
        b = torch.zeros(n, device='cuda')
        y = torch.empty(m, n, device='cuda', dtype=torch.bfloat16)
        dy = torch.randn(m, n, device='cuda').bfloat16()
        dx = torch.empty(m, k, device='cuda', dtype=torch.bfloat16)
        dw = torch.zeros(n, k, device='cuda')
        tf = timeit(lambda: L.tok_linear_fwd(m, n, k, _p(x), _p(w), _p(b), _p(y), _st()))
        tf0 = timeit(lambda: L.tok_linear_fwd(m, n, k, _p(x), _p(w), None, _p(y), _st()))
        td = timeit(lambda: L.tok_linear_dgrad(m, n, k, _p(dy), _p(w), _p(dx), _st()))
        tw = timeit(lambda: L.tok_linear_wgrad(m, n, k, _p(x), _p(dy), _p(dw), _st()))
        byt = 2.0 * m * (k + n)
        fl = 2.0 * m * n * k
        floor = max(byt / 6.55e12, fl / 1.38e15) * 1e6
        print(f'{hw:3d} {name:5s} M={m} K={k} N={n}: floor {floor:6.1f} us | fwd {tf:6.1f} ({floor / tf:.2f}) nobias {tf0:6.1f} '
              f'dgrad {td:6.1f} ({floor / td:.2f}) wgrad {tw:6.1f} ({floor / tw:.2f})')
