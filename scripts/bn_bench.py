"""Micro-benchmark of the BatchNorm backward reduction at given (rows, C) shapes: CUDA events around the C-ABI call."""
import sys, torch
sys.path.insert(0, '.')
from torchok_b200._lib import lib
from torchok_b200.kernels import _p

L = lib()
dev = 'cuda'
shapes = [(12544, 2048), (50176, 1024), (200704, 512), (802816, 256), (802816, 64), (200704, 128), (50176, 256), (12544, 512)]
st = torch.cuda.current_stream().cuda_stream
for rows, c in shapes:
    g = torch.randn(rows, c, device=dev).bfloat16()
    y = torch.randn(rows, c, device=dev).bfloat16()
    bits = torch.randint(0, 255, (rows * c // 8,), dtype=torch.uint8, device=dev)
    small = torch.randn(4, c, device=dev).abs() + 0.5
    acc = torch.zeros(6, c, device=dev)
    gamma = torch.ones(c, device=dev)
    coefs = torch.empty(3, c, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)
    for mode in (1, 2):
        for fin in (1,):
            def run():
                if fin:
                    L.tok_bn_bwd_reduce2_finalize(rows, c, _p(g), None, _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                                                  _p(acc[2]), _p(acc[3]), _p(small[2]), _p(small[3]), _p(gamma), _p(coefs[0]),
                                                  _p(coefs[1]), _p(coefs[2]), _p(dg), _p(db), 1, cnt.data_ptr(), st)
                else:
                    L.tok_bn_bwd_reduce2(rows, c, _p(g), None, _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                                         _p(acc[2]), _p(acc[3]), st)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            byts = rows * c * 4 + (rows * c // 8 if mode == 2 else 0)
            print(f'rows {rows} C {c} mode {mode} fin {fin}: {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s', flush=True)
