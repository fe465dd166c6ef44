"""Window-attention forward alone (Swin-T stage shapes at bs256): CUDA-event time and, with TOK_ATTN_PROFILE=1, the clock64
phase sums of CTA 0 (threads 0 and 128).  Bring-up aid for csrc/tok_swin.cu: window_attn_fwd_tc2_kernel."""
import ctypes as C
import sys

import torch

from torchok_b200._lib import lib
from torchok_b200.kernels import _p, _st

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = [(56, 96, 3), (28, 192, 6), (14, 384, 12), (7, 768, 24)]
L = lib()
for hw, c, heads in shapes:
    for shift in (0, 3 if hw > 7 else 0):
        qkv = torch.randn(B * hw * hw, 3 * c, device='cuda').bfloat16()
        bias = torch.randn(heads, 49, 49, device='cuda')
        ls = torch.full((heads,), 2.3, device='cuda')
        out = torch.empty(B * hw * hw, c, device='cuda', dtype=torch.bfloat16)
        for _ in range(3):
            L.tok_window_attn_fwd(B, hw, hw, c, heads, 7, shift, _p(qkv), _p(ls), _p(bias), _p(out), _st())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.tok_window_attn_fwd(B, hw, hw, c, heads, 7, shift, _p(qkv), _p(ls), _p(bias), _p(out), _st())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 5
        pairs = B * (hw // 7) ** 2 // 2
        groups = min(296 // heads, pairs)
        iters = -(-pairs // groups)
        gout = torch.randn(B * hw * hw, c, device='cuda').bfloat16()
        dqkv = torch.empty_like(qkv)
        dbias = torch.zeros_like(bias)
        dls = torch.zeros_like(ls)
        col = torch.zeros(3 * c, device='cuda')
        for _ in range(2):
            L.tok_window_attn_bwd(B, hw, hw, c, heads, 7, shift, _p(qkv), _p(ls), _p(bias), _p(gout), _p(dqkv), _p(dbias),
                                  _p(dls), _p(col), _st())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            L.tok_window_attn_bwd(B, hw, hw, c, heads, 7, shift, _p(qkv), _p(ls), _p(bias), _p(gout), _p(dqkv), _p(dbias),
                                  _p(dls), _p(col), _st())
        e1.record()
        torch.cuda.synchronize()
        us_b = e0.elapsed_time(e1) * 1000 / 5
        buf = (C.c_longlong * 64)()
        n = L.tok_debug_attn_profile(buf, 64)
        line = f'hw={hw} C={c} heads={heads} shift={shift}: {us:.1f} us, {iters} pairs/CTA, {us * 1000 / iters:.0f} ns/pair; bwd {us_b:.1f} us'
        print(line)
        if n:
            for th in range(2):
                v = [buf[th * 14 + i] / iters for i in range(14)]
                print('   thread', th * 128, ' '.join(f'{x:.0f}' for x in v), ' total', f'{sum(v):.0f}', 'clk/pair')
