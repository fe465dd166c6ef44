#!/usr/bin/env python
"""Parity + timing A/B of the halo 3x3 kernel (tok_conv3.cu) against the generic persistent kernel and torch's fp32
convolution: fprop (+BatchNorm statistics), dgrad, dgrad + addend, on ResNet / HRNet 3x3 shapes incl. ragged ones.
python scripts/check_halo_conv.py [quick]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from torchok_b200 import kernels as K  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


shapes = [  # n, cin, h, w, cout
    (2, 64, 8, 8, 64), (3, 24, 9, 13, 24), (2, 40, 17, 30, 40), (2, 128, 7, 7, 128), (2, 64, 56, 56, 64),
    (4, 32, 20, 128, 24), (2, 24, 5, 254, 40), (1, 72, 12, 12, 72), (2, 128, 28, 28, 128), (2, 16, 3, 3, 8),
]
if not quick:
    shapes += [(256, 64, 56, 56, 64), (256, 128, 28, 28, 128), (32, 24, 128, 128, 24), (32, 40, 64, 64, 40),
               (32, 72, 32, 32, 72)]
ok = True
for (n, c, h, w_, k) in shapes:
    torch.manual_seed(c * 1000 + k + h)
    x = torch.randn(n, c, h, w_, device=dev).to(torch.bfloat16)
    w = (torch.randn(k, c, 3, 3, device=dev) / (c * 9) ** 0.5).to(torch.bfloat16)
    d, p, q = K.conv_desc(n, h, w_, c, k, 3, 3, 1, 1, 1)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = w.permute(0, 2, 3, 1).contiguous()
    dy = torch.randn(n, h, w_, k, device=dev).to(torch.bfloat16)
    add = torch.randn(n, h, w_, c, device=dev).to(torch.bfloat16)
    res = {}
    for halo in ('0', '1'):
        os.environ['TOK_CONV_HALO'] = halo
        y = torch.full((n, h, w_, k), float('nan'), device=dev, dtype=torch.bfloat16)
        stats = torch.zeros(2, k, device=dev)
        K.conv_fprop(d, xn, wk, y, stats)
        dx = torch.full((n, h, w_, c), float('nan'), device=dev, dtype=torch.bfloat16)
        K.conv_dgrad(d, dy, wk, dx)
        dxa = torch.full((n, h, w_, c), float('nan'), device=dev, dtype=torch.bfloat16)
        K.conv_dgrad(d, dy, wk, dxa, add)
        torch.cuda.synchronize()
        dw = torch.zeros(k, 3, 3, c, device=dev)
        K.conv_wgrad(d, xn, dy, dw)
        dw_keep = dw.clone()
        t_w = timed(lambda: K.conv_wgrad(d, xn, dy, dw))
        t_f = timed(lambda: K.conv_fprop(d, xn, wk, y, stats))
        t_d = timed(lambda: K.conv_dgrad(d, dy, wk, dx))
        stats.zero_()
        K.conv_fprop(d, xn, wk, y, stats)
        res[halo] = (y, stats.clone(), dx, dxa, t_f, t_d, dw_keep, t_w)
    ref = F.conv2d(x.float(), w.float(), padding=1).permute(0, 2, 3, 1)
    refd = torch.nn.grad.conv2d_input((n, c, h, w_), w.float(), dy.permute(0, 3, 1, 2).float(),
                                      padding=1).permute(0, 2, 3, 1)
    y, stats, dx, dxa, t_f, t_d, dw, t_w = res['1']
    y0, stats0, dx0, dxa0, t_f0, t_d0, dw0, t_w0 = res['0']
    e_f = float((y.float() - ref).abs().max() / ref.abs().max())
    e_d = float((dx.float() - refd).abs().max() / refd.abs().max())
    e_a = float((dxa.float() - (refd + add.float())).abs().max() / (refd + add.float()).abs().max())
    ysum = y.float().sum((0, 1, 2))
    ysq = (y.float() ** 2).sum((0, 1, 2))
    e_s = float((stats[0] - ysum).abs().max() / (ysum.abs().max() + 1e-6))
    e_q = float((stats[1] - ysq).abs().max() / ysq.abs().max())
    refw = torch.nn.grad.conv2d_weight(x.float(), (k, c, 3, 3), dy.permute(0, 3, 1, 2).float(), padding=1).permute(0, 2, 3, 1)
    e_w = float((dw - refw).abs().max() / refw.abs().max())
    e_w0 = float((dw0 - refw).abs().max() / refw.abs().max())
    same_f = float((y.float() - y0.float()).abs().max())
    same_d = float((dx.float() - dx0.float()).abs().max())
    good = e_f < 1e-2 and e_d < 1e-2 and e_a < 1e-2 and e_s < 1e-3 and e_q < 1e-3 and e_w < 2e-3
    ok &= good
    print(f'{"PASS" if good else "FAIL"} n{n} {c}x{h}x{w_}->{k}: fprop {e_f:.1e} dgrad {e_d:.1e} dgrad+add {e_a:.1e} '
          f'sum {e_s:.1e} sq {e_q:.1e} wgrad {e_w:.1e} (generic {e_w0:.1e}) | vs generic: fprop {same_f:.1e} dgrad {same_d:.1e} | us halo {t_f:.1f}/{t_d:.1f} '
          f'generic {t_f0:.1f}/{t_d0:.1f} | wgrad us {t_w:.1f} vs {t_w0:.1f}', flush=True)
print('HALO CONV CHECK', 'OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
