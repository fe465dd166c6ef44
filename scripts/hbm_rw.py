"""HBM bandwidth by access mix: write-only (fill), read-only (sum), copy — 2 GiB buffers, CUDA events, best of 5."""
import torch
dev = 'cuda'
n = 1 << 30   # bf16 elements = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)


def best(fn, it=5):
    ts = []
    for _ in range(it):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


gb = a.numel() * 2 / 1e9
t = best(lambda: b.fill_(1.0)); print(f'write-only  fill_      {gb / t * 1e3:8.0f} GB/s')
t = best(lambda: b.zero_()); print(f'write-only  zero_      {gb / t * 1e3:8.0f} GB/s')
t = best(lambda: a.float().sum() if False else torch.sum(a.view(torch.int16))); print(f'read-only   sum        {gb / t * 1e3:8.0f} GB/s')
t = best(lambda: b.copy_(a)); print(f'copy        r+w        {2 * gb / t * 1e3:8.0f} GB/s')
c = torch.empty(n // 4, dtype=torch.bfloat16, device=dev).normal_()
# 1 read : 4 writes (an expansion 1x1: reads M x 64, writes M x 256)
t = best(lambda: b.view(4, -1).copy_(c.view(1, -1).expand(4, -1))); print(f'1r:4w       expand     {(gb + gb / 4) / t * 1e3:8.0f} GB/s')
