"""ncu target: one BatchNorm backward reduce (+finalize) and one backward apply on a small L2-resident tensor."""
import sys, torch
sys.path.insert(0, '.')
from torchok_b200._lib import lib
from torchok_b200.kernels import _p
L = lib()
rows, c = int(sys.argv[1]), int(sys.argv[2])
dev = 'cuda'
g = torch.randn(rows, c, device=dev).bfloat16(); y = torch.randn(rows, c, device=dev).bfloat16()
bits = torch.randint(0, 255, (rows * c // 8,), dtype=torch.uint8, device=dev)
small = torch.randn(4, c, device=dev).abs() + 0.5; acc = torch.zeros(6, c, device=dev)
gamma = torch.ones(c, device=dev); coefs = torch.empty(3, c, device=dev)
dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev); cnt = torch.zeros(4, dtype=torch.int32, device=dev)
dy = torch.empty_like(g)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    L.tok_bn_bwd_reduce2_finalize(rows, c, _p(g), None, _p(y), 1, _p(bits), _p(small[0]), _p(small[1]), _p(acc[2]), _p(acc[3]),
                                  _p(small[2]), _p(small[3]), _p(gamma), _p(coefs[0]), _p(coefs[1]), _p(coefs[2]), _p(dg), _p(db), 1,
                                  cnt.data_ptr(), st)
    L.tok_bn_bwd_apply2(rows, c, _p(g), None, _p(y), 1, _p(bits), _p(small[0]), _p(small[1]), _p(coefs[0]), _p(coefs[1]),
                        _p(coefs[2]), _p(dy), None, st)
torch.cuda.synchronize()
