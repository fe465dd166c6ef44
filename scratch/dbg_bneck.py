import torch, sys
sys.path.insert(0, '/root/repo')
from oracle import models as om
from torchok_b200.models.backbones import resnet as pr
torch.manual_seed(256+64+1)
def bf(t): return t.to(torch.bfloat16).float()
inpl, planes, stride, hw, n = 256, 64, 1, 14, 4
o = om.Bottleneck(inpl, planes, stride, None); om.dedegenerate_(o, 5)
with torch.no_grad():
    for mod in o.modules():
        if isinstance(mod, torch.nn.Conv2d): mod.weight.copy_(bf(mod.weight))
m = pr.Bottleneck(inpl, planes, stride, None); m.load_state_dict(o.state_dict()); m.cuda()
x = bf(torch.randn(n, inpl, hw, hw))
xo, xm = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
with om.amp_bf16():
    yo = o(xo); r = bf(torch.randn_like(yo)); (yo*r).sum().backward()
ym = m(xm); (ym.float()*r.cuda()).sum().backward()
d = (xm.grad.float().cpu() - xo.grad).abs()
mx = xo.grad.abs().max()
bad = (d > 1e-2*mx)
print('bad count', bad.sum().item(), 'of', bad.numel())
idx = bad.nonzero()
print('by n:', torch.bincount(idx[:,0], minlength=n).tolist())
print('by h:', torch.bincount(idx[:,2], minlength=hw).tolist())
print('by w:', torch.bincount(idx[:,3], minlength=hw).tolist())
cc = torch.bincount(idx[:,1], minlength=inpl)
print('by c (nonzero):', [(i,int(v)) for i,v in enumerate(cc.tolist()) if v][:40])
# mask agreement of block outputs
mo = (yo>0); mm = (ym.float().cpu()>0)
print('mask mismatches', (mo!=mm).sum().item())
print('fwd diff max', (ym.float().cpu()-yo).abs().max().item())
# is the error equal to r at those positions (a g-flip)?
print('sample errors', d[bad][:10].tolist(), 'r there', r[bad][:10].tolist())
