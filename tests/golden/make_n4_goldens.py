"""Golden vectors for the SURVEY §8f N4 parts whose kernels are not written yet (OCRSegmentationHead, UnetNeck),
produced by the REFERENCE's own files executed by path (same mechanism as make_reference_goldens.py):

    torchok/models/heads/segmentation/ocr.py, torchok/models/necks/segmentation/unet.py (+ modules/blocks/scse.py,
    modules/bricks/convbnact.py, models/base.py)

Output: tests/golden/n4_goldens.pt — state_dict, seeded inputs, outputs (train and eval mode) and every gradient;
tests/test_oracle_n4_goldens.py replays them through oracle/models.py.   python tests/golden/make_n4_goldens.py
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_reference_goldens import bf, install_stub_tree, load, randomize_bn_  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'n4_goldens.pt')


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def running(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items() if 'running_' in k}


def seeded_state(module, seed):
    """A full state dict that both the generator and the replaying test can rebuild from `seed` alone: conv / linear
    weights ~ N(0, 1/sqrt(fan_in)) rounded to bf16, BatchNorm gamma ~ U(.5, 1.5), beta / running_mean ~ N(0, .1),
    running_var ~ U(.5, 1.5), in state-dict key order."""
    gen = torch.Generator().manual_seed(seed)
    state = {}
    for k, v in module.state_dict().items():
        if k.endswith('num_batches_tracked'):
            state[k] = v.clone()
        elif k.endswith('running_var'):
            state[k] = torch.rand(v.shape, generator=gen) + 0.5
        elif k.endswith('running_mean'):
            state[k] = torch.randn(v.shape, generator=gen) * 0.1
        elif v.dim() > 1:
            state[k] = bf(torch.randn(v.shape, generator=gen) / (v[0].numel() ** 0.5))
        elif k.endswith('weight'):
            state[k] = torch.rand(v.shape, generator=gen) + 0.5
        else:
            state[k] = torch.randn(v.shape, generator=gen) * 0.1
    return state


def main():
    install_stub_tree()
    pkg = types.ModuleType('torchok.models.modules.blocks')
    pkg.__path__ = []
    sys.modules['torchok.models.modules.blocks'] = pkg
    load('torchok.models.base')
    load('torchok.models.modules.bricks.convbnact')
    load('torchok.models.modules.blocks.scse')
    ocr = load('torchok.models.heads.segmentation.ocr')
    unet = load('torchok.models.necks.segmentation.unet')
    g = torch.Generator().manual_seed(2024)
    out = {}

    def prepare(m):
        randomize_bn_(m, g)
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() > 1:
                    p.copy_(bf(p * 2))
        return m

    cases = []
    for cin, ncls, mid, key, size, fsize, train in [(24, 5, 32, 16, 32, 8, False), (24, 5, 32, 16, 32, 8, True),
                                                    (16, 1, 32, 8, 24, 6, False)]:
        torch.manual_seed(cin + ncls)
        m = prepare(ocr.OCRSegmentationHead(cin, ncls, ocr_mid_channels=mid, ocr_key_channels=key))
        for mod in m.modules():                       # Dropout2d is random in train mode: parity needs p = 0
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        state = sd(m)
        m.train(train)
        image = torch.zeros(2, 3, size, size)
        f = bf(torch.randn(2, cin, fsize, fsize, generator=g)).requires_grad_(True)
        y = m([image, f])
        ys = y if isinstance(y, tuple) else (y,)
        rs = [bf(torch.randn(t.shape, generator=g)) for t in ys]
        sum((t * r).sum() for t, r in zip(ys, rs)).backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        cases.append(dict(args=(cin, ncls, mid, key), train=train, state=state, image=image, f=f.detach(), rs=rs,
                          ys=[t.detach() for t in ys], df=f.grad.clone(), grads=grads, state_after=running(m)))
    out['OCRSegmentationHead'] = cases

    cases = []
    for chans, dec, center, use_bn, train in [((8, 12, 16, 24), (16, 12, 8, 8), True, True, False),
                                              ((8, 12, 16, 24), (16, 12, 8, 8), True, True, True),
                                              ((8, 16, 24), (16, 8, 8), False, False, False)]:
        torch.manual_seed(len(chans))
        m = prepare(unet.UnetNeck(list(chans), decoder_channels=dec, use_batchnorm=use_bn, center=center))
        state = sd(m)
        m.train(train)
        size = 32
        feats = [torch.zeros(2, 3, size, size)] + \
            [bf(torch.randn(2, c, size >> (i + 1), size >> (i + 1), generator=g)).requires_grad_(True)
             for i, c in enumerate(chans)]
        image, y = m(feats)
        r = bf(torch.randn(y.shape, generator=g))
        (y * r).sum().backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        cases.append(dict(args=(chans, dec, center, use_bn), train=train, state=state,
                          feats=[f.detach() for f in feats], r=r, y=y.detach(), dfeats=[f.grad.clone() for f in feats[1:]],
                          grads=grads, state_after=running(m)))
    out['UnetNeck'] = cases
    # ---- HRNetClassificationNeck (necks/classification/hrnet.py:12-85).  The file imports timm's Bottleneck; timm is not
    # installed, so torchvision's independent Bottleneck (same constructor order, same v1.5 arithmetic, same state-dict
    # keys) stands in for it — everything else (layer construction, the S7 overwrite quirk in forward) is the
    # reference's code.  The 3.9 M parameters are not stored: both sides build them with `seeded_state`.
    import torchvision
    for name in ('timm', 'timm.models'):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    tres = types.ModuleType('timm.models.resnet')
    tres.Bottleneck = torchvision.models.resnet.Bottleneck
    sys.modules['timm.models.resnet'] = tres
    for name in ('torchok.models.necks.classification',):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    hn = load('torchok.models.necks.classification.hrnet')
    cases = []
    for chans, sizes, train in [((18, 36, 72, 144), (8, 4, 2, 1), False), ((18, 36, 72, 144), (8, 4, 2, 1), True)]:
        m = hn.HRNetClassificationNeck(list(chans))
        m.load_state_dict(seeded_state(m, 77))
        m.train(train)
        feats = [bf(torch.randn(3, c, s, s, generator=g)).requires_grad_(True) for c, s in zip(chans, sizes)]
        y = m(feats)
        r = bf(torch.randn(y.shape, generator=g))
        (y * r).sum().backward()
        cases.append(dict(chans=chans, train=train, seed=77, feats=[f.detach() for f in feats], r=r, y=y.detach(),
                          dfeats=[None if f.grad is None else f.grad.clone() for f in feats],
                          grad_norms={n: float(p.grad.norm()) for n, p in m.named_parameters() if p.grad is not None},
                          state_after=running(m) if train else {}))
    out['HRNetClassificationNeck'] = cases

    torch.save(out, OUT)
    print(f'wrote {OUT}: ' + ', '.join(f'{k} x{len(v)}' for k, v in out.items()))


if __name__ == '__main__':
    main()
