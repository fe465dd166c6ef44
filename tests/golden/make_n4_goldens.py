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
    torch.save(out, OUT)
    print(f'wrote {OUT}: ' + ', '.join(f'{k} x{len(v)}' for k, v in out.items()))


if __name__ == '__main__':
    main()
