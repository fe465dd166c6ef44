"""Weight loading helpers.  `pretrained: true` cannot download in an offline box (the reference pulls timm URLs,
torchok/models/backbones/resnet.py:566-568): weights are taken from $TORCHOK_B200_PRETRAINED/<variant>.pth when that
file exists, otherwise the model keeps its random init and a warning says so."""
import os
import warnings

import torch


def load_pretrained(model, variant):
    root = os.environ.get('TORCHOK_B200_PRETRAINED', '')
    path = os.path.join(root, f'{variant}.pth') if root else ''
    if path and os.path.exists(path):
        state = torch.load(path, map_location='cpu')
        state = state.get('state_dict', state)
        own = model.state_dict()
        state = {k: v for k, v in state.items() if k in own and own[k].shape == v.shape}
        model.load_state_dict(state, strict=False)
        return True
    warnings.warn(f'pretrained weights for {variant} are not available offline '
                  f'(set TORCHOK_B200_PRETRAINED to a directory holding {variant}.pth); using random init')
    return False
