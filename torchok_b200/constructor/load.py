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


# ----------------------------------------------------------------------------------------------------------------
# task.load_checkpoint (torchok/constructor/load.py:9-227; called from BaseTask.on_fit_start / on_test_start /
# on_predict_start, torchok/tasks/base.py:113-123).  Pinned by the known answers of the reference's
# tests/base_tests/constructor/test_load_checkpoint.py:43-140 (mirrored in tests/test_front_door.py).
# ----------------------------------------------------------------------------------------------------------------
def load_state_dict(checkpoint_path, map_location='cpu'):
    """State dict stored at `checkpoint_path`: either the file itself or its 'state_dict' entry (Lightning layout and
    the layout runner.Runner.save_checkpoint writes)."""
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    return checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint


def _with_prefix(prefix, state_dict):
    prefix = prefix.strip(' .') + '.'
    return {(k if k.startswith(prefix) else prefix + k): v for k, v in state_dict.items()}


def generate_required_state_dict(base_state_dict, overridden_name2state_dict, exclude_keys, model_keys,
                                 initial_state_dict):
    """The state dict `load_checkpoint` hands to `load_state_dict`:

    1. start from the base checkpoint;
    2. lay the per-module override checkpoints over it, shallow module names first, so that the deepest override
       (most dots in its module name) wins; override keys are given relative to their module or absolute;
    3. every model key that starts with one of `exclude_keys` is taken from the model's own initial state instead.
       An exclude key matching nothing in the model is an error (ValueError), as in the reference (load.py:187-192).
    """
    by_depth = sorted(overridden_name2state_dict.items(), key=lambda kv: kv[0].count('.'))   # stable: ties keep order
    required = dict(base_state_dict)
    for name, state in by_depth:
        required.update(_with_prefix(name, state))
    for exclude in exclude_keys:
        hits = [k for k in model_keys if k.startswith(exclude)]
        if not hits:
            raise ValueError(f'Load checkpoint. Found exclude key {exclude} which not in model_keys.')
        for k in hits:
            required[k] = initial_state_dict[k]
    return required


def load_checkpoint(model, base_ckpt_path=None, overridden_name2ckpt_path=None, exclude_keys=None, strict=True):
    """`task.load_checkpoint` block of a config (config_structure.py:107-112).  No paths at all is a no-op."""
    if base_ckpt_path is None and overridden_name2ckpt_path is None:
        return
    initial = model.state_dict()
    base = load_state_dict(base_ckpt_path) if base_ckpt_path is not None else {}
    overrides = {name: load_state_dict(path) for name, path in (overridden_name2ckpt_path or {}).items()}
    required = generate_required_state_dict(base, overrides, list(exclude_keys or ()), list(initial.keys()), initial)
    model.load_state_dict(required, strict=strict)
