"""`paramwise_cfg` of the optimizer block: per-parameter learning-rate and weight-decay multipliers.

Restates the rules of the reference's `Constructor.add_params` (torchok/constructor/constructor.py:162-251), which walks
the module tree and builds one torch.optim param group per parameter:

  * `custom_keys`: {substring: {lr_mult, decay_mult}} — keys are tried longest first (ties alphabetical); the first key
    contained in the parameter's dotted name wins and every other rule is skipped for that parameter;
  * otherwise `bias_lr_mult` scales the lr of every parameter called `bias` that does not belong to a normalisation
    layer, and the weight decay is scaled by `norm_decay_mult` (parameters of BatchNorm / InstanceNorm / GroupNorm /
    LayerNorm), else `dwconv_decay_mult` (depth-wise Conv2d: in_channels == groups), else `bias_decay_mult` (`bias`);
  * frozen parameters keep the defaults.  (`dcn_offset_lr_mult` needs deformable-conv modules, which this package does
    not have; the key is accepted and ignored, as it is in the reference for models without DCN.)

Here the result is a {parameter: (lr_mult, decay_mult)} map: the arena optimizers (engine.ArenaSGD / ArenaAdam) keep ONE
flat step kernel and look the multipliers up in a per-parameter segment table instead of launching per group.
Decay multipliers only matter when the optimizer has a weight decay (reference: `base_wd is not None`).
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm
from torch.nn.modules.instancenorm import _InstanceNorm

_NORMS = (_BatchNorm, _InstanceNorm, torch.nn.GroupNorm, torch.nn.LayerNorm)


def paramwise_multipliers(module, paramwise_cfg=None, prefix=''):
    """{parameter: (lr_mult, decay_mult)} for every parameter below `module` (reference traversal order)."""
    out = {}
    _walk(out, module, dict(paramwise_cfg or {}), prefix)
    return out


def _walk(out, module, cfg, prefix):
    custom = cfg.get('custom_keys', {}) or {}
    keys = sorted(sorted(custom.keys()), key=len, reverse=True)
    bias_lr_mult = float(cfg.get('bias_lr_mult', 1.))
    bias_decay_mult = float(cfg.get('bias_decay_mult', 1.))
    norm_decay_mult = float(cfg.get('norm_decay_mult', 1.))
    dwconv_decay_mult = float(cfg.get('dwconv_decay_mult', 1.))
    is_norm = isinstance(module, _NORMS)
    is_dwconv = isinstance(module, torch.nn.Conv2d) and module.in_channels == module.groups
    for name, param in module.named_parameters(recurse=False):
        lr_mult, decay_mult = 1., 1.
        if param.requires_grad:
            full = f'{prefix}.{name}'
            for key in keys:
                if key in full:
                    lr_mult = float(custom[key].get('lr_mult', 1.))
                    decay_mult = float(custom[key].get('decay_mult', 1.))
                    break
            else:
                if name == 'bias' and not is_norm:
                    lr_mult = bias_lr_mult
                if is_norm:
                    decay_mult = norm_decay_mult
                elif is_dwconv:
                    decay_mult = dwconv_decay_mult
                elif name == 'bias':
                    decay_mult = bias_decay_mult
        if param not in out:   # shared parameters: the first visit decides (torch.optim would reject duplicates)
            out[param] = (lr_mult, decay_mult)
    for child_name, child in module.named_children():
        _walk(out, child, cfg, f'{prefix}.{child_name}' if prefix else child_name)
