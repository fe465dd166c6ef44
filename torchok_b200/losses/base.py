"""JointLoss: weighted sum of loss modules fed through per-loss kwarg mappings.

Same behaviour as torchok/losses/base.py:7-113 (weights all-or-none, optional normalisation to sum 1, tagged values
returned alongside the total, ValueError when a mapped output is missing, KeyError for an unknown tag); pinned by the
known answers of tests/base_tests/losses/test_base_losses.py:19-77.
"""
from torch.nn import Module, ModuleList


class JointLoss(Module):
    def __init__(self, losses, mappings, tags, weights, normalize_weights=True):
        super().__init__()
        self.losses = ModuleList(losses)
        self.tags = tags
        self.mappings = mappings
        self.tag2loss = {t: m for t, m in zip(tags, self.losses) if t is not None}
        given = [w for w in weights if w is not None]
        if given and len(given) != len(losses):
            raise ValueError('Loss weights must be either specified for each loss function or '
                             'not specified for any loss function')
        self.weights = list(weights) if given else [1.] * len(self.losses)
        if normalize_weights:
            total = sum(self.weights)
            self.weights = [w / total for w in self.weights]

    def forward(self, **kwargs):
        total_loss = 0.
        tagged = {}
        for module, mapping, tag, weight in zip(self.losses, self.mappings, self.tags, self.weights):
            value = module(**self._select(mapping, kwargs))
            total_loss = total_loss + value * weight
            if tag is not None:
                tagged[tag] = value
        return total_loss, tagged

    def __getitem__(self, tag):
        if tag not in self.tag2loss:
            raise KeyError(f'Cannot access loss {tag}. You should tag your losses for direct access with a tag key')
        return self.tag2loss[tag]

    @staticmethod
    def _select(mapping, outputs):
        picked = {}
        for dst, src in mapping.items():
            if src not in outputs:
                raise ValueError(f'Cannot find {src} for your mapping {dst} : {src}. You should either add {src} '
                                 f'output to your model or remove the mapping from configuration')
            picked[dst] = outputs[src]
        return picked
