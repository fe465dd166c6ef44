"""Name -> callable plug-in registries: the drop-in boundary of the hot path.

Behavioural mirror of torchok/constructor/registry.py:10-138 (register_class keyed by __name__, KeyError text of
`get`, `list_models` wildcard filtering with natural sort) without the timm dependency the reference pulls in at
registry.py:7 for its sort key.
"""
import fnmatch
import re
import sys
from collections import defaultdict


def _natural(text):
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r'(\d+)', text.lower())]


class Registry:
    def __init__(self, name):
        self.name = name
        self.entrypoints = {}
        self.object_to_module = {}
        self.module_to_objects = defaultdict(set)

    # -- lookup ---------------------------------------------------------------------------------------------------
    def get(self, key):
        try:
            return self.entrypoints[key]
        except KeyError:
            raise KeyError(f'{key} is not in the {self.name} registry') from None

    __getitem__ = get

    def __contains__(self, key):
        return key in self.entrypoints

    def __repr__(self):
        return f'{type(self).__name__}(name={self.name}, items={list(self.entrypoints)})'

    # -- registration ---------------------------------------------------------------------------------------------
    def register_class(self, fn):
        if not callable(fn):
            raise TypeError(f'{fn} must be callable')
        key = fn.__name__
        if key in self.entrypoints:
            raise KeyError(f'{key} is already registered in {self.name}')
        owner = sys.modules.get(fn.__module__)
        if owner is not None:
            exported = getattr(owner, '__all__', None)
            if exported is None:
                owner.__all__ = [key]
            else:
                exported.append(key)
        leaf = fn.__module__.rsplit('.', 1)[-1]
        self.entrypoints[key] = fn
        self.object_to_module[key] = leaf
        self.module_to_objects[leaf].add(key)
        return fn

    # -- listing --------------------------------------------------------------------------------------------------
    def list_models(self, filter='', module='', exclude_filters=''):
        pool = list(self.module_to_objects[module]) if module else list(self.entrypoints)
        if filter:
            patterns = filter if isinstance(filter, (tuple, list)) else [filter]
            chosen = set()
            for pat in patterns:
                chosen.update(fnmatch.filter(pool, pat))
        else:
            chosen = set(pool)
        if exclude_filters:
            patterns = exclude_filters if isinstance(exclude_filters, (tuple, list)) else [exclude_filters]
            for pat in patterns:
                chosen.difference_update(fnmatch.filter(chosen, pat))
        return sorted(chosen, key=_natural)
