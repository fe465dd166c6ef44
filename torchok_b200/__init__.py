"""torchok_b200 — the forward/backward hot path of eora-ai/torchok on hand-written sm_100a CUDA.

Importing the package registers the backbones / necks / poolings / heads / losses / tasks under the reference's
registry names (torchok/constructor/__init__.py:4-17).  The kernels live in libtokb200.so (include/tokb200.h) and are
loaded on first use; nothing in the product path imports oracle/.
"""
from . import callbacks, constructor, data, losses, metrics, models, optim, tasks  # noqa: F401
from .constructor import (BACKBONES, CALLBACKS, DATASETS, DETECTION_NECKS, HEADS, LOSSES, METRICS, NECKS, OPTIMIZERS,
                          POOLINGS, SAMPLERS, SCHEDULERS, TASKS, TRANSFORMS)  # noqa: F401
from .constructor.config import Config, load_config  # noqa: F401

__version__ = '0.1.0'
