from ..constructor import LOSSES
from .base import JointLoss  # noqa: F401
from .classification import CrossEntropyLoss  # noqa: F401
from . import pairwise  # noqa: F401
from .segmentation import DiceLoss  # noqa: F401
