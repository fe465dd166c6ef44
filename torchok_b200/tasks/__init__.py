from .base import BaseTask  # noqa: F401
from .classification import ClassificationTask  # noqa: F401
from . import pairwise_task  # noqa: F401
from . import segmentation  # noqa: F401
