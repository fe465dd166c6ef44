from .classification import ClassificationHead, LinearHead  # noqa: F401
from . import arcface  # noqa: F401
from . import segmentation  # noqa: F401
