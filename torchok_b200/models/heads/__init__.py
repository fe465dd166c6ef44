from .classification import ClassificationHead, LinearHead  # noqa: F401
