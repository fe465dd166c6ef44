from .classification import Pooling, PoolingLinear  # noqa: F401
