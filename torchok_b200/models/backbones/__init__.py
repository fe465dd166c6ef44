from . import resnet  # noqa: F401
from . import hrnet  # noqa: F401
from . import swin  # noqa: F401
