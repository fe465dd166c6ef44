from . import hrnet  # noqa: F401
from . import fpn  # noqa: F401
from . import unet  # noqa: F401
