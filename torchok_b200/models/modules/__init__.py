from .layers import BatchNorm2d, Conv2d, MaxPool2d, ReLU, conv_bn_act  # noqa: F401
