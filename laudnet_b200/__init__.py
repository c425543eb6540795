"""laudnet_b200 - B200-native (sm_100a) implementation of LAUDNet's dynamic-operator hot path.

Public surface mirrors the reference's `imagenet_classification/models` package:
`uni_resnet50`, `uni_resnet101` and the operators in `laudnet_b200.utils`.
Importing the package does not need a GPU; running any operator does, and
needs the in-tree CUDA library (`python -m laudnet_b200.build`).
"""
from ._lib import LaudError, LIB_PATH  # noqa: F401
from .laud_resnet import Bottleneck, ResNet, uni_resnet50, uni_resnet101  # noqa: F401
from .utils import (ExpandMask, Masker_channel_conv_linear, Masker_channel_MLP, Masker_spatial,  # noqa: F401
                    apply_channel_mask, apply_spatial_mask)

__version__ = "0.1.0"
