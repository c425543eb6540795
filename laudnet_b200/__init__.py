"""laudnet_b200 - B200-native (sm_100a) implementation of LAUDNet's dynamic-operator hot path.

Public surface mirrors the reference's `imagenet_classification/models` package:
`uni_resnet50`, `uni_resnet101`, `lad_regnet_y_*` and the operators in `laudnet_b200.utils`.
Importing the package does not need a GPU; running any operator does, and
needs the in-tree CUDA library (`python -m laudnet_b200.build`).
"""
from ._lib import LaudError, LIB_PATH  # noqa: F401
from .laud_resnet import Bottleneck, ResNet, uni_resnet50, uni_resnet101  # noqa: F401
from .laud_regnet import (LAD_RegNet, lad_regnet_y_400mf, lad_regnet_y_800mf, lad_regnet_y_1_6gf,  # noqa: F401
                          lad_regnet_y_3_2gf, lad_regnet_y_8gf, lad_regnet_y_16gf)
from .mmdet_adapter import LAD_MMDet_ResNet  # noqa: F401
from .adavit import AdaViT, ada_deit_tiny_patch16_224, ada_deit_small_patch16_224, ada_deit_base_patch16_224  # noqa: F401
from .utils import (ExpandMask, Masker_channel_conv_linear, Masker_channel_MLP, Masker_spatial,  # noqa: F401
                    apply_channel_mask, apply_spatial_mask)

__version__ = "0.2.0"
