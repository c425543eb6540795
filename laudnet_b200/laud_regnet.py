"""Drop-in LAUD-RegNet-Y backbone (`lad_regnet_y_400mf` ... `lad_regnet_y_16gf`).

Mirrors the public interface of the reference's `imagenet_classification/models/laud_regnet.py`:

  * factories `lad_regnet_y_*(**kw)` and `LAD_RegNet(block_params, ...)` with the same keyword arguments (:468-488);
  * the same module tree, hence the same `state_dict` keys (`stem.0/1`, `trunk_output.block{s}.block{s}-{i}.
    {proj.0/1, f.a.0/1, f.b.0/1, f.se.fc1/fc2, f.c.0/1, f.masker_spatial.conv, f.masker_channel.*}`, `fc`);
  * `forward(x, temperature)` returning the 7-tuple `(logits, rho_conv3[4], rho_conv2[4], rho_conv1[4],
    rho_channel[4], flops_perc[n_blocks], flops)` (:574-613) and `get_optim_policies()` (:615-657).

Execution (eval mode, CUDA): the 1x1 convolutions a / c / proj run on the tcgen05 conv kernel shared with LAUD-ResNet
(c fuses BN + spatial gate + residual + ReLU), conv b on the grouped-conv kernel, SE as GAP + gate + channel scale;
see `_engine_regnet.py`.  Parity mode only: SE pools the DENSE conv-b output exactly as the
reference does, so a/b are evaluated everywhere and the spatial gate applies after conv c.
RegNet-X variants are rejected: the reference itself cannot run them (its transform calls `self.se` unconditionally,
laud_regnet.py:194).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ._lib import LaudError
from .utils import ExpandMask, Masker_channel_conv_linear, Masker_channel_MLP, Masker_spatial

__all__ = ["LAD_RegNet", "BlockParams", "stage_params", "lad_regnet_y_400mf", "lad_regnet_y_800mf",
           "lad_regnet_y_1_6gf", "lad_regnet_y_3_2gf", "lad_regnet_y_8gf", "lad_regnet_y_16gf"]


def _divisible(v: float, d: int) -> int:
    """Nearest multiple of d, never below d and never more than 10 % below v (torchvision `_make_divisible`)."""
    r = max(d, int(v + d / 2) // d * d)
    return r + d if r < 0.9 * v else r


def stage_params(depth: int, w_0: int, w_a: float, w_m: float, group_width: int,
                 bottleneck_multiplier: float = 1.0) -> Tuple[List[int], List[int], List[int]]:
    """(stage widths, stage depths, group widths) of a RegNet from its design-space parameters: block j has the
    quantised width round8(w_0 * w_m ** round(log_{w_m}((w_0 + w_a j) / w_0))); runs of equal widths form the stages;
    widths are then made divisible by the group width.  Restates BlockParams.from_init_params (laud_regnet.py:374-465,
    the algorithm of "Designing Network Design Spaces")."""
    if w_a < 0 or w_0 <= 0 or w_m <= 1 or w_0 % 8:
        raise ValueError("Invalid RegNet settings")
    cont = torch.arange(depth) * w_a + w_0
    expo = torch.round(torch.log(cont / w_0) / math.log(w_m))
    per_block = (torch.round(w_0 * torch.pow(w_m, expo) / 8) * 8).int().tolist()
    widths: List[int] = []
    depths: List[int] = []
    for w in per_block:
        if widths and widths[-1] == w:
            depths[-1] += 1
        else:
            widths.append(w)
            depths.append(1)
    gws = []
    for i, w in enumerate(widths):
        wb = int(w * bottleneck_multiplier)
        g = min(group_width, wb)
        widths[i] = int(_divisible(wb, g) / bottleneck_multiplier)
        gws.append(g)
    return widths, depths, gws


class BlockParams:
    """Per-stage settings (same fields as the reference's BlockParams, laud_regnet.py:357-372)."""

    def __init__(self, depths, widths, group_widths, bottleneck_multipliers, strides, se_ratio=None):
        self.depths, self.widths, self.group_widths = depths, widths, group_widths
        self.bottleneck_multipliers, self.strides, self.se_ratio = bottleneck_multipliers, strides, se_ratio

    @classmethod
    def from_init_params(cls, depth, w_0, w_a, w_m, group_width, bottleneck_multiplier=1.0, se_ratio=None, **kwargs):
        widths, depths, gws = stage_params(depth, w_0, w_a, w_m, group_width, bottleneck_multiplier)
        n = len(widths)
        return cls(depths, widths, gws, [bottleneck_multiplier] * n, [2] * n, se_ratio)

    def _get_expanded_params(self):
        return zip(self.widths, self.strides, self.depths, self.group_widths, self.bottleneck_multipliers)


def _conv_bn(cin, cout, k, stride, groups=1, act=True):
    """Parameter container with the key layout of torchvision's ConvNormActivation: `0` conv, `1` norm."""
    layers = [nn.Conv2d(cin, cout, k, stride, padding=(k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if act:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class _SqueezeExcitation(nn.Module):
    """Parameter container with torchvision SqueezeExcitation's keys (`fc1`, `fc2`: 1x1 convolutions with bias)."""

    def __init__(self, channels, squeeze):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, squeeze, 1)
        self.fc2 = nn.Conv2d(squeeze, channels, 1)


class BottleneckTransform(nn.Module):
    def __init__(self, width_in, width_out, stride, group_width, bottleneck_multiplier, se_ratio,
                 spatial_mask_channel_group=1, channel_dyn_granularity=1, output_size=56, mask_spatial_granularity=1,
                 dyn_mode="both", channel_masker="conv_linear", channel_masker_layers=2, reduction=16):
        super().__init__()
        assert dyn_mode in ["channel", "spatial", "both"]
        assert channel_masker in ["conv_linear", "MLP"]
        if not se_ratio:
            raise LaudError("RegNet-X (no Squeeze-Excitation) is not runnable in the reference either "
                            "(laud_regnet.py:194 calls self.se unconditionally); use a RegNet-Y variant")
        self.dyn_mode = dyn_mode
        w_b = int(round(width_out * bottleneck_multiplier))
        g = w_b // group_width
        self.group_width = group_width
        self.stride = stride
        self.a = _conv_bn(width_in, w_b, 1, 1)
        self.b = _conv_bn(w_b, w_b, 3, stride, groups=g)
        width_se_out = int(round(se_ratio * width_in))
        self.se = _SqueezeExcitation(w_b, width_se_out)
        self.c = _conv_bn(w_b, width_out, 1, 1, act=False)
        assert channel_dyn_granularity <= w_b
        self.channel_dyn_granularity = channel_dyn_granularity
        channel_dyn_group = w_b // channel_dyn_granularity
        self.channel_dyn_group = channel_dyn_group
        self.spatial_mask_channel_group = spatial_mask_channel_group
        self.conv1_flops_per_pixel = width_in * w_b
        self.conv2_flops_per_pixel = w_b * w_b * 9 // g
        self.conv3_flops_per_pixel = w_b * width_out
        self.se_flops_per_pixel = w_b * width_se_out * 2
        self.output_size = output_size
        self.mask_spatial_granularity = mask_spatial_granularity
        self.mask_size = output_size // mask_spatial_granularity
        self.masker_spatial = None
        self.masker_channel = None
        if dyn_mode in ["spatial", "both"]:
            self.masker_spatial = Masker_spatial(width_in, spatial_mask_channel_group, self.mask_size)
            self.mask_expander2 = ExpandMask(stride=1, padding=0, mask_channel_group=spatial_mask_channel_group)
            self.mask_expander1 = ExpandMask(stride=stride, padding=1, mask_channel_group=spatial_mask_channel_group)
        if dyn_mode in ["channel", "both"]:
            if channel_masker == "conv_linear":
                self.masker_channel = Masker_channel_conv_linear(width_in, channel_dyn_group, reduction=reduction)
            else:
                self.masker_channel = Masker_channel_MLP(width_in, channel_dyn_group, layers=channel_masker_layers,
                                                         reduction=reduction)


class ResBottleneckBlock(nn.Module):
    def __init__(self, width_in, width_out, stride, group_width=1, bottleneck_multiplier=1.0, se_ratio=None, **dyn):
        super().__init__()
        self.proj = None
        if width_in != width_out or stride != 1:
            self.proj = _conv_bn(width_in, width_out, 1, stride, act=False)
            self.downsample_flops = width_in * width_out
        self.f = BottleneckTransform(width_in, width_out, stride, group_width, bottleneck_multiplier, se_ratio, **dyn)
        self.activation = nn.ReLU(inplace=True)
        self.dyn_mode = dyn.get("dyn_mode", "both")
        self.width_in, self.width_out, self.stride = width_in, width_out, stride


class AnyStage(nn.Sequential):
    def __init__(self, width_in, width_out, stride, depth, group_width, bottleneck_multiplier, se_ratio, stage_index,
                 **dyn):
        super().__init__()
        for i in range(depth):
            self.add_module(f"block{stage_index}-{i}",
                            ResBottleneckBlock(width_in if i == 0 else width_out, width_out, stride if i == 0 else 1,
                                               group_width, bottleneck_multiplier, se_ratio, **dyn))


class LAD_RegNet(nn.Module):
    def __init__(self, block_params: BlockParams, num_classes: int = 1000, stem_width: int = 32, stem_type=None,
                 block_type=None, norm_layer=None, activation=None, input_size=224,
                 spatial_mask_channel_group=[1, 1, 1, 1], mask_spatial_granularity=[1, 1, 1, 1],
                 channel_dyn_granularity=[1, 1, 1, 1], dyn_mode=["both", "both", "both", "both"],
                 channel_masker=["MLP", "MLP", "MLP", "MLP"], channel_masker_layers=[1, 1, 1, 1],
                 reduction_ratio=[16, 16, 16, 16], lr_mult=1.0, **kwargs):
        super().__init__()
        if stem_type is not None or block_type is not None or activation is not None:
            raise LaudError("custom stem / block / activation types are not supported by the CUDA path")
        self.dyn_mode = dyn_mode
        assert lr_mult is not None
        self.lr_mult = lr_mult
        self.input_size = input_size
        self.stem = _conv_bn(3, stem_width, 3, 2)
        current = stem_width
        blocks = []
        for i, (width_out, stride, depth, group_width, bm) in enumerate(block_params._get_expanded_params()):
            blocks.append((f"block{i + 1}", AnyStage(
                current, width_out, stride, depth, group_width, bm, block_params.se_ratio, stage_index=i + 1,
                spatial_mask_channel_group=spatial_mask_channel_group[i],
                channel_dyn_granularity=channel_dyn_granularity[i], output_size=input_size // (2 ** (i + 2)),
                mask_spatial_granularity=mask_spatial_granularity[i], dyn_mode=dyn_mode[i],
                channel_masker=channel_masker[i], channel_masker_layers=channel_masker_layers[i],
                reduction=reduction_ratio[i])))
            current = width_out
        self.trunk_output = nn.Sequential(OrderedDict(blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(current, num_classes)
        for name, m in self.named_modules():            # reference init, laud_regnet.py:562-572
            if isinstance(m, nn.Conv2d) and "masker" not in name:
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, mean=0.0, std=math.sqrt(2.0 / fan_out))
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear) and "masker" not in name:
                nn.init.normal_(m.weight, mean=0.0, std=0.01)
                nn.init.zeros_(m.bias)
        from ._engine_regnet import RegNetEngine
        self._engine = RegNetEngine(self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._engine.prepared_for = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if hasattr(self, "_engine"):
            self._invalidate()
        return out

    def blocks(self):
        return [blk for stage in self.trunk_output for blk in stage]

    def forward(self, x, temperature=1.0, keep=None, forced=None):
        if self.training:
            raise LaudError("LAD_RegNet: training mode (Gumbel gates, reference utils.py:56-58) is not part of the "
                            "CUDA inference path; call .eval()")
        logits, stats = self._engine.forward(x, keep, forced=forced)
        r3, r2, r1, rc, perc, flops = self._engine.split_stats(stats)
        return logits, r3, r2, r1, rc, perc, flops

    def forward_logits(self, x):
        if self.training:
            raise LaudError("LAD_RegNet: call .eval() first")
        return self._engine.forward(x)[0]

    def capture(self, x_example):
        if self.training:
            raise LaudError("LAD_RegNet: call .eval() first")
        from ._engine import GraphedForward
        return GraphedForward(self._engine, x_example)

    def get_optim_policies(self):
        """Same two parameter groups as the reference (laud_regnet.py:615-657)."""
        groups = {"backbone_params": [], "masker_params": []}
        for name, m in self.named_modules():
            key = "masker_params" if "masker" in name else "backbone_params"
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                groups[key].extend(list(m.parameters())[:2])
            elif isinstance(m, nn.BatchNorm2d) or (key == "masker_params" and isinstance(m, nn.BatchNorm1d)):
                groups[key].extend(list(m.parameters()))
        return [
            {"params": groups["backbone_params"], "lr_mult": self.lr_mult, "decay_mult": 1.0, "name": "backbone_params"},
            {"params": groups["masker_params"], "lr_mult": 1.0, "decay_mult": 1.0, "name": "masker_params"},
        ]


def _lad_regnet(block_params: BlockParams, pretrained: bool, **kwargs) -> LAD_RegNet:
    if pretrained:
        raise LaudError("pretrained=True needs network access to the torchvision model zoo; "
                        "load a checkpoint with load_state_dict instead")
    kwargs.pop("norm_layer", None)
    return LAD_RegNet(block_params, **kwargs)


def lad_regnet_y_400mf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=16, w_0=48, w_a=27.89, w_m=2.09, group_width=8, se_ratio=0.25), pretrained, **kw)


def lad_regnet_y_800mf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=14, w_0=56, w_a=38.84, w_m=2.4, group_width=16, se_ratio=0.25), pretrained, **kw)


def lad_regnet_y_1_6gf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=27, w_0=48, w_a=20.71, w_m=2.65, group_width=24, se_ratio=0.25), pretrained, **kw)


def lad_regnet_y_3_2gf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=21, w_0=80, w_a=42.63, w_m=2.66, group_width=24, se_ratio=0.25), pretrained, **kw)


def lad_regnet_y_8gf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=17, w_0=192, w_a=76.82, w_m=2.19, group_width=56, se_ratio=0.25), pretrained, **kw)


def lad_regnet_y_16gf(pretrained=False, progress=True, **kw):
    return _lad_regnet(BlockParams.from_init_params(depth=18, w_0=200, w_a=106.23, w_m=2.48, group_width=112, se_ratio=0.25), pretrained, **kw)
