"""Detection-backbone adapter: the interface of the reference's `LAD_MMDet_ResNet`
(mmdetection-2.21.0/mmdet/models/backbones/lad_mmdet_resnet.py:331-762; the 3.3.0 copy differs by `out_indices`
selection, mmdetection-3.3.0/.../lad_mmdet_resnet.py:750-754) on top of the CUDA engine.

Same constructor keywords (the LAUD ones and the ResNet ones that matter for inference), same parameter names - the
reference builds its norm layers with `build_norm_layer(postfix=n)`, i.e. `bn1/bn2/bn3`, so a detection checkpoint's
`backbone.*` keys are `conv1, bn1, layer{1-4}.{i}.{conv1..3, bn1..3, downsample.0/1, masker_channel.*, masker_spatial.*}`:
exactly this module's `state_dict()` - and the same return value of `forward(x, iter_now=0, len_loader=100)`:

    (tuple of the stage feature maps selected by out_indices  [fp32 NCHW],
     additional = {spatial_sparsity_conv3/2/1: [4 x Tensor], channel_sparsity: [4 x Tensor], flops_perc_list, flops, dense_flops},
     model_configs = {dyn_mode, sparsity_target})

mmcv / mmengine are not installed in this image, so the class is NOT registered in a `BACKBONES` registry here;
INTEGRATION.md shows the two lines a maintainer adds (`@BACKBONES.register_module()` on a subclass).

What the reference supports in this backbone and so does this adapter: `dyn_mode` 'channel' and 'layer' per stage
(maskers are only built for those, :161-176), pytorch style, no DCN / plugins / deep stem / avg_down.  The gates run
their eval branch; with `gumbel_noise=` they run the training branch on supplied samples with frozen BN (`norm_eval=True`,
the reference's default) - see `ResNet.forward`.
Limitation of this round: the engine's kernels take H x W maps but the host sequencing assumes SQUARE inputs whose side
is a multiple of 32; a rectangular detection input (800 x 1333) raises `LaudError`.  Changing the input side re-prepares.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from ._lib import LaudError
from .laud_resnet import Bottleneck, ResNet

__all__ = ["LAD_MMDet_ResNet"]


class LAD_MMDet_ResNet(nn.Module):
    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    def __init__(self, depth, in_channels=3, stem_channels=None, base_channels=64, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style="pytorch", deep_stem=False, avg_down=False,
                 frozen_stages=-1, conv_cfg=None, norm_cfg=None, norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False, zero_init_residual=True,
                 pretrained=None, init_cfg=None, sparsity_target=None, temperature_0=None, temperature_t=None,
                 spatial_mask_channel_group=[1, 1, 1, 1], mask_spatial_granularity=[1, 1, 1, 1],
                 channel_dyn_granularity=[1, 1, 1, 1], dyn_mode=["both", "both", "both", "both"],
                 channel_masker=["MLP", "MLP", "MLP", "MLP"], channel_masker_layers=[1, 1, 1, 1],
                 reduction_ratio=[16, 16, 16, 16], input_size=224):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f"invalid depth {depth} for resnet")
        unsupported = dict(deep_stem=deep_stem, avg_down=avg_down, dcn=dcn, plugins=plugins, with_cp=with_cp)
        for k, v in unsupported.items():
            if v:
                raise LaudError(f"LAD_MMDet_ResNet: {k} is not supported by the CUDA path")
        if in_channels != 3 or num_stages != 4 or tuple(strides) != (1, 2, 2, 2) or tuple(dilations) != (1, 1, 1, 1) \
                or style != "pytorch" or base_channels != 64 or stem_channels not in (None, 64):
            raise LaudError("LAD_MMDet_ResNet: only the standard 4-stage pytorch-style ResNet geometry is supported")
        for m in dyn_mode:
            if m not in ("channel", "layer"):        # the reference builds maskers only for these (:161-176)
                raise LaudError(f"LAD_MMDet_ResNet: dyn_mode '{m}' - the reference's detection backbone implements "
                                "'channel' and 'layer' only")
        self.depth, self.out_indices = depth, tuple(out_indices)
        self.dyn_mode, self.sparsity_target = dyn_mode, sparsity_target
        self.temperature_0, self.temperature_t = temperature_0, temperature_t
        self.norm_eval, self.frozen_stages = norm_eval, frozen_stages
        net = ResNet(Bottleneck, list(self.arch_settings[depth]), num_classes=1, zero_init_residual=zero_init_residual,
                     input_size=input_size, spatial_mask_channel_group=spatial_mask_channel_group,
                     mask_spatial_granularity=mask_spatial_granularity, channel_dyn_granularity=channel_dyn_granularity,
                     dyn_mode=dyn_mode, channel_masker=channel_masker, channel_masker_layers=channel_masker_layers,
                     reduction_ratio=reduction_ratio)
        # the reference's parameter names: register the trunk's modules here (no `fc`, no prefix); the ResNet wrapper that
        # owns the engine stays outside the module registry
        self.conv1, self.bn1 = net.conv1, net.bn1
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        object.__setattr__(self, "_net", net)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._net._invalidate())

    @property
    def norm1(self):                      # mmdet's name for the stem norm (`self.norm1_name`)
        return self.bn1

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        net = self.__dict__.get("_net")
        if net is not None:
            net.fc._apply(fn)             # the (unused) 1-class head follows the trunk's device
            net._invalidate()
        return out

    def train(self, mode=True):
        """Normalisation layers stay in eval mode (`norm_eval`, the reference's default, :753-762); the CUDA path has no
        batch-statistics BatchNorm at all."""
        super().train(mode)
        if mode and not self.norm_eval:
            raise LaudError("LAD_MMDet_ResNet: norm_eval=False (batch-statistics BatchNorm) is not part of the CUDA path")
        for m in self.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
        return self

    def _set_input_size(self, h: int, w: int) -> None:
        """Feature-map geometry of an (h, w) input: stage s runs at (h, w) / (4 << s); the gates follow the actual feature
        size (lad_mmdet_resnet.py:274).  Non-square inputs (detection batches are padded to a multiple of 32 per side)."""
        net = self._net
        if net.input_size == h and (getattr(net, "input_w", None) or net.input_size) == w:
            return
        net.input_size, net.input_w = h, w
        for s, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            for blk in layer:
                blk.output_size, blk.output_w = h // (4 << s), w // (4 << s)
                blk.mask_size = 1 if blk.dyn_mode == "layer" else max(1, blk.output_size // blk.mask_spatial_granularity)
                blk._solo_engine = None
        net._invalidate()

    def forward(self, x, iter_now=0, len_loader=100, gumbel_noise=None, keep=None, forced=None):
        if x.dim() != 4 or x.shape[2] % 32 or x.shape[3] % 32:
            raise LaudError(f"LAD_MMDet_ResNet: expected [B, 3, H, W] with H and W multiples of 32, got {tuple(x.shape)}")
        gates_train = any(m.training for m in self.modules() if "Masker" in type(m).__name__)
        if gates_train and gumbel_noise is None:
            raise LaudError("LAD_MMDet_ResNet: gates in training mode draw Gumbel noise (utils.py:56-58); pass gumbel_noise= or .eval()")
        self._set_input_size(int(x.shape[2]), int(x.shape[3]))
        net = self._net
        eng = net._engine
        temperature = self.temperature_0 if self.temperature_0 is not None else 1.0
        outs = []
        _, stats = eng.forward(x, keep, forced=forced, gumbel_noise=gumbel_noise, temperature=temperature, stage_outputs=outs)
        r3, r2, r1, rc, perc, flops = eng.split_stats(stats)
        # dense FLOPs of stem + trunk (:688-693 and the per-block dense_flops of :243-300): static for a given input size
        hh, ww = int(x.shape[2]), int(x.shape[3])
        c0 = self.conv1.weight.shape[0]
        dense = 3 * c0 * (hh // 2) * (ww // 2) * 49 + c0 * (hh // 4) * (ww // 4) * 9
        consts = eng.stats_consts
        dense += int(consts[:, 0:6].sum())
        additional = {"spatial_sparsity_conv3": r3, "spatial_sparsity_conv2": r2, "spatial_sparsity_conv1": r1,
                      "channel_sparsity": rc, "flops_perc_list": perc, "flops": flops,
                      "dense_flops": torch.tensor(float(dense), device=x.device)}
        model_configs = {"dyn_mode": self.dyn_mode, "sparsity_target": self.sparsity_target}
        return tuple(outs[i] for i in self.out_indices), additional, model_configs
