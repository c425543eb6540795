"""Drop-in LAUD-ResNet backbone (`uni_resnet50` / `uni_resnet101`).

Mirrors the public interface of the reference's
`imagenet_classification/models/laud_resnet.py`:

  * constructors `uni_resnet50(**kw)`, `uni_resnet101(**kw)` and
    `ResNet(block, layers, ...)` with the same keyword arguments (:169-181);
  * the same module tree, hence the same `state_dict` keys (`conv1`, `bn1`,
    `layer{1-4}.{i}.{conv1..3,bn1..3,downsample.0/1,masker_channel.*,
    masker_spatial.*}`, `fc`) - checkpoints load unchanged;
  * `forward(x, temperature)` returning the 7-tuple
    `(logits, rho_conv3[4], rho_conv2[4], rho_conv1[4], rho_channel[4],
    flops_perc[n_blocks], flops)` (:363);
  * `Bottleneck.forward((x, l3, l2, l1, lc, lperc, flops), temperature)`
    threading the same 7-tuple (:88-165), and `get_optim_policies()` (:365-401).

What differs is how it executes: in eval mode on a CUDA device the forward is a
sequence of hand-written sm_100a kernels (see `_engine.py`, `csrc/`).  The
modules below hold parameters only; their torch `forward`s are never called.
Training mode and CPU tensors raise `LaudError` - there is no fallback path.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from ._engine import BlockOutputs, BlockPlan, ResNetEngine, fold_bn, pack_conv_weight
from ._lib import LaudError
from .utils import (ExpandMask, Masker_channel_conv_linear, Masker_channel_MLP, Masker_spatial, conv1x1, conv3x3,
                    to_nchw_f32, to_nhwc_f16)

__all__ = ["uni_resnet50", "uni_resnet101", "ResNet", "Bottleneck"]


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, group_width=1, dilation=1, norm_layer=None,
                 spatial_mask_channel_group=1, channel_dyn_granularity=1, output_size=56,
                 mask_spatial_granularity=1, dyn_mode="both", channel_masker="conv_linear",
                 channel_masker_layers=2, reduction=16):
        super().__init__()
        assert dyn_mode in ["channel", "spatial", "both", "layer"]
        assert channel_masker in ["conv_linear", "MLP"]
        if dilation != 1:
            raise LaudError("dilated bottlenecks are not supported by the CUDA path")
        self.dyn_mode = dyn_mode
        norm_layer = norm_layer or nn.BatchNorm2d
        width = int(planes * (64 / 64.0)) * group_width
        assert channel_dyn_granularity <= width
        self.channel_dyn_granularity = channel_dyn_granularity
        self.channel_dyn_group = width // channel_dyn_granularity
        self.spatial_mask_channel_group = spatial_mask_channel_group
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = norm_layer(width)
        self.conv2 = conv3x3(width, width, stride, group_width, dilation)
        self.bn2 = norm_layer(width)
        self.conv3 = conv1x1(width, planes * self.expansion)
        self.bn3 = norm_layer(planes * self.expansion)
        self.downsample = downsample
        self.stride = stride
        self.conv1_flops_per_pixel = inplanes * width
        self.conv2_flops_per_pixel = width * width * 9 // self.conv2.groups
        self.conv3_flops_per_pixel = width * planes * self.expansion
        if downsample is not None:
            self.downsample_flops = inplanes * planes * self.expansion
        self.output_size = output_size
        self.mask_spatial_granularity = mask_spatial_granularity
        self.mask_size = output_size // mask_spatial_granularity if dyn_mode != "layer" else 1
        self.masker_spatial = None
        self.masker_channel = None
        if dyn_mode in ["spatial", "layer", "both"]:
            self.masker_spatial = Masker_spatial(inplanes, spatial_mask_channel_group, self.mask_size)
            self.mask_expander2 = ExpandMask(stride=1, padding=0, mask_channel_group=spatial_mask_channel_group)
            self.mask_expander1 = ExpandMask(stride=stride, padding=1, mask_channel_group=spatial_mask_channel_group)
        if dyn_mode in ["channel", "both"]:
            if channel_masker == "conv_linear":
                self.masker_channel = Masker_channel_conv_linear(inplanes, self.channel_dyn_group, reduction=reduction)
            else:
                self.masker_channel = Masker_channel_MLP(inplanes, self.channel_dyn_group,
                                                         layers=channel_masker_layers, reduction=reduction)
        self._solo_engine = None

    # Stand-alone block call with the reference's calling convention: NCHW in,
    # NCHW out, the running lists extended by this block's scalars.
    def forward(self, x, temperature=1.0, forced_channel_mask=None, forced_spatial_mask=None, keep=None):
        x, l3, l2, l1, lc, lperc, flops = x
        if self.training:
            raise LaudError("Bottleneck: training mode is not part of the CUDA inference path; call .eval()")
        _lib.require_cuda(x, "Bottleneck")
        with torch.cuda.device(x.device):
            return self._forward_cuda(x, l3, l2, l1, lc, lperc, flops, forced_channel_mask, forced_spatial_mask, keep)

    def _forward_cuda(self, x, l3, l2, l1, lc, lperc, flops, forced_channel_mask, forced_spatial_mask, keep):
        plan = self._plan()
        B, C, H, W = x.shape
        if C != plan.inplanes or H != plan.H_in or W != plan.H_in:
            raise LaudError(f"Bottleneck: expected [B,{plan.inplanes},{plan.H_in},{plan.H_in}], got {tuple(x.shape)}")
        eng = self._solo_engine
        ws = eng._workspace_for_block(plan, B, x.device)
        ws["counts"].zero_()
        xin = to_nhwc_f16(x)
        out = torch.empty((B, plan.H_out, plan.H_out, plan.outplanes), dtype=torch.float16, device=x.device)
        idb = torch.empty_like(out)
        res = eng.run_block(plan, xin.view(-1), out.view(-1), idb.view(-1), B, ws, keep, forced_channel_mask,
                            forced_spatial_mask)
        out = res[:out.numel()].view(out.shape)      # an in-place layer skip returns the (converted) input buffer
        stats = torch.empty(6, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().laud_forward_stats(_lib.ptr(ws["counts"]), _lib.ptr(ws["consts"]), 1, 0, 0, 0,
                                                 _lib.ptr(stats), _lib.stream_ptr()), "laud_forward_stats")
        r3, r2, r1, rc, perc, sparse = (stats[i] for i in range(6))
        cat = lambda lst, v: v.unsqueeze(0) if lst is None else torch.cat((lst, v.unsqueeze(0)), dim=0)
        flops = flops + sparse
        return to_nchw_f32(out), cat(l3, r3), cat(l2, r2), cat(l1, r1), cat(lc, rc), cat(lperc, perc), flops

    def _plan(self) -> BlockPlan:
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise LaudError("Bottleneck: parameters must live on a CUDA device")
        if self._solo_engine is None or self._solo_engine.prepared_for != dev:
            self._solo_engine = _SoloEngine(self)
        return self._solo_engine.plans[0]


class _SoloEngine(ResNetEngine):
    """Engine over a single block, for the stand-alone Bottleneck.forward."""

    def __init__(self, blk: Bottleneck):
        self.model = None
        self.impl = _lib.CONV_AUTO
        self.channel_exec = os.environ.get("LAUD_CHANNEL_EXEC", "nskip")
        self.layer_exec = os.environ.get("LAUD_LAYER_EXEC", "skip")
        self.spatial_exec = os.environ.get("LAUD_SPATIAL_EXEC", "mask")
        self._ws = {}
        dev = blk.conv1.weight.device
        p = BlockPlan(index=0, stage=0, inplanes=blk.conv1.weight.shape[1], width=blk.conv1.weight.shape[0],
                      outplanes=blk.conv3.weight.shape[0], stride=blk.stride, H_in=blk.output_size * blk.stride,
                      H_out=blk.output_size, mode=blk.dyn_mode, gran=blk.channel_dyn_granularity,
                      G=blk.channel_dyn_group, g_spatial=blk.spatial_mask_channel_group, mask_size=blk.mask_size,
                      module=blk)
        p.w1, p.w2, p.w3 = (pack_conv_weight(c.weight) for c in (blk.conv1, blk.conv2, blk.conv3))
        p.s1, p.t1 = fold_bn(blk.bn1)
        p.s2, p.t2 = fold_bn(blk.bn2)
        p.s3, p.t3 = fold_bn(blk.bn3)
        if blk.downsample is not None:
            p.wd = pack_conv_weight(blk.downsample[0].weight)
            p.sd, p.td = fold_bn(blk.downsample[1])
        if p.use_c:
            p.pack_channel_mode(blk)
        self.plans = [p]
        self.stats_consts = self._stats_consts(dev)
        self.prepared_for = dev

    def _workspace_for_block(self, p: BlockPlan, B: int, dev) -> dict:
        key = (B, dev)
        ws = self._ws.get(key)
        if ws is None:
            i32 = dict(dtype=torch.int32, device=dev)
            f16 = dict(dtype=torch.float16, device=dev)
            hw = p.H_in * p.H_in
            ws = dict(
                a1=torch.empty(B * hw * (p.width + 16), **f16), a2=torch.empty(B * hw * (p.width + 16), **f16),
                partial=torch.empty(B * (_lib.GAP_SPLITS + 1) * max(p.inplanes, p.outplanes), dtype=torch.float32, device=dev),
                cmask=torch.empty((B, p.G), dtype=torch.uint8, device=dev), cidx=torch.empty((B, p.G), **i32),
                ccnt=torch.empty((B,), **i32),
                pb2=torch.empty(B * 16 * p.width, dtype=torch.float32, device=dev),
                pb3=torch.empty(B * p.outplanes, dtype=torch.float32, device=dev),
                inact=torch.empty(B * p.width, **f16), T=torch.empty(B * (9 * p.width + p.outplanes), **f16),
                smask=torch.empty(B * p.g_spatial * hw, dtype=torch.uint8, device=dev),
                m3=torch.empty(B * p.g_spatial * hw, dtype=torch.uint8, device=dev),
                m2=torch.empty(B * p.g_spatial * hw, dtype=torch.uint8, device=dev),
                m1=torch.empty(B * p.g_spatial * hw, dtype=torch.uint8, device=dev),
                srows=torch.empty((B,), **i32), scnt=torch.zeros((1,), **i32),
                cws=torch.zeros((max(64, B * hw // 2048 + 2),), **i32),
                rows1=torch.empty((B * hw,), **i32), rows2=torch.empty((B * hw,), **i32), rcnt=torch.zeros((2,), **i32),
                lidx=torch.empty((B * p.g_spatial,), **i32), lcnt=torch.empty((B,), **i32),
                counts=torch.zeros((1, 4), **i32), mkz=None, mkpool=None)
            consts = self.stats_consts.clone()
            S = min(p.mask_size, p.H_in)
            consts[0, 6] = B * p.G
            consts[0, 7] = B * p.g_spatial * S * S
            consts[0, 8] = B * p.g_spatial * p.H_out * p.H_out
            consts[0, 9] = B * p.g_spatial * hw
            ws["consts"] = consts.to(dev)
            self._ws[key] = ws
        return ws


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None, width_mult=1.0, input_size=224,
                 spatial_mask_channel_group=[1, 1, 1, 1], mask_spatial_granularity=[1, 1, 1, 1],
                 channel_dyn_granularity=[1, 1, 1, 1], dyn_mode=["both", "both", "both", "both"],
                 channel_masker=["MLP", "MLP", "MLP", "MLP"], channel_masker_layers=[1, 1, 1, 1],
                 reduction_ratio=[16, 16, 16, 16], lr_mult=1.0, **kwargs):
        super().__init__()
        self.dyn_mode = dyn_mode
        assert lr_mult is not None
        self.lr_mult = lr_mult
        self.input_size = input_size
        self._norm_layer = norm_layer or nn.BatchNorm2d
        if replace_stride_with_dilation not in (None, [False, False, False], (False, False, False)):
            raise LaudError("replace_stride_with_dilation is not supported by the CUDA path")
        if groups != 1 or width_per_group != 64:
            raise LaudError("grouped / wide bottlenecks are not supported by the CUDA path")
        self.inplanes = int(64 * width_mult)
        self.groups = groups
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = self._norm_layer(self.inplanes)
        stage_planes = [int(c * width_mult) for c in (64, 128, 256, 512)]
        for s in range(4):
            layer = self._make_layer(block, stage_planes[s], layers[s], stride=1 if s == 0 else 2,
                                     output_size=input_size // (4 << s),
                                     spatial_mask_channel_group=spatial_mask_channel_group[s],
                                     mask_spatial_granularity=mask_spatial_granularity[s],
                                     channel_dyn_granularity=channel_dyn_granularity[s], dyn_mode=dyn_mode[s],
                                     channel_masker=channel_masker[s],
                                     channel_masker_layers=channel_masker_layers[s],
                                     reduction_ratio=reduction_ratio[s])
            setattr(self, f"layer{s + 1}", layer)
        self.fc = nn.Linear(int(512 * width_mult * block.expansion), num_classes)
        for name, m in self.named_modules():       # reference init, laud_resnet.py:255-260
            if isinstance(m, nn.Conv2d) and "masker" not in name:
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
        self._engine = ResNetEngine(self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._engine.prepared_for = None

    def _apply(self, fn, *a, **kw):           # .cuda()/.to() move parameters: re-prepare lazily
        out = super()._apply(fn, *a, **kw)
        if hasattr(self, "_engine"):
            self._invalidate()
        return out

    def _make_layer(self, block, planes, blocks, stride, output_size, spatial_mask_channel_group,
                    mask_spatial_granularity, channel_dyn_granularity, dyn_mode, channel_masker,
                    channel_masker_layers, reduction_ratio):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       self._norm_layer(planes * block.expansion))
        common = dict(group_width=self.groups, norm_layer=self._norm_layer, output_size=output_size,
                      spatial_mask_channel_group=spatial_mask_channel_group,
                      mask_spatial_granularity=mask_spatial_granularity,
                      channel_dyn_granularity=channel_dyn_granularity, dyn_mode=dyn_mode,
                      channel_masker=channel_masker, channel_masker_layers=channel_masker_layers,
                      reduction=reduction_ratio)
        seq = [block(inplanes=self.inplanes, planes=planes, stride=stride, downsample=downsample, **common)]
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes, **common) for _ in range(1, blocks)]
        return nn.ModuleList(seq)

    def prepare(self):
        """Pack weights for the kernels (fp16 K-major convs, folded BN).  Called
        lazily by forward; call it again after mutating parameters in place."""
        self._engine.prepare()
        return self

    def set_conv_impl(self, impl: int):
        self._engine.impl = impl
        return self

    def forward(self, x, temperature=1.0, keep=None, forced=None, gumbel_noise=None):
        """gumbel_noise: per block (channel sample [B,2G] | None, spatial sample [B,2g,S,S] | None).  With it the gates
        run the reference's TRAINING branch - hard Gumbel-softmax at `temperature` (utils.py:56-58,123-125) - on the
        supplied samples; BatchNorm uses its running statistics (eval), the configuration in which the mmdet backbones
        train the gates (lad_mmdet_resnet.py, norm_eval).  Full training mode (batch-statistics BN, backward) is not
        part of the CUDA path."""
        if self.training and gumbel_noise is None:
            raise LaudError("ResNet: training mode draws Gumbel noise from torch's generator (reference utils.py:56-58); pass "
                            "the samples as gumbel_noise=[(channel, spatial), ...] or call .eval()")
        logits, stats = self._engine.forward(x, keep, forced=forced, gumbel_noise=gumbel_noise, temperature=temperature)
        r3, r2, r1, rc, perc, flops = self._engine.split_stats(stats)
        return logits, r3, r2, r1, rc, perc, flops

    def forward_logits(self, x):
        """Logits only (serving path): skips building the statistics views."""
        if self.training:
            raise LaudError("ResNet: call .eval() first")
        return self._engine.forward(x)[0]

    def capture(self, x_example):
        """CUDA-graph the eval forward for this input shape (see _engine.GraphedForward)."""
        if self.training:
            raise LaudError("ResNet: call .eval() first")
        return self._engine.capture(x_example)

    def get_optim_policies(self):
        """Same two parameter groups as the reference (laud_resnet.py:365-401)."""
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


def _resnet(arch, block, layers, pretrained, progress, **kwargs):
    if pretrained:
        raise LaudError("pretrained=True needs network access to the torchvision model zoo; "
                        "load a checkpoint with load_state_dict instead")
    return ResNet(block, layers, **kwargs)


def uni_resnet50(pretrained=False, progress=True, **kwargs):
    return _resnet("resnet50", Bottleneck, [3, 4, 6, 3], pretrained, progress, **kwargs)


def uni_resnet101(pretrained=False, progress=True, **kwargs):
    return _resnet("resnet101", Bottleneck, [3, 4, 23, 3], pretrained, progress, **kwargs)
