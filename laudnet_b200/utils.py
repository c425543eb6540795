"""Drop-in mirrors of the reference's dynamic operators, backed by liblaud_b200.so.

Same class / function names, constructor arguments, parameter names
(`state_dict` keys) and return values as
`imagenet_classification/models/utils.py` of the reference:

    Masker_spatial(in_channels, mask_channel_group, mask_size)   utils.py:35-65
    ExpandMask(stride, padding=1, mask_channel_group=1)          utils.py:67-89
    Masker_channel_MLP(in_channels, channel_dyn_group, layers, reduction)   :92-131
    Masker_channel_conv_linear(in_channels, channel_dyn_group, reduction)   :133-169
    apply_channel_mask(x, mask) / apply_spatial_mask(x, mask)    utils.py:18-33

`forward` takes what the reference takes (an NCHW float tensor and a
temperature) and returns what it returns `(mask, sparsity, flops)`, but runs
the eval-mode gate as hand-written CUDA on the tensor's device.  There is no
CPU path and no training (Gumbel) path: both raise `LaudError`.

The NHWC/fp16 entry points (`gate_nhwc`) are what the network-level engine
calls; they additionally produce the compact active-index lists the
mask-conditioned convolutions consume.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import LaudError, check, lib, ptr, stream_ptr


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    """Parameter container for a 3x3 convolution (same signature as the reference)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation,
                     groups=groups, bias=False, dilation=dilation)


def conv1x1(in_planes, out_planes, stride=1, bias=False):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=bias)


# ---------------------------------------------------------------------------
# layout plumbing
# ---------------------------------------------------------------------------
def to_nhwc_f16(x: torch.Tensor, ld: Optional[int] = None) -> torch.Tensor:
    """NCHW fp32/fp16 CUDA tensor -> fp16 [B,H,W,ld] (channels last, pitch ld)."""
    _lib.require_cuda(x, "to_nhwc_f16")
    if x.dtype not in (torch.float32, torch.float16):
        raise LaudError(f"to_nhwc_f16: unsupported dtype {x.dtype}")
    x = x.contiguous()
    b, c, h, w = x.shape
    ld = ld or c
    out = torch.zeros if ld != c else torch.empty
    y = out((b, h, w, ld), dtype=torch.float16, device=x.device)
    check(lib().laud_nchw_to_nhwc_f16(ptr(x), int(x.dtype == torch.float32), b, c, h, w, ptr(y), ld, stream_ptr()),
          "laud_nchw_to_nhwc_f16")
    return y


def to_nchw_f32(y: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    """fp16 [B,H,W,ld] -> fp32 NCHW [B,channels,H,W]."""
    b, h, w, ld = y.shape
    c = channels or ld
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=y.device)
    check(lib().laud_nhwc_f16_to_nchw_f32(ptr(y), ld, b, c, h, w, ptr(out), stream_ptr()),
          "laud_nhwc_f16_to_nchw_f32")
    return out


def _no_training(mod: nn.Module, name: str, noise=None) -> None:
    if mod.training and noise is None:
        raise LaudError(f"{name}: training mode draws Gumbel noise from torch's generator (reference utils.py:56-58), which "
                        "a kernel cannot reproduce: pass the sample as `noise=` (same shape as the gate logits), or call "
                        ".eval() for the argmax gate")


def gumbel_noise_like(shape, generator=None, device=None) -> torch.Tensor:
    """A Gumbel(0,1) sample drawn the way F.gumbel_softmax draws it (-log(Exp(1))): the `noise=` input of the
    training-mode gates."""
    return -torch.empty(shape, device=device).exponential_(generator=generator).log()


def gate_from_logits(logits: torch.Tensor, noise: Optional[torch.Tensor], tau: float, G: int, inner: int,
                     mask: torch.Tensor, idx: Optional[torch.Tensor] = None, cnt: Optional[torch.Tensor] = None,
                     total: Optional[torch.Tensor] = None) -> None:
    """laud_gate_from_logits: hard two-way decision from fp32 logits [B,2,G,inner] (+ Gumbel noise) at temperature tau."""
    if noise is not None:
        if noise.numel() != logits.numel():
            raise LaudError(f"Gumbel noise has {noise.numel()} elements, the gate logits {logits.numel()}")
        noise = noise.to(device=logits.device, dtype=torch.float32).contiguous()
    check(lib().laud_gate_from_logits(ptr(logits), ptr(noise), logits.shape[0], G, inner, float(tau), ptr(mask), ptr(idx),
                                      ptr(cnt), ptr(total), stream_ptr()), "laud_gate_from_logits")


# ---------------------------------------------------------------------------
# mask application (reference utils.py:18-33) - thin device ops on NCHW tensors
# ---------------------------------------------------------------------------
def apply_channel_mask(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x[b,c] *= mask[b, c // (C/G)]  (consecutive channels share a group)."""
    _lib.require_cuda(x, "apply_channel_mask")
    b, c, h, w = x.shape
    g = mask.shape[1]
    # The product path never materialises this multiply (masked channels are
    # skipped by the gather-GEMM); the helper exists for API parity.
    per_channel = mask.to(x.dtype).repeat_interleave(c // g, dim=1).view(b, c, 1, 1)
    return x * per_channel


def apply_spatial_mask(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(x, "apply_spatial_mask")
    c, g = x.shape[1], mask.shape[1]
    m = mask.to(x.dtype)
    if g > 1 and g != c:
        m = m.repeat_interleave(c // g, dim=1)
    return x * m


# ---------------------------------------------------------------------------
# spatial / layer masker
# ---------------------------------------------------------------------------
class Masker_spatial(nn.Module):
    def __init__(self, in_channels, mask_channel_group, mask_size):
        super().__init__()
        self.mask_channel_group = mask_channel_group
        self.mask_size = mask_size
        self.conv = conv1x1(in_channels, mask_channel_group * 2, bias=True)
        self.conv_flops_pp = self.conv.weight.shape[0] * self.conv.weight.shape[1] + self.conv.weight.shape[1]
        with torch.no_grad():   # the reference's keep-biased init (utils.py:42-43), same off-by-one
            self.conv.bias[:mask_channel_group] = 5.0
            self.conv.bias[mask_channel_group + 1:] = 0.0

    def _weights(self):
        """(weight fp32 [2g, C] contiguous, bias fp32 [2g]) - packed once, see Masker_channel_MLP._weights."""
        pk = getattr(self, "_packed", None)
        if pk is None:
            w = self.conv.weight.detach()
            pk = (w.reshape(w.shape[0], w.shape[1]).float().contiguous(), self.conv.bias.detach().float().contiguous())
            self._packed = pk
        return pk

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._packed = None
        return super()._load_from_state_dict(*a, **kw)

    def gate_nhwc(self, x: torch.Tensor, total: Optional[torch.Tensor] = None,
                  want_logits: bool = False):
        """x fp16 [B,H,W,C] -> (mask u8 [B,g,S,S], logits fp32 [B,2g,S,S] | None)."""
        b, h, w, c = x.shape
        g = self.mask_channel_group
        s = self.mask_size if self.mask_size < h else h
        mask = torch.empty((b, g, s, s), dtype=torch.uint8, device=x.device)
        logits = torch.empty((b, 2 * g, s, s), dtype=torch.float32, device=x.device) if want_logits else None
        wt, bias = self._weights()
        check(lib().laud_masker_spatial(ptr(x), b, h, w, c, ptr(wt), ptr(bias), g, s,
                                        ptr(logits), ptr(mask), ptr(total), stream_ptr()), "laud_masker_spatial")
        return mask, logits

    def forward(self, x, temperature, noise=None):
        """noise (training mode): Gumbel sample [B, 2g, S, S] -> the hard Gumbel-softmax gate of utils.py:56-58."""
        _no_training(self, "Masker_spatial", noise)
        _lib.require_cuda(x, "Masker_spatial")
        b, c, h, w = x.shape
        total = torch.zeros(1, dtype=torch.int32, device=x.device)
        if self.training:
            mask_u8, logits = self.gate_nhwc(to_nhwc_f16(x), None, want_logits=True)
            g, s_ = self.mask_channel_group, mask_u8.shape[-1]
            gate_from_logits(logits, noise, temperature, g, s_ * s_, mask_u8, total=total)
        else:
            mask_u8, _ = self.gate_nhwc(to_nhwc_f16(x), total)
        s = mask_u8.shape[-1]
        flops = c * s * s + self.conv_flops_pp * s * s
        mask = mask_u8.float()
        sparsity = total[0].float() / float(mask_u8.numel())
        return mask, sparsity, flops


class ExpandMask(nn.Module):
    def __init__(self, stride, padding=1, mask_channel_group=1):
        super().__init__()
        self.stride = stride
        self.padding = padding
        self.mask_channel_group = mask_channel_group

    def expand_u8(self, mask_u8: torch.Tensor, total: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, g, h, w = mask_u8.shape
        out = torch.empty((b, g, h * self.stride, w * self.stride), dtype=torch.uint8, device=mask_u8.device)
        check(lib().laud_expand_mask(ptr(mask_u8), b, g, h, w, self.stride, self.padding, ptr(out), ptr(total),
                                     stream_ptr()), "laud_expand_mask")
        return out

    def forward(self, x):
        _lib.require_cuda(x, "ExpandMask")
        m = (x > 0.5).to(torch.uint8).contiguous() if x.dtype != torch.uint8 else x.contiguous()
        return self.expand_u8(m).bool()


# ---------------------------------------------------------------------------
# channel maskers
# ---------------------------------------------------------------------------
class _ChannelGate:
    """Outputs of a channel gate on the device."""
    __slots__ = ("mask", "idx", "cnt", "logits", "pooled")

    def __init__(self, mask, idx, cnt, logits, pooled):
        self.mask, self.idx, self.cnt, self.logits, self.pooled = mask, idx, cnt, logits, pooled


class Masker_channel_MLP(nn.Module):
    def __init__(self, in_channels, channel_dyn_group, layers=2, reduction=16):
        super().__init__()
        assert layers in [1, 2]
        self.channel_dyn_group = channel_dyn_group
        self.layers = layers
        width = max(channel_dyn_group // reduction, 16)
        self.conv = nn.Sequential(
            nn.Linear(in_channels, width), nn.ReLU(), nn.Linear(width, channel_dyn_group * 2, bias=True)
        ) if layers == 2 else nn.Linear(in_channels, channel_dyn_group * 2, bias=True)
        self.conv_flops = in_channels * width + width * channel_dyn_group * 2 if layers == 2 \
            else in_channels * channel_dyn_group * 2
        last = self.conv[-1] if layers == 2 else self.conv
        with torch.no_grad():   # reference init (utils.py:106-111), same off-by-one
            last.bias[:channel_dyn_group] = 2.0
            last.bias[channel_dyn_group + 1:] = -2.0

    def _weights(self):
        """fp32 contiguous views of the MLP parameters (the kernels read `const float*`): the parameters themselves
        for an fp32 model, packed copies after `.half()` / `.bfloat16()`.  Cached until the parameters move or are
        reloaded (`_apply`, load_state_dict) - the engine's `prepare()` refreshes the cache."""
        pk = getattr(self, "_packed", None)
        if pk is None:
            f = lambda t: t.detach().float().contiguous()
            if self.layers == 2:
                l1, l2 = self.conv[0], self.conv[2]
                pk = (f(l1.weight), f(l1.bias), l1.weight.shape[0], f(l2.weight), f(l2.bias))
            else:
                pk = (f(self.conv.weight), f(self.conv.bias), 0, None, None)
            self._packed = pk
        return pk

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._packed = None
        return super()._load_from_state_dict(*a, **kw)

    def gate_nhwc(self, x: torch.Tensor, total: Optional[torch.Tensor] = None, want_logits: bool = False,
                  out: Optional[_ChannelGate] = None, partial_ws: Optional[torch.Tensor] = None) -> _ChannelGate:
        """x fp16 [B,H,W,C] (or [B,HW,C]).  One fused launch pair: GAP partials, then
        MLP + decision + ordered index compaction."""
        b, c = x.shape[0], x.shape[-1]
        hw = x.numel() // (b * c)
        G = self.channel_dyn_group
        dev = x.device
        if out is None:
            out = _ChannelGate(torch.empty((b, G), dtype=torch.uint8, device=dev),
                               torch.empty((b, G), dtype=torch.int32, device=dev),
                               torch.empty((b,), dtype=torch.int32, device=dev),
                               torch.empty((b, 2 * G), dtype=torch.float32, device=dev) if want_logits else None,
                               torch.empty((b, c), dtype=torch.float32, device=dev) if want_logits else None)
        if partial_ws is None:
            partial_ws = torch.empty((b, _lib.GAP_SPLITS, c), dtype=torch.float32, device=dev)
        w1, b1, hidden, w2, b2 = self._weights()
        check(lib().laud_masker_channel_mlp(ptr(x), b, hw, c, self.layers, ptr(w1), ptr(b1), hidden, ptr(w2), ptr(b2),
                                            G, ptr(partial_ws), ptr(out.pooled), ptr(out.logits), ptr(out.mask),
                                            ptr(out.idx), ptr(out.cnt), ptr(total), stream_ptr()),
              "laud_masker_channel_mlp")
        return out

    def gate_from_partials(self, partials: torch.Tensor, b: int, hw: int, c: int, gap_tiles: int,
                           total: Optional[torch.Tensor], out: _ChannelGate) -> _ChannelGate:
        """Same decision from the fused-GAP partial sums [B,gap_tiles,C] the producing convolution left
        (laud_conv_desc::gap_partial): the activations are not read again."""
        w1, b1, hidden, w2, b2 = self._weights()
        check(lib().laud_masker_channel_from_partials(ptr(partials), b, hw, c, gap_tiles, self.layers, ptr(w1), ptr(b1),
                                                      hidden, ptr(w2), ptr(b2), self.channel_dyn_group, ptr(out.pooled),
                                                      ptr(out.logits), ptr(out.mask), ptr(out.idx), ptr(out.cnt),
                                                      ptr(total), stream_ptr()), "laud_masker_channel_from_partials")
        return out

    def forward(self, x, temperature, noise=None):
        """noise (training mode): Gumbel sample [B, 2G] -> the hard Gumbel-softmax gate of utils.py:123-125."""
        _no_training(self, "Masker_channel_MLP", noise)
        _lib.require_cuda(x, "Masker_channel_MLP")
        b, c, h, w = x.shape
        total = torch.zeros(1, dtype=torch.int32, device=x.device)
        if self.training:
            gate = self.gate_nhwc(to_nhwc_f16(x), None, want_logits=True)
            gate_from_logits(gate.logits, noise, temperature, self.channel_dyn_group, 1, gate.mask, gate.idx, gate.cnt, total)
        else:
            gate = self.gate_nhwc(to_nhwc_f16(x), total)
        flops = c * h * w + self.conv_flops
        mask = gate.mask.float()
        sparsity = total[0].float() / float(gate.mask.numel())
        return mask, sparsity, flops


class Masker_channel_conv_linear(nn.Module):
    def __init__(self, in_channels, channel_dyn_group, reduction=16):
        super().__init__()
        self.channel_dyn_group = channel_dyn_group
        self.conv = nn.Sequential(conv1x1(in_channels, in_channels // reduction),
                                  nn.BatchNorm2d(in_channels // reduction), nn.ReLU())
        self.linear = nn.Linear(in_channels // reduction, channel_dyn_group * 2, bias=True)
        with torch.no_grad():
            self.linear.bias[:channel_dyn_group] = 2.0
            self.linear.bias[channel_dyn_group + 1:] = -2.0
        self.masker_flops = in_channels * in_channels // reduction + in_channels // reduction * channel_dyn_group * 2

    def _weights(self):
        """Packed once (fp16 K-major 1x1 conv weight, folded BN scale / shift, fp32 linear weight / bias): the gate is
        then allocation-free apart from its workspaces and safe to capture in a CUDA graph."""
        pk = getattr(self, "_packed", None)
        if pk is None:
            from ._engine import fold_bn, pack_conv_weight     # local import: engine depends on this module
            scale, shift = fold_bn(self.conv[1])
            wc = pack_conv_weight(self.conv[0].weight)
            wl = self.linear.weight.detach().float().contiguous()
            cr = wc.shape[0]
            crp = (cr + 7) // 8 * 8
            if crp != cr:
                # the fp16 kernels take channel counts in multiples of 8 (64 input channels / reduction 16 = 4, the
                # reference Bottleneck default at stage 1): pad with channels that are exactly 0 after BN + ReLU
                # (zero weights, scale = shift = 0) and give them zero columns in the linear layer
                pad = crp - cr
                wc = torch.cat([wc, wc.new_zeros((pad,) + tuple(wc.shape[1:]))]).contiguous()
                scale = torch.cat([scale, scale.new_zeros(pad)]).contiguous()
                shift = torch.cat([shift, shift.new_zeros(pad)]).contiguous()
                wl = torch.cat([wl, wl.new_zeros(wl.shape[0], pad)], dim=1).contiguous()
            pk = (wc, scale, shift, wl, self.linear.bias.detach().float().contiguous())
            self._packed = pk
        return pk

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._packed = None
        return super()._load_from_state_dict(*a, **kw)

    def gate_nhwc(self, x: torch.Tensor, total: Optional[torch.Tensor] = None, want_logits: bool = False,
                  out: Optional[_ChannelGate] = None, partial_ws=None, impl: int = _lib.CONV_AUTO,
                  z_ws: Optional[torch.Tensor] = None, pooled_ws: Optional[torch.Tensor] = None) -> _ChannelGate:
        """1x1 conv + BN + ReLU (the conv kernel) -> deterministic GAP -> Linear -> decision (utils.py:150-169).
        z_ws / pooled_ws / partial_ws: caller-owned workspaces (fp16 >= B*H*W*C/r, fp32 >= B*C/r, fp32 >=
        B*GAP_SPLITS*C/r); allocated here when absent."""
        from ._engine import run_conv
        b, h, w, c = x.shape
        G = self.channel_dyn_group
        dev = x.device
        wc, scale, shift, wl, bl = self._weights()
        cr = wc.shape[0]                           # reduced width, padded to a multiple of 8
        z = (z_ws[:b * h * w * cr] if z_ws is not None else torch.empty(b * h * w * cr, dtype=torch.float16, device=dev)).view(b, h, w, cr)
        run_conv(x, wc, z, b, h, w, c, h, w, cr, 1, 1, 0, scale=scale, shift=shift, relu=_lib.RELU_ALL, impl=impl,
                 tag="masker.conv")
        pooled = (pooled_ws[:b * cr] if pooled_ws is not None else torch.empty(b * cr, dtype=torch.float32, device=dev)).view(b, cr)
        pws = partial_ws if partial_ws is not None else torch.empty((b, _lib.GAP_SPLITS, cr), dtype=torch.float32, device=dev)
        check(lib().laud_global_avg_pool(ptr(z), b, h * w, cr, cr, ptr(pws), ptr(pooled), stream_ptr()),
              "laud_global_avg_pool")
        if out is None:
            out = _ChannelGate(torch.empty((b, G), dtype=torch.uint8, device=dev),
                               torch.empty((b, G), dtype=torch.int32, device=dev),
                               torch.empty((b,), dtype=torch.int32, device=dev),
                               torch.empty((b, 2 * G), dtype=torch.float32, device=dev) if want_logits else None,
                               pooled)
        check(lib().laud_masker_channel_from_pooled(ptr(pooled), b, cr, 1, ptr(wl), ptr(bl), 0, None, None, G,
                                                    ptr(out.logits), ptr(out.mask), ptr(out.idx), ptr(out.cnt),
                                                    ptr(total), stream_ptr()), "laud_masker_channel_from_pooled")
        return out

    def forward(self, x, temperature, noise=None):
        """noise (training mode, BatchNorm frozen - its running statistics are used, as the mmdet backbones run it):
        Gumbel sample [B, 2G] -> the hard Gumbel-softmax gate of utils.py:161-163."""
        _no_training(self, "Masker_channel_conv_linear", noise)
        _lib.require_cuda(x, "Masker_channel_conv_linear")
        b, c, h, w = x.shape
        total = torch.zeros(1, dtype=torch.int32, device=x.device)
        if self.training:
            gate = self.gate_nhwc(to_nhwc_f16(x), None, want_logits=True)
            gate_from_logits(gate.logits, noise, temperature, self.channel_dyn_group, 1, gate.mask, gate.idx, gate.cnt, total)
        else:
            gate = self.gate_nhwc(to_nhwc_f16(x), total)
        cr = self.conv[0].weight.shape[0]
        flops = cr * h * w + self.masker_flops
        mask = gate.mask.float()
        sparsity = total[0].float() / float(gate.mask.numel())
        return mask, sparsity, flops
