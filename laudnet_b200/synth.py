"""Deterministic synthetic weights / inputs for the LAUD hot path.

There is no network access for the model-zoo checkpoints (reference
README.md:61-64), so parity tests and the benchmark run on seeded synthetic
parameters.  Everything here is numpy `RandomState`-based so that the build
container (where the golden fixtures are generated from the reference) and the
GPU box produce bit-identical parameters from the same seed, independent of
torch's RNG implementation.

Recipe:
  * backbone convs: Kaiming-normal fan_out, as the reference initialises them
    (imagenet_classification/models/laud_resnet.py:255-260);
  * BatchNorm: gamma~N(1,.2), beta~N(0,.5), running_mean~N(0,.5),
    running_var~U(.5,2) so the mask-before-BN constants relu(bn(0)) are
    non-trivial;
  * masker layers: torch's default Linear/Conv init (U(+-1/sqrt(fan_in)));
    the final 2-way bias is then *calibrated* (see `calibrate_two_way_bias`) so
    the eval-mode activation rate hits the requested target - random init
    alone gives density ~1.0 because of the +2/-2 (+5/0) keep bias
    (imagenet_classification/models/utils.py:42-43,107-108).
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.RandomState:
    return np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 32))


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """One parameter/buffer, chosen by its state_dict key."""
    r = _rng(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    is_masker = "masker" in name
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if leaf == "running_mean":
        a = r.standard_normal(shape) * 0.5
    elif leaf == "running_var":
        a = r.uniform(0.5, 2.0, size=shape)
    elif leaf == "weight" and len(shape) == 1:          # BN gamma
        a = 1.0 + 0.2 * r.standard_normal(shape)
    elif leaf == "bias" and not is_masker and name != "fc.bias" and ".se." not in name:   # BN beta
        a = 0.5 * r.standard_normal(shape)
    elif leaf == "weight" and len(shape) == 4 and not is_masker:
        fan_out = shape[0] * shape[2] * shape[3]
        a = r.standard_normal(shape) * np.sqrt(2.0 / fan_out)
    elif leaf == "weight":                              # Linear / masker conv
        fan_in = int(np.prod(shape[1:]))
        a = r.uniform(-1.0, 1.0, size=shape) / np.sqrt(fan_in)
    else:                                               # Linear / masker / SE bias
        a = r.uniform(-1.0, 1.0, size=shape) * 0.1
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))


def synth_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int,
                     fp16_weights: bool = True) -> Dict[str, torch.Tensor]:
    """Fill every key of `shapes` (name -> shape).  With `fp16_weights` the
    convolution / fc weights are rounded to fp16-representable values (kept in
    fp32), which is the model-preparation step of the CUDA path: the oracle is
    then fed exactly the weights the tensor cores see."""
    sd = {}
    for name, shape in shapes.items():
        t = synth_tensor(name, tuple(shape), seed)
        if fp16_weights and t.dtype == torch.float32 and t.dim() in (2, 4) and "masker" not in name:
            t = t.half().float()
        sd[name] = t
    return sd


def synth_images(batch: int, size: int, seed: int, start: int = 0) -> torch.Tensor:
    """fp16-representable images with per-sample gain/offset so that pooled
    features (and therefore gating decisions) differ between samples.
    Sample i depends only on (seed, start+i): shards of a batch agree with the
    unsharded batch."""
    out = np.empty((batch, 3, size, size), dtype=np.float32)
    for i in range(batch):
        r = np.random.RandomState((seed * 7919 + 104729 * (start + i) + 1) % (2 ** 32))
        gain = r.uniform(0.5, 1.5)
        offs = r.standard_normal((3, 1, 1)) * 0.5
        yy, xx = np.meshgrid(np.linspace(-1, 1, size), np.linspace(-1, 1, size), indexing="ij")
        tilt = r.standard_normal((3, 1, 1)) * yy[None] + r.standard_normal((3, 1, 1)) * xx[None]
        out[i] = gain * r.standard_normal((3, size, size)) + offs + 0.5 * tilt
    return torch.from_numpy(out).half().float()


def calibrate_two_way_bias(margin: torch.Tensor, rate: float, per_group: bool) -> torch.Tensor:
    """Bias shift that makes a 2-way gate fire at `rate`.

    `margin` is keep-logit minus drop-logit, shape [N, G] (N = samples or
    samples*positions).  Returns delta[G] to SUBTRACT from the keep bias so
    that a fraction `rate` of the margins stay >= 0.  The threshold sits midway
    between two order statistics, so no calibration sample lands on a tie.
    """
    m = margin.detach().double().cpu()
    if not per_group:
        m = m.reshape(-1, 1)
    n = m.shape[0]
    k = int(round((1.0 - rate) * n))            # number of samples to switch off
    srt, _ = torch.sort(m, dim=0)
    if k <= 0:
        thr = srt[0] - 1.0
    elif k >= n:
        thr = srt[-1] + 1.0
    else:
        thr = 0.5 * (srt[k - 1] + srt[k])
    thr = thr.float()
    return thr if per_group else thr.expand(margin.shape[1]).clone()


# --------------------------------------------------------------------------
# Data-driven calibration: the stand-in for training.
# --------------------------------------------------------------------------
def _set_bn_from_data(sd, prefix: str, z: torch.Tensor, seed: int) -> None:
    """Point a BatchNorm's running statistics at the statistics of the tensor it
    will actually normalise (jittered), the way a trained network's would be.
    Without this a randomly initialised 33-block residual chain is
    linear-homogeneous in its input and overflows fp16."""
    mean = z.mean(dim=(0, 2, 3))
    var = z.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6)
    r = _rng(seed, prefix + "jitter")
    c = mean.numel()
    jm = torch.from_numpy(r.standard_normal(c).astype(np.float32)).to(z.device)
    jv = torch.from_numpy(r.uniform(0.5, 2.0, size=c).astype(np.float32)).to(z.device)
    sd[prefix + "running_mean"] = (mean + 0.5 * var.sqrt() * jm).float().cpu()
    sd[prefix + "running_var"] = (var * jv).float().cpu()


def _bn_eval(z, sd, prefix, eps=1e-5):
    dev = z.device
    g, b = sd[prefix + "weight"].to(dev), sd[prefix + "bias"].to(dev)
    m, v = sd[prefix + "running_mean"].to(dev), sd[prefix + "running_var"].to(dev)
    scale = g / torch.sqrt(v + eps)
    return z * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def calibrate_resnet(sd: Dict[str, torch.Tensor], geoms, images: torch.Tensor, seed: int,
                     channel_rate: float = 0.6, spatial_rate: float = 0.4,
                     layer_rate: float = 0.47) -> Dict[str, torch.Tensor]:
    """One sequential pass over a calibration batch that (i) sets every
    BatchNorm's running stats from data and (ii) shifts every masker's keep
    bias so that its gate fires at the requested rate (per channel group for
    the channel masker; per mask group for the spatial/layer masker).

    `geoms` is a list of objects with the attributes of a bottleneck
    (prefix, inplanes, width, stride, output_size, mask_size, dyn_mode,
    groups_channel, masker_kind, masker_layers, has_downsample).
    Uses stock torch ops on `images.device`: this is weight *synthesis* (it
    replaces training), it is not part of the inference hot path.
    """
    import torch.nn.functional as F
    dev = images.device
    W = lambda k: sd[k].to(dev)
    with torch.no_grad():
        z = F.conv2d(images.float(), W("conv1.weight"), stride=2, padding=3)
        _set_bn_from_data(sd, "bn1.", z, seed)
        x = F.max_pool2d(torch.relu(_bn_eval(z, sd, "bn1.")), 3, 2, 1)
        for g in geoms:
            p = g.prefix
            b = x.shape[0]
            cmask = None
            m3 = None
            if g.dyn_mode in ("channel", "both"):
                mp = p + "masker_channel."
                G = g.groups_channel
                if g.masker_kind == "MLP":
                    pooled = x.mean(dim=(2, 3))
                    if g.masker_layers == 2:
                        h = torch.relu(F.linear(pooled, W(mp + "conv.0.weight"), W(mp + "conv.0.bias")))
                        bias_key = mp + "conv.2.bias"
                        logits = F.linear(h, W(mp + "conv.2.weight"), W(bias_key))
                    else:
                        bias_key = mp + "conv.bias"
                        logits = F.linear(pooled, W(mp + "conv.weight"), W(bias_key))
                else:
                    zc = F.conv2d(x, W(mp + "conv.0.weight"))
                    _set_bn_from_data(sd, mp + "conv.1.", zc, seed)
                    pooled = torch.relu(_bn_eval(zc, sd, mp + "conv.1.")).mean(dim=(2, 3))
                    bias_key = mp + "linear.bias"
                    logits = F.linear(pooled, W(mp + "linear.weight"), W(bias_key))
                margin = logits[:, :G] - logits[:, G:]
                delta = calibrate_two_way_bias(margin, channel_rate, per_group=True).to(dev)
                nb = sd[bias_key].clone()
                nb[:G] -= delta.cpu()
                sd[bias_key] = nb
                cmask = ((margin - delta) >= 0).float()
                cmask = cmask.repeat_interleave(g.width // G, dim=1).view(b, g.width, 1, 1)
            if g.dyn_mode in ("spatial", "layer", "both"):
                wk, bk = p + "masker_spatial.conv.weight", p + "masker_spatial.conv.bias"
                q = F.adaptive_avg_pool2d(x, g.mask_size) if g.mask_size < x.shape[2] else x
                logits = F.conv2d(q, W(wk), W(bk))
                gs = logits.shape[1] // 2
                margin = logits[:, :gs] - logits[:, gs:]
                rate = layer_rate if g.dyn_mode == "layer" else spatial_rate
                flat = margin.permute(0, 2, 3, 1).reshape(-1, gs)
                delta = calibrate_two_way_bias(flat, rate, per_group=True).to(dev)
                nb = sd[bk].clone()
                nb[:gs] -= delta.cpu()
                sd[bk] = nb
                small = ((margin - delta.view(1, gs, 1, 1)) >= 0).float()
                S = small.shape[-1]
                idx = torch.div(torch.arange(g.output_size, device=dev) * S, g.output_size, rounding_mode="floor")
                m3 = small[:, :, idx][:, :, :, idx]
                cout = g.outplanes
                if gs > 1 and gs != cout:
                    m3 = m3.repeat_interleave(cout // gs, dim=1)
            z1 = F.conv2d(x, W(p + "conv1.weight"))
            _set_bn_from_data(sd, p + "bn1.", z1, seed)
            a1 = torch.relu(_bn_eval(z1 * cmask if cmask is not None else z1, sd, p + "bn1."))
            z2 = F.conv2d(a1, W(p + "conv2.weight"), stride=g.stride, padding=1)
            _set_bn_from_data(sd, p + "bn2.", z2, seed)
            a2 = torch.relu(_bn_eval(z2 * cmask if cmask is not None else z2, sd, p + "bn2."))
            z3 = F.conv2d(a2, W(p + "conv3.weight"))
            _set_bn_from_data(sd, p + "bn3.", z3, seed)
            y = _bn_eval(z3, sd, p + "bn3.")
            if m3 is not None:
                y = y * m3
            ident = x
            if g.has_downsample:
                zd = F.conv2d(x, W(p + "downsample.0.weight"), stride=g.stride)
                _set_bn_from_data(sd, p + "downsample.1.", zd, seed)
                ident = _bn_eval(zd, sd, p + "downsample.1.")
            x = torch.relu(y + ident)
    return sd


def calibrate_regnet(sd: Dict[str, torch.Tensor], geoms, images: torch.Tensor, seed: int,
                     channel_rate: float = 0.6, spatial_rate: float = 0.4) -> Dict[str, torch.Tensor]:
    """`calibrate_resnet` for the LAUD-RegNet-Y trunk (reference laud_regnet.py): one sequential pass that sets every
    BatchNorm's running statistics from data and shifts every masker's keep bias to the requested firing rate.
    `geoms`: objects with the attributes of oracle.RegBlockGeom."""
    import torch.nn.functional as F
    dev = images.device
    W = lambda k: sd[k].to(dev)
    with torch.no_grad():
        z = F.conv2d(images.float(), W("stem.0.weight"), stride=2, padding=1)
        _set_bn_from_data(sd, "stem.1.", z, seed)
        x = torch.relu(_bn_eval(z, sd, "stem.1."))
        for g in geoms:
            p = g.prefix + "f."
            b = x.shape[0]
            cmask = None
            m3 = None
            if g.dyn_mode in ("channel", "both"):
                mp = p + "masker_channel."
                G = g.groups_channel
                if g.masker_kind == "MLP":
                    pooled = x.mean(dim=(2, 3))
                    if g.masker_layers == 2:
                        h = torch.relu(F.linear(pooled, W(mp + "conv.0.weight"), W(mp + "conv.0.bias")))
                        bias_key = mp + "conv.2.bias"
                        logits = F.linear(h, W(mp + "conv.2.weight"), W(bias_key))
                    else:
                        bias_key = mp + "conv.bias"
                        logits = F.linear(pooled, W(mp + "conv.weight"), W(bias_key))
                else:
                    zc = F.conv2d(x, W(mp + "conv.0.weight"))
                    _set_bn_from_data(sd, mp + "conv.1.", zc, seed)
                    pooled = torch.relu(_bn_eval(zc, sd, mp + "conv.1.")).mean(dim=(2, 3))
                    bias_key = mp + "linear.bias"
                    logits = F.linear(pooled, W(mp + "linear.weight"), W(bias_key))
                margin = logits[:, :G] - logits[:, G:]
                delta = calibrate_two_way_bias(margin, channel_rate, per_group=True).to(dev)
                nb = sd[bias_key].clone()
                nb[:G] -= delta.cpu()
                sd[bias_key] = nb
                cmask = ((margin - delta) >= 0).float()
                cmask = cmask.repeat_interleave(g.w_b // G, dim=1).view(b, g.w_b, 1, 1)
            if g.dyn_mode in ("spatial", "both"):
                wk, bk = p + "masker_spatial.conv.weight", p + "masker_spatial.conv.bias"
                q = F.adaptive_avg_pool2d(x, g.mask_size) if g.mask_size < x.shape[2] else x
                logits = F.conv2d(q, W(wk), W(bk))
                gs = logits.shape[1] // 2
                margin = logits[:, :gs] - logits[:, gs:]
                flat = margin.permute(0, 2, 3, 1).reshape(-1, gs)
                delta = calibrate_two_way_bias(flat, spatial_rate, per_group=True).to(dev)
                nb = sd[bk].clone()
                nb[:gs] -= delta.cpu()
                sd[bk] = nb
                small = ((margin - delta.view(1, gs, 1, 1)) >= 0).float()
                S = small.shape[-1]
                idx = torch.div(torch.arange(g.output_size, device=dev) * S, g.output_size, rounding_mode="floor")
                m3 = small[:, :, idx][:, :, :, idx]
                if gs > 1 and gs != g.w_out:
                    m3 = m3.repeat_interleave(g.w_out // gs, dim=1)
            za = F.conv2d(x, W(p + "a.0.weight"))
            _set_bn_from_data(sd, p + "a.1.", za, seed)
            a1 = torch.relu(_bn_eval(za, sd, p + "a.1."))
            if cmask is not None:
                a1 = a1 * cmask
            zb = F.conv2d(a1, W(p + "b.0.weight"), stride=g.stride, padding=1, groups=g.conv_groups)
            _set_bn_from_data(sd, p + "b.1.", zb, seed)
            a2 = torch.relu(_bn_eval(zb, sd, p + "b.1."))
            if cmask is not None:
                a2 = a2 * cmask
            sq = a2.mean(dim=(2, 3), keepdim=True)
            sq = torch.relu(F.conv2d(sq, W(p + "se.fc1.weight"), W(p + "se.fc1.bias")))
            a2 = a2 * torch.sigmoid(F.conv2d(sq, W(p + "se.fc2.weight"), W(p + "se.fc2.bias")))
            zc3 = F.conv2d(a2, W(p + "c.0.weight"))
            _set_bn_from_data(sd, p + "c.1.", zc3, seed)
            y = _bn_eval(zc3, sd, p + "c.1.")
            if m3 is not None:
                y = y * m3
            ident = x
            if g.has_proj:
                zd = F.conv2d(x, W(g.prefix + "proj.0.weight"), stride=g.stride)
                _set_bn_from_data(sd, g.prefix + "proj.1.", zd, seed)
                ident = _bn_eval(zd, sd, g.prefix + "proj.1.")
            x = torch.relu(ident + y)
    return sd


# --------------------------------------------------------------------------
# Workload construction shared by bench.py, the tests and smoke().
# --------------------------------------------------------------------------
class _Geom:
    """Bottleneck geometry read off a drop-in module (same attribute names as
    the oracle's BlockGeom, so `calibrate_resnet` accepts either)."""

    def __init__(self, prefix, blk):
        self.prefix = prefix
        self.inplanes = blk.conv1.weight.shape[1]
        self.width = blk.conv1.weight.shape[0]
        self.outplanes = blk.conv3.weight.shape[0]
        self.stride = blk.stride
        self.output_size = blk.output_size
        self.mask_size = blk.mask_size
        self.dyn_mode = blk.dyn_mode
        self.groups_channel = blk.channel_dyn_group
        self.groups_spatial = blk.spatial_mask_channel_group
        mk = blk.masker_channel
        self.masker_kind = "MLP" if (mk is None or hasattr(mk, "layers")) else "conv_linear"
        self.masker_layers = getattr(mk, "layers", 2)
        self.has_downsample = blk.downsample is not None


def geometry_of(model):
    return [_Geom(f"layer{s + 1}.{i}.", blk)
            for s in range(4) for i, blk in enumerate(getattr(model, f"layer{s + 1}"))]


class _RegGeom:
    """ResBottleneckBlock geometry read off a drop-in LAD_RegNet (attribute names of the oracle's RegBlockGeom)."""

    def __init__(self, prefix, blk):
        f = blk.f
        self.prefix = prefix
        self.w_in, self.w_out, self.stride = blk.width_in, blk.width_out, blk.stride
        self.w_b = f.a[0].weight.shape[0]
        self.conv_groups = f.b[0].groups
        self.output_size, self.mask_size, self.dyn_mode = f.output_size, f.mask_size, f.dyn_mode
        self.groups_channel, self.groups_spatial = f.channel_dyn_group, f.spatial_mask_channel_group
        mk = f.masker_channel
        self.masker_kind = "MLP" if (mk is None or hasattr(mk, "layers")) else "conv_linear"
        self.masker_layers = getattr(mk, "layers", 2)
        self.has_proj = blk.proj is not None
        self.se_width = f.se.fc1.weight.shape[0]


def regnet_geometry_of(model):
    return [_RegGeom(f"trunk_output.{sname}.{bname}.", blk)
            for sname, stage in model.trunk_output.named_children() for bname, blk in stage.named_children()]


SPATIAL_KWARGS = dict(           # spatial 4-4-2-1, one mask group (BASELINE configs[0] / [4])
    input_size=224, dyn_mode=["spatial"] * 4, mask_spatial_granularity=[4, 4, 2, 1],
    spatial_mask_channel_group=[1] * 4, channel_dyn_granularity=[1] * 4, channel_masker=["MLP"] * 4,
    channel_masker_layers=[2] * 4, reduction_ratio=[16] * 4)
LAYER_KWARGS = dict(SPATIAL_KWARGS, dyn_mode=["layer"] * 4, mask_spatial_granularity=[56, 28, 14, 7])   # train_scripts.sh:22


HEADLINE_KWARGS = dict(          # LAUD-ResNet101 channel-2222
    input_size=224, dyn_mode=["channel"] * 4, channel_dyn_granularity=[2, 2, 2, 2],
    channel_masker=["MLP"] * 4, channel_masker_layers=[2] * 4, reduction_ratio=[16] * 4,
    spatial_mask_channel_group=[1] * 4, mask_spatial_granularity=[4, 4, 2, 1], lr_mult=1.0)


def synth_calibrated_state_dict(model, seed: int, calib_images: torch.Tensor, channel_rate: float = 0.6,
                                spatial_rate: float = 0.4, layer_rate: float = 0.47):
    """Seeded weights for `model`'s architecture, BN statistics and gate biases
    calibrated on `calib_images` (on that tensor's device, stock torch ops)."""
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(shapes, seed)
    if hasattr(model, "trunk_output"):           # LAUD-RegNet-Y
        return calibrate_regnet(sd, regnet_geometry_of(model), calib_images, seed, channel_rate=channel_rate,
                                spatial_rate=spatial_rate)
    return calibrate_resnet(sd, geometry_of(model), calib_images, seed, channel_rate=channel_rate,
                            spatial_rate=spatial_rate, layer_rate=layer_rate)


# --------------------------------------------------------------------------
# AdaViT (BASELINE configs[3]): seeded DeiT weights + policy heads calibrated to target keep rates
# --------------------------------------------------------------------------
ADAVIT_RATES = dict(token_rate=0.65, head_rate=0.7, layer_rate=0.85)   # AdaViT reports ~2x fewer FLOPs on DeiT-S


def synth_adavit_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    r = _rng(seed, "adavit." + name)
    leaf = name.rsplit(".", 1)[-1]
    if name in ("cls_token", "pos_embed"):
        a = 0.5 * r.standard_normal(shape)
    elif leaf == "weight" and len(shape) == 1:                 # LayerNorm gamma
        a = 1.0 + 0.1 * r.standard_normal(shape)
    elif leaf == "bias" and ("norm" in name):                  # LayerNorm beta
        a = 0.1 * r.standard_normal(shape)
    elif leaf == "weight":                                     # Linear / patch conv
        fan_in = int(np.prod(shape[1:]))
        gain = 0.5 if (".proj." in name and "patch_embed" not in name) or ".fc2." in name else 1.0   # tame residual growth
        a = gain * r.standard_normal(shape) / np.sqrt(fan_in)
    else:                                                      # Linear bias
        a = 0.1 * r.standard_normal(shape)
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))


def synth_adavit_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    """Backbone Linear / conv weights are rounded to fp16-representable values (the oracle is fed the weights the
    tensor cores see); LayerNorm and policy parameters stay fp32."""
    sd = {}
    for name, shape in shapes.items():
        t = synth_adavit_tensor(name, tuple(shape), seed)
        if t.dim() >= 2 and "_select" not in name and name not in ("cls_token", "pos_embed"):
            t = t.half().float()
        sd[name] = t
    return sd


def calibrate_adavit(sd: Dict[str, torch.Tensor], model_kwargs: dict, images: torch.Tensor, token_rate: float = 0.65,
                     head_rate: float = 0.7, layer_rate: float = 0.85) -> Dict[str, torch.Tensor]:
    """Shift the biases of the policy heads, block by block on a calibration batch, so that the eval decisions fire at
    the target rates (stands in for training the policies).  Plain torch on `images.device`; weight synthesis only."""
    import torch.nn.functional as F
    dev = images.device
    s = {k: v.to(dev) for k, v in sd.items()}
    D, H, P = model_kwargs["embed_dim"], model_kwargs["num_heads"], model_kwargs["patch_size"]
    depth, keep_layers = model_kwargs["depth"], model_kwargs.get("keep_layers", 1)
    d = D // H
    ln = lambda t, p: F.layer_norm(t, (D,), s[p + "weight"], s[p + "bias"], 1e-6)
    with torch.no_grad():
        t = F.conv2d(images.float(), s["patch_embed.proj.weight"], s["patch_embed.proj.bias"], stride=P).flatten(2).transpose(1, 2)
        x = torch.cat([s["cls_token"].expand(t.shape[0], -1, -1), t], 1) + s["pos_embed"]
        B, L, _ = x.shape
        for i in range(depth):
            p = f"blocks.{i}."
            tok = torch.ones(B, L, dtype=torch.bool, device=dev)
            head = torch.ones(B, H, dtype=torch.bool, device=dev)
            layer = torch.ones(B, 2, dtype=torch.bool, device=dev)
            if i >= keep_layers and (p + "norm_policy.weight") in s:
                pt = ln(x[:, 0], p + "norm_policy.")
                if p + "layer_select.weight" in s:
                    lg = F.linear(pt, s[p + "layer_select.weight"])
                    s[p + "layer_select.bias"] = -calibrate_two_way_bias(lg, layer_rate, per_group=True).to(dev)
                    layer = (lg + s[p + "layer_select.bias"]) >= 0
                if p + "head_select.weight" in s:
                    lg = F.linear(pt, s[p + "head_select.weight"])
                    s[p + "head_select.bias"] = -calibrate_two_way_bias(lg, head_rate, per_group=True).to(dev)
                    head = (lg + s[p + "head_select.bias"]) >= 0
                if p + "token_select.weight" in s:
                    lg = F.linear(ln(x[:, 1:], p + "norm1."), s[p + "token_select.weight"]).squeeze(-1)
                    s[p + "token_select.bias"] = -calibrate_two_way_bias(lg.reshape(-1, 1), token_rate, per_group=True).to(dev)
                    tok = torch.cat([tok[:, :1], (lg + s[p + "token_select.bias"]) >= 0], 1)
            qkv = F.linear(ln(x, p + "norm1."), s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"]).view(B, L, 3, H, d).permute(2, 0, 3, 1, 4)
            a = (qkv[0] @ qkv[1].transpose(-1, -2)) * d ** -0.5
            a = a.masked_fill(~tok[:, None, None, :], float("-inf")).softmax(-1) @ qkv[2]
            a = (a * head[:, :, None, None]).transpose(1, 2).reshape(B, L, D)
            a = F.linear(a, s[p + "attn.proj.weight"], s[p + "attn.proj.bias"])
            x = x + a * tok[:, :, None] * layer[:, 0, None, None]
            m = F.linear(F.gelu(F.linear(ln(x, p + "norm2."), s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"])),
                         s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"])
            x = x + m * tok[:, :, None] * layer[:, 1, None, None]
    return {k: v.float().cpu() for k, v in s.items()}
