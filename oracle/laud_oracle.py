"""CPU oracle for the LAUDNet dynamic-operator hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-fp32 (CPU) *restatement* of the reference's eval-mode
forward for the masker -> mask-conditioned bottleneck path.  It is written
functionally (state_dict in, tensors out) and shares no code with the
reference; every function cites the reference file:line it restates (paths
relative to the reference checkout, `imagenet_classification/models/...`).

Who may use it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` - as the checker or the timed CPU arm,
never as the product path.  Nothing under `laudnet_b200/` imports this module.

Parity pin: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, generated in the build container by `tests/golden/make_golden.py`
(which imports /root/reference) and committed under `tests/golden/*.npz`.
`tests/test_oracle_golden.py` replays them on every CPU test run.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5


# --------------------------------------------------------------------------
# configuration (mirrors the ctor kwargs of laud_resnet.py:169-181)
# --------------------------------------------------------------------------
@dataclass
class ResNetCfg:
    layers: Sequence[int] = (3, 4, 23, 3)
    width_mult: float = 1.0
    input_size: int = 224
    num_classes: int = 1000
    dyn_mode: Sequence[str] = ("channel",) * 4
    channel_dyn_granularity: Sequence[int] = (2, 2, 2, 2)
    spatial_mask_channel_group: Sequence[int] = (1, 1, 1, 1)
    mask_spatial_granularity: Sequence[int] = (4, 4, 2, 1)
    channel_masker: Sequence[str] = ("MLP",) * 4
    channel_masker_layers: Sequence[int] = (2, 2, 2, 2)
    reduction_ratio: Sequence[int] = (16, 16, 16, 16)

    def kwargs(self) -> dict:
        """The kwargs the reference / drop-in constructors take."""
        return dict(
            input_size=self.input_size,
            width_mult=self.width_mult,
            num_classes=self.num_classes,
            dyn_mode=list(self.dyn_mode),
            channel_dyn_granularity=list(self.channel_dyn_granularity),
            spatial_mask_channel_group=list(self.spatial_mask_channel_group),
            mask_spatial_granularity=list(self.mask_spatial_granularity),
            channel_masker=list(self.channel_masker),
            channel_masker_layers=list(self.channel_masker_layers),
            reduction_ratio=list(self.reduction_ratio),
        )


@dataclass
class BlockGeom:
    """Static geometry of one bottleneck (laud_resnet.py:28-86)."""
    prefix: str
    inplanes: int
    width: int
    outplanes: int
    stride: int
    output_size: int
    mask_size: int
    dyn_mode: str
    groups_channel: int          # G = width // granularity
    groups_spatial: int          # g = spatial_mask_channel_group
    masker_kind: str
    masker_layers: int
    has_downsample: bool


def resnet_geometry(cfg: ResNetCfg) -> List[BlockGeom]:
    """Block list in execution order (laud_resnet.py:208-250, 269-314)."""
    geoms: List[BlockGeom] = []
    inplanes = int(64 * cfg.width_mult)
    for s in range(4):
        planes = int(64 * (2 ** s) * cfg.width_mult)
        out_size = cfg.input_size // (4 * 2 ** s)
        mode = cfg.dyn_mode[s]
        for i in range(cfg.layers[s]):
            stride = 2 if (s > 0 and i == 0) else 1
            has_ds = i == 0 and (stride != 1 or inplanes != planes * 4)
            msize = 1 if mode == "layer" else out_size // cfg.mask_spatial_granularity[s]
            geoms.append(BlockGeom(
                prefix=f"layer{s + 1}.{i}.", inplanes=inplanes, width=planes,
                outplanes=planes * 4, stride=stride, output_size=out_size,
                mask_size=msize, dyn_mode=mode,
                groups_channel=planes // cfg.channel_dyn_granularity[s],
                groups_spatial=cfg.spatial_mask_channel_group[s],
                masker_kind=cfg.channel_masker[s],
                masker_layers=cfg.channel_masker_layers[s],
                has_downsample=has_ds))
            inplanes = planes * 4
    return geoms


# --------------------------------------------------------------------------
# L2 operators (models/utils.py)
# --------------------------------------------------------------------------
def two_way_decision(logits: Tensor) -> Tensor:
    """Eval-mode gate: first half of the channel axis = keep logits, second
    half = drop logits, ties keep (models/utils.py:55-60, 122-127)."""
    half = logits.shape[1] // 2
    return (logits[:, :half] >= logits[:, half:]).to(torch.float32)


def gumbel_two_way_decision(logits: Tensor, noise: Tensor, tau: float) -> Tensor:
    """Training-mode gate (models/utils.py:56-58, 123-125, 161-163):
        F.gumbel_softmax(logits.view(b, 2, ...), dim=1, tau=temperature, hard=True)[:, 0]
    with the Gumbel(0,1) sample `noise` (same shape as `logits`) GIVEN instead of drawn - torch computes
    y_soft = softmax((logits + g) / tau) over the pair and returns one_hot(argmax) (+ y_soft - y_soft.detach(), which is
    the one-hot value in the forward pass up to one rounding).  Returns the 0/1 keep mask."""
    b = logits.shape[0]
    half = logits.shape[1] // 2
    lg = torch.stack([logits[:, :half], logits[:, half:]], dim=1)
    g = torch.stack([noise[:, :half], noise[:, half:]], dim=1)
    y_soft = ((lg + g) / tau).softmax(dim=1)
    index = y_soft.max(dim=1)[1]                     # ties -> index 0 = keep
    return (index == 0).to(torch.float32)


def _decide(logits: Tensor, noise: Optional[Tensor], tau: float) -> Tensor:
    return two_way_decision(logits) if noise is None else gumbel_two_way_decision(logits, noise, tau)


def masker_channel_mlp(x: Tensor, sd: Dict[str, Tensor], prefix: str, layers: int,
                       noise: Optional[Tensor] = None, tau: float = 1.0) -> Tuple[Tensor, Tensor, int, Tensor]:
    """models/utils.py:113-131.  Returns (mask[B,G], sparsity, flops, logits).  noise: training-mode Gumbel sample
    [B, 2G] (None = the eval branch)."""
    b, c, h, w = x.shape
    pooled = x.mean(dim=(2, 3))                                   # :116 GAP
    if layers == 2:
        w1, b1 = sd[prefix + "conv.0.weight"], sd[prefix + "conv.0.bias"]
        w2, b2 = sd[prefix + "conv.2.weight"], sd[prefix + "conv.2.bias"]
        hidden = torch.relu(F.linear(pooled, w1, b1))
        logits = F.linear(hidden, w2, b2)
        mlp_flops = c * w1.shape[0] + w1.shape[0] * w2.shape[0]    # :105
    else:
        w1, b1 = sd[prefix + "conv.weight"], sd[prefix + "conv.bias"]
        logits = F.linear(pooled, w1, b1)
        mlp_flops = c * w1.shape[0]
    mask = _decide(logits, noise, tau)
    return mask, mask.mean(), c * h * w + mlp_flops, logits


def masker_channel_conv_linear(x: Tensor, sd: Dict[str, Tensor], prefix: str,
                               noise: Optional[Tensor] = None, tau: float = 1.0) -> Tuple[Tensor, Tensor, int, Tensor]:
    """models/utils.py:150-169: 1x1 conv -> BN -> ReLU -> GAP -> Linear."""
    wc = sd[prefix + "conv.0.weight"]
    z = F.conv2d(x, wc)
    z = F.batch_norm(z, sd[prefix + "conv.1.running_mean"], sd[prefix + "conv.1.running_var"],
                     sd[prefix + "conv.1.weight"], sd[prefix + "conv.1.bias"], False, 0.0, BN_EPS)
    z = torch.relu(z)
    _, cr, h, w = z.shape
    pooled = z.mean(dim=(2, 3))
    wl, bl = sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]
    logits = F.linear(pooled, wl, bl)
    mask = _decide(logits, noise, tau)
    cin = x.shape[1]
    flops = cr * h * w + cin * cr + cr * wl.shape[0]              # :148,153,157
    return mask, mask.mean(), flops, logits


def masker_spatial(x: Tensor, weight: Tensor, bias: Tensor, mask_size: int,
                   noise: Optional[Tensor] = None, tau: float = 1.0) -> Tuple[Tensor, Tensor, int, Tensor]:
    """models/utils.py:47-65.  Returns (mask[B,g,S,S], sparsity, flops, logits)."""
    q = F.adaptive_avg_pool2d(x, mask_size) if mask_size < x.shape[2] else x   # :48
    flops = q.shape[1] * q.shape[2] * q.shape[3]
    logits = F.conv2d(q, weight, bias)                                          # :51
    per_pixel = weight.shape[0] * weight.shape[1] + weight.shape[1]             # :41
    flops += per_pixel * logits.shape[2] * logits.shape[3]
    mask = _decide(logits, noise, tau)
    return mask, mask.mean(), flops, logits


def expand_mask(mask: Tensor, stride: int, padding: int) -> Tensor:
    """models/utils.py:74-89 restated without convolutions.

    The reference zero-inserts by `stride` with a one-hot transposed conv and
    then box-sums over a (2p+1)^2 window AND over all mask groups with an
    all-ones [g,g,k,k] kernel, thresholding at 0.5.  For a 0/1 mask that is
    'any active cell in the window, in any group', broadcast to every group.
    """
    m = mask.to(torch.float32)
    b, g, h, w = m.shape
    if stride > 1:
        up = torch.zeros(b, g, h * stride, w * stride, dtype=m.dtype)
        up[:, :, ::stride, ::stride] = m
        m = up
    anyg = m.amax(dim=1, keepdim=True)
    k = 2 * padding + 1
    dil = F.max_pool2d(anyg, kernel_size=k, stride=1, padding=padding) if k > 1 else anyg
    return (dil > 0.5).expand(b, g, dil.shape[2], dil.shape[3])


def apply_channel_mask(x: Tensor, mask: Tensor) -> Tensor:
    """models/utils.py:18-25: group j gates the consecutive channels
    [j*c/G, (j+1)*c/G)."""
    b, c = x.shape[:2]
    g = mask.shape[1]
    per_channel = mask.repeat_interleave(c // g, dim=1)
    return x * per_channel.view(b, c, 1, 1)


def apply_spatial_mask(x: Tensor, mask: Tensor) -> Tensor:
    """models/utils.py:27-33."""
    c, g = x.shape[1], mask.shape[1]
    if g > 1 and g != c:
        mask = mask.repeat_interleave(c // g, dim=1)
    return x * mask


def nearest_resize(mask: Tensor, size) -> Tensor:
    """F.interpolate(mode='nearest') as used at laud_resnet.py:106:
    src index = floor(dst * in / out).  `size` is an int (square maps) or (h, w): the detection backbone resizes to the
    actual feature size (mmdet/models/backbones/lad_mmdet_resnet.py:274)."""
    h, w = (size, size) if isinstance(size, int) else size
    iy = torch.div(torch.arange(h) * mask.shape[-2], h, rounding_mode="floor")
    ix = torch.div(torch.arange(w) * mask.shape[-1], w, rounding_mode="floor")
    return mask[:, :, iy][:, :, :, ix]


def _bn(z: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    return F.batch_norm(z, sd[p + "running_mean"], sd[p + "running_var"],
                        sd[p + "weight"], sd[p + "bias"], False, 0.0, BN_EPS)


# --------------------------------------------------------------------------
# L3: bottleneck + network (models/laud_resnet.py)
# --------------------------------------------------------------------------
@dataclass
class BlockTrace:
    """Everything a parity test may want to look at for one block."""
    x: Optional[Tensor] = None
    channel_mask: Optional[Tensor] = None
    channel_logits: Optional[Tensor] = None
    spatial_mask_small: Optional[Tensor] = None
    spatial_logits: Optional[Tensor] = None
    mask_conv3: Optional[Tensor] = None
    mask_conv2: Optional[Tensor] = None
    mask_conv1: Optional[Tensor] = None
    a1: Optional[Tensor] = None
    a2: Optional[Tensor] = None
    out: Optional[Tensor] = None
    stats: Dict[str, float] = field(default_factory=dict)


def bottleneck_forward(x: Tensor, sd: Dict[str, Tensor], g: BlockGeom,
                       trace: Optional[BlockTrace] = None,
                       forced_channel_mask: Optional[Tensor] = None,
                       forced_spatial_mask: Optional[Tensor] = None,
                       noise: Optional[Tuple[Optional[Tensor], Optional[Tensor]]] = None, tau: float = 1.0):
    """laud_resnet.py:88-165, eval-mode BatchNorm.  noise = (channel [B,2G] | None, spatial [B,2g,S,S] | None): the
    maskers take their TRAINING branch (hard Gumbel-softmax at temperature tau) with these Gumbel samples - the
    configuration of the mmdet backbones, which train the gates with frozen BN (lad_mmdet_resnet.py norm_eval).

    Returns (out, rho3, rho2, rho1, rho_c, sparse_flops, dense_flops) with the
    densities as 0-dim tensors like the reference.  `forced_*` feed a given
    (e.g. device-computed) small mask instead of the oracle's own decision, for
    teacher-forced comparisons.
    """
    p = g.prefix
    one = torch.tensor(1.0)
    mode = g.dyn_mode
    use_c = mode in ("channel", "both")
    use_s = mode in ("spatial", "layer", "both")
    cm = sm3 = sm2 = sm1 = None
    rho_c = rho1 = rho2 = rho3 = one
    cflops = sflops = 0
    if use_c:                                                       # :94,:102
        if g.masker_kind == "MLP":
            cm, rho_c, cflops, clog = masker_channel_mlp(x, sd, p + "masker_channel.", g.masker_layers,
                                                         noise[0] if noise else None, tau)
        else:
            cm, rho_c, cflops, clog = masker_channel_conv_linear(x, sd, p + "masker_channel.",
                                                                 noise[0] if noise else None, tau)
        if forced_channel_mask is not None:
            cm = forced_channel_mask.to(torch.float32)
            rho_c = cm.mean()
        if trace is not None:
            trace.channel_mask, trace.channel_logits = cm, clog
    if use_s:                                                       # :98,:103
        small, rho3, sflops, slog = masker_spatial(
            x, sd[p + "masker_spatial.conv.weight"], sd[p + "masker_spatial.conv.bias"], g.mask_size,
            noise[1] if noise else None, tau)
        if forced_spatial_mask is not None:
            small = forced_spatial_mask.to(torch.float32)
            rho3 = small.mean()
        sm3 = nearest_resize(small, (x.shape[2] // g.stride, x.shape[3] // g.stride))   # :106 (= g.output_size for square
        #                                                                                  inputs; lad_mmdet_resnet.py:274)
        sm2 = expand_mask(sm3, 1, 0)                                # :107
        rho2 = sm2.float().mean()
        sm1 = expand_mask(sm2, g.stride, 1)                         # :109
        rho1 = sm1.float().mean()
        if trace is not None:
            trace.spatial_mask_small, trace.spatial_logits = small, slog
            trace.mask_conv3, trace.mask_conv2, trace.mask_conv1 = sm3, sm2, sm1

    sparse = cflops + sflops
    dense = cflops + sflops

    z = F.conv2d(x, sd[p + "conv1.weight"])                         # :115
    if use_c:
        z = apply_channel_mask(z, cm)                               # :116 (mask BEFORE bn)
    a1 = torch.relu(_bn(z, sd, p + "bn1."))
    hw_in = a1.shape[2] * a1.shape[3]
    c1 = g.inplanes * g.width
    dense = dense + c1 * hw_in
    sparse = sparse + c1 * hw_in * rho_c * rho1                     # :121

    z = F.conv2d(a1, sd[p + "conv2.weight"], stride=g.stride, padding=1)
    if use_c:
        z = apply_channel_mask(z, cm)                               # :124
    a2 = torch.relu(_bn(z, sd, p + "bn2."))
    hw = a2.shape[2] * a2.shape[3]
    c2 = g.width * g.width * 9
    dense = dense + c2 * hw
    sparse = sparse + c2 * hw * rho_c ** 2 * rho2                   # :129

    y = _bn(F.conv2d(a2, sd[p + "conv3.weight"]), sd, p + "bn3.")   # :131-132
    if use_s:
        y = apply_spatial_mask(y, sm3)                              # :133 (mask AFTER bn)
    c3 = g.width * g.outplanes
    dense = dense + c3 * hw
    sparse = sparse + c3 * hw * rho_c * rho3                        # :136

    identity = x
    if g.has_downsample:                                            # :138-141
        identity = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=g.stride), sd, p + "downsample.1.")
        ds = g.inplanes * g.outplanes * hw
        dense = dense + ds
        sparse = sparse + ds
    out = torch.relu(y + identity)                                  # :143-144
    if trace is not None:
        trace.x, trace.a1, trace.a2, trace.out = x, a1, a2, out
    return out, rho3, rho2, rho1, rho_c, sparse, dense


def stem_forward(x: Tensor, sd: Dict[str, Tensor]) -> Tuple[Tensor, int]:
    """laud_resnet.py:317-324: conv7x7/2 -> bn -> relu -> maxpool3x3/2 and the
    python-int flop counter that goes with them."""
    cin = x.shape[1]
    z = torch.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1."))
    flops = cin * z.shape[1] * z.shape[2] * z.shape[3] * 49
    z = F.max_pool2d(z, kernel_size=3, stride=2, padding=1)
    flops += z.shape[1] * z.shape[2] * z.shape[3] * 9
    return z, flops


def resnet_forward(sd: Dict[str, Tensor], cfg: ResNetCfg, x: Tensor,
                   traces: Optional[List[BlockTrace]] = None, noise: Optional[List] = None, tau: float = 1.0):
    """laud_resnet.py:316-363.  Returns the reference's 7-tuple:
    (logits, rho3[4], rho2[4], rho1[4], rho_c[4], flops_perc[n_blocks], flops).
    noise: per block (channel Gumbel sample | None, spatial Gumbel sample | None) -> training-branch gates at
    temperature tau with eval-mode BatchNorm (see bottleneck_forward)."""
    feat, flops = stem_forward(x, sd)
    per_stage = {k: [[] for _ in range(4)] for k in ("r3", "r2", "r1", "rc")}
    perc: List[Tensor] = []
    for bi, g in enumerate(resnet_geometry(cfg)):
        tr = BlockTrace() if traces is not None else None
        feat, r3, r2, r1, rc, sparse, dense = bottleneck_forward(feat, sd, g, tr, noise=noise[bi] if noise else None, tau=tau)
        s = int(g.prefix[5]) - 1
        for key, val in (("r3", r3), ("r2", r2), ("r1", r1), ("rc", rc)):
            per_stage[key][s].append(val.reshape(1))
        flops = flops + sparse                                      # :146
        perc.append((sparse / dense).reshape(1))                    # :147
        if traces is not None:
            traces.append(tr)
    pooled = feat.mean(dim=(2, 3))                                  # :349
    flops = flops + pooled.shape[1]
    logits = F.linear(pooled, sd["fc.weight"], sd["fc.bias"])       # :355
    flops = flops + pooled.shape[1] * logits.shape[1]
    cat = lambda key: [torch.cat(v) for v in per_stage[key]]
    flops_t = flops if torch.is_tensor(flops) else torch.tensor(float(flops))
    return logits, cat("r3"), cat("r2"), cat("r1"), cat("rc"), torch.cat(perc), flops_t


def dense_flops(cfg: ResNetCfg) -> int:
    """MACs of the same network with every density = 1 (static ResNet)."""
    s = cfg.input_size
    total = 3 * int(64 * cfg.width_mult) * (s // 2) ** 2 * 49 + int(64 * cfg.width_mult) * (s // 4) ** 2 * 9
    for g in resnet_geometry(cfg):
        hin = g.output_size * g.stride
        total += g.inplanes * g.width * hin * hin
        total += 9 * g.width * g.width * g.output_size ** 2
        total += g.width * g.outplanes * g.output_size ** 2
        if g.has_downsample:
            total += g.inplanes * g.outplanes * g.output_size ** 2
    feat = int(512 * cfg.width_mult) * 4
    return total + feat * (s // 32) ** 2 + feat * cfg.num_classes


# ==========================================================================
# LAUD-RegNet-Y (models/laud_regnet.py) - same operators on the RegNet trunk
# ==========================================================================
@dataclass
class RegNetCfg:
    """Constructor arguments of LAD_RegNet (laud_regnet.py:468-488) with the per-stage block parameters spelled out
    (the reference derives them with BlockParams.from_init_params, :374-446; `laudnet_b200.laud_regnet.stage_params`
    restates that derivation and tests compare it with the values stored in the golden fixtures)."""
    widths: Sequence[int] = (64, 144, 320, 784)            # RegNetY-800MF
    depths: Sequence[int] = (1, 3, 8, 2)
    group_widths: Sequence[int] = (16, 16, 16, 16)
    strides: Sequence[int] = (2, 2, 2, 2)
    bottleneck_multiplier: float = 1.0
    se_ratio: float = 0.25
    stem_width: int = 32
    input_size: int = 224
    num_classes: int = 1000
    dyn_mode: Sequence[str] = ("spatial",) * 4
    channel_dyn_granularity: Sequence[int] = (1, 1, 1, 1)
    spatial_mask_channel_group: Sequence[int] = (1, 1, 1, 1)
    mask_spatial_granularity: Sequence[int] = (4, 4, 2, 1)
    channel_masker: Sequence[str] = ("MLP",) * 4
    channel_masker_layers: Sequence[int] = (2, 2, 2, 2)
    reduction_ratio: Sequence[int] = (16, 16, 16, 16)

    def kwargs(self) -> dict:
        return dict(
            input_size=self.input_size, num_classes=self.num_classes, stem_width=self.stem_width,
            dyn_mode=list(self.dyn_mode), channel_dyn_granularity=list(self.channel_dyn_granularity),
            spatial_mask_channel_group=list(self.spatial_mask_channel_group),
            mask_spatial_granularity=list(self.mask_spatial_granularity),
            channel_masker=list(self.channel_masker), channel_masker_layers=list(self.channel_masker_layers),
            reduction_ratio=list(self.reduction_ratio))


@dataclass
class RegBlockGeom:
    """Static geometry of one ResBottleneckBlock (laud_regnet.py:74-155, 221-279)."""
    prefix: str                  # "trunk_output.block{s}.block{s}-{i}."
    w_in: int
    w_b: int
    w_out: int
    conv_groups: int             # groups of the 3x3 conv = w_b // group_width
    stride: int
    output_size: int
    mask_size: int
    dyn_mode: str
    groups_channel: int
    groups_spatial: int
    masker_kind: str
    masker_layers: int
    has_proj: bool
    se_width: int


def regnet_geometry(cfg: RegNetCfg) -> List[RegBlockGeom]:
    """Blocks in execution order (laud_regnet.py:521-561 builds the stages, :326-346 the blocks of a stage)."""
    geoms: List[RegBlockGeom] = []
    w_prev = cfg.stem_width
    for s in range(len(cfg.widths)):
        w_out, gw = cfg.widths[s], cfg.group_widths[s]
        out_size = cfg.input_size // (2 ** (s + 2))
        for i in range(cfg.depths[s]):
            w_in = w_prev if i == 0 else w_out
            stride = cfg.strides[s] if i == 0 else 1
            w_b = int(round(w_out * cfg.bottleneck_multiplier))
            geoms.append(RegBlockGeom(
                prefix=f"trunk_output.block{s + 1}.block{s + 1}-{i}.", w_in=w_in, w_b=w_b, w_out=w_out,
                conv_groups=w_b // gw, stride=stride, output_size=out_size,
                mask_size=out_size // cfg.mask_spatial_granularity[s], dyn_mode=cfg.dyn_mode[s],
                groups_channel=w_b // cfg.channel_dyn_granularity[s],
                groups_spatial=cfg.spatial_mask_channel_group[s], masker_kind=cfg.channel_masker[s],
                masker_layers=cfg.channel_masker_layers[s], has_proj=(w_in != w_out) or stride != 1,
                se_width=int(round(cfg.se_ratio * w_in))))
        w_prev = w_out
    return geoms


def squeeze_excitation(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """torchvision.ops.misc.SqueezeExcitation as used at laud_regnet.py:128-132,194:
    x * sigmoid(fc2(relu(fc1(avgpool(x))))), fc1/fc2 are 1x1 convolutions with bias."""
    s = x.mean(dim=(2, 3), keepdim=True)
    s = torch.relu(F.conv2d(s, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[p + "fc2.weight"], sd[p + "fc2.bias"]))
    return x * s


def regnet_block_forward(x: Tensor, sd: Dict[str, Tensor], g: RegBlockGeom, trace: Optional[BlockTrace] = None,
                         forced_channel_mask: Optional[Tensor] = None, forced_spatial_mask: Optional[Tensor] = None,
                         noise: Optional[Tuple[Optional[Tensor], Optional[Tensor]]] = None, tau: float = 1.0):
    """BottleneckTransform.forward + ResBottleneckBlock.forward (laud_regnet.py:157-217, 281-295), eval mode.
    Returns (out, rho3, rho2, rho1, rho_c, sparse_flops, dense_flops, se_flops, sparse_flops_of_the_transform,
    projection_flops) - the last two because the reference adds them to `flops` separately."""
    p = g.prefix + "f."
    one = torch.tensor(1.0)
    use_c = g.dyn_mode in ("channel", "both")
    use_s = g.dyn_mode in ("spatial", "both")
    cm = sm3 = None
    rho_c = rho1 = rho2 = rho3 = one
    cflops = sflops = 0
    if use_c:                                                       # :161,:169
        if g.masker_kind == "MLP":
            cm, rho_c, cflops, clog = masker_channel_mlp(x, sd, p + "masker_channel.", g.masker_layers,
                                                         noise[0] if noise else None, tau)
        else:
            cm, rho_c, cflops, clog = masker_channel_conv_linear(x, sd, p + "masker_channel.",
                                                                 noise[0] if noise else None, tau)
        if forced_channel_mask is not None:
            cm = forced_channel_mask.to(torch.float32)
            rho_c = cm.mean()
        if trace is not None:
            trace.channel_mask, trace.channel_logits = cm, clog
    if use_s:                                                       # :165,:170,:172-177
        small, rho3, sflops, slog = masker_spatial(
            x, sd[p + "masker_spatial.conv.weight"], sd[p + "masker_spatial.conv.bias"], g.mask_size,
            noise[1] if noise else None, tau)
        if forced_spatial_mask is not None:
            small = forced_spatial_mask.to(torch.float32)
            rho3 = small.mean()
        sm3 = nearest_resize(small, g.output_size)
        sm2 = expand_mask(sm3, 1, 0)
        rho2 = sm2.float().mean()
        sm1 = expand_mask(sm2, g.stride, 1)
        rho1 = sm1.float().mean()
        if trace is not None:
            trace.spatial_mask_small, trace.spatial_logits = small, slog
            trace.mask_conv3, trace.mask_conv2, trace.mask_conv1 = sm3, sm2, sm1
    sparse = cflops + sflops
    dense = cflops + sflops

    a1 = torch.relu(_bn(F.conv2d(x, sd[p + "a.0.weight"]), sd, p + "a.1."))          # :182
    if use_c:
        a1 = apply_channel_mask(a1, cm)                                               # :183 (mask AFTER bn+relu)
    hw_in = a1.shape[2] * a1.shape[3]
    c1 = g.w_in * g.w_b
    dense = dense + c1 * hw_in
    sparse = sparse + c1 * hw_in * rho_c * rho1                                       # :186

    a2 = torch.relu(_bn(F.conv2d(a1, sd[p + "b.0.weight"], stride=g.stride, padding=1, groups=g.conv_groups),
                        sd, p + "b.1."))                                              # :188
    if use_c:
        a2 = apply_channel_mask(a2, cm)                                               # :189
    hw = a2.shape[2] * a2.shape[3]
    c2 = g.w_b * g.w_b * 9 // g.conv_groups
    dense = dense + c2 * hw
    sparse = sparse + c2 * hw * rho_c ** 2 * rho2                                     # :192

    a2s = squeeze_excitation(a2, sd, p + "se.")                                       # :194 (pools the DENSE a2)
    se_flops = g.w_b * g.se_width * 2                                                 # :151,:195

    y = _bn(F.conv2d(a2s, sd[p + "c.0.weight"]), sd, p + "c.1.")                      # :197
    if use_s:
        y = apply_spatial_mask(y, sm3)                                                # :198
    c3 = g.w_b * g.w_out
    dense = dense + c3 * hw
    sparse = sparse + c3 * hw * rho_c * rho3                                          # :201

    sparse_t = sparse                                                                 # what the transform returns
    ds = 0
    if g.has_proj:                                                                    # :284-288
        ident = _bn(F.conv2d(x, sd[g.prefix + "proj.0.weight"], stride=g.stride), sd, g.prefix + "proj.1.")
        ds = g.w_in * g.w_out * hw
        sparse = sparse + ds
        dense = dense + ds
    else:
        ident = x
    out = torch.relu(ident + y)                                                       # :295
    if trace is not None:
        trace.x, trace.a1, trace.a2, trace.out = x, a1, a2s, out
    return out, rho3, rho2, rho1, rho_c, sparse, dense, se_flops, sparse_t, ds


def regnet_stem_forward(x: Tensor, sd: Dict[str, Tensor]) -> Tuple[Tensor, int]:
    """SimpleStemIN: conv3x3/2 -> BN -> ReLU (laud_regnet.py:59-71, 574-577)."""
    z = torch.relu(_bn(F.conv2d(x, sd["stem.0.weight"], stride=2, padding=1), sd, "stem.1."))
    return z, x.shape[1] * z.shape[1] * z.shape[2] * z.shape[3] * 9


def regnet_forward(sd: Dict[str, Tensor], cfg: RegNetCfg, x: Tensor, traces: Optional[List[BlockTrace]] = None):
    """LAD_RegNet.forward (laud_regnet.py:574-613): the reference's 7-tuple."""
    feat, flops = regnet_stem_forward(x, sd)
    per_stage = {k: [[] for _ in range(4)] for k in ("r3", "r2", "r1", "rc")}
    perc: List[Tensor] = []
    for g in regnet_geometry(cfg):
        tr = BlockTrace() if traces is not None else None
        feat, r3, r2, r1, rc, sparse, dense, se_flops, sparse_t, ds = regnet_block_forward(feat, sd, g, tr)
        s = int(g.prefix.split(".")[1][5:]) - 1
        for key, val in (("r3", r3), ("r2", r2), ("r1", r1), ("rc", rc)):
            per_stage[key][s].append(val.reshape(1))
        flops = flops + se_flops                                                      # :195 (before the sparse flops)
        flops = flops + sparse_t                                                      # :203
        if ds:
            flops = flops + ds                                                        # :288
        perc.append((sparse / dense).reshape(1))
        if traces is not None:
            traces.append(tr)
    c = feat.shape[1]
    pooled = feat.mean(dim=(2, 3))
    flops = flops + c                                                                 # :596
    logits = F.linear(pooled, sd["fc.weight"], sd["fc.bias"])
    flops = flops + c * logits.shape[1]                                               # :602
    cat = lambda lst: [torch.cat(v) for v in lst]
    return (logits, cat(per_stage["r3"]), cat(per_stage["r2"]), cat(per_stage["r1"]), cat(per_stage["rc"]),
            torch.cat(perc), flops)
