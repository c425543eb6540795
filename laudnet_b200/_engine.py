"""Host-side executor of the LAUD-ResNet forward on top of the C ABI.

Holds the *prepared* model (fp16 K-major conv weights, folded BatchNorm
scale/shift, fp32 masker parameters) and the activation / mask workspaces, and
sequences the kernels of one forward pass on torch's current CUDA stream.
torch is used for device memory and streams only; every arithmetic kernel
launched from here lives in liblaud_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import ConvDesc, LaudError, check, lib, ptr, stream_ptr


def fold_bn(bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm as y = x*scale + shift (fp32)."""
    with torch.no_grad():
        scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
        shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
    return scale, shift


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, kh, kw] fp32 -> fp16 [C_out, kh*kw, C_in] (K-major: in-channel fastest)."""
    with torch.no_grad():
        co, ci, kh, kw = w.shape
        return w.detach().permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(torch.float16).contiguous()


def pack_conv_weight_t(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, kh, kw] fp32 -> fp16 [kh*kw, C_in, C_out] (out-channel fastest): the
    transposed copy the K-row-gather path reads (16-byte gathers of active INPUT channels)."""
    with torch.no_grad():
        co, ci, kh, kw = w.shape
        return w.detach().permute(2, 3, 1, 0).reshape(kh * kw, ci, co).to(torch.float16).contiguous()


def round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def gap_tiles(hw: int) -> int:
    """Partial-sum slots per sample of the fused GAP (laud_conv_desc::gap_tiles): 128-pixel tiles of the flat pixel
    list that can hold pixels of one sample."""
    return (hw - 1) // 128 + 2


class conv_profile:
    """Context manager: bracket every laud_conv_forward launch with CUDA events
    on the launching stream (bench.py's per-kernel timing).  `.total_ms()` after
    a synchronize gives the summed device time of the conv launches."""
    active = None

    def __enter__(self):
        self.records = []
        conv_profile.active = self
        return self

    def __exit__(self, *exc):
        conv_profile.active = None
        return False

    def total_ms(self) -> float:
        return sum(a.elapsed_time(b) for _, a, b in self.records)

    def by_tag(self):
        out = {}
        for tag, a, b in self.records:
            n, t = out.get(tag, (0, 0.0))
            out[tag] = (n + 1, t + a.elapsed_time(b))
        return out


conv_tags: Optional[list] = None      # when a list: every run_conv appends its tag (launch order of laud_conv_profile records)


def run_conv(x, w, y, B, H_in, W_in, C_in, H_out, W_out, C_out, ksize, stride, pad, *,
             ldx=None, ldy=None, scale=None, shift=None, relu=_lib.RELU_NONE, residual=None, ldr=0,
             k_idx=None, k_cnt=None, k_gran=1, n_idx=None, n_cnt=None, n_gran=1,
             pre_bias=None, pre_bias_classes=0, pre_bias_ld=0, out_mask=None, mask_groups=1,
             sample_idx=None, sample_cnt=None, row_idx=None, row_cnt=None, n_pad_align=0,
             impl=_lib.CONV_AUTO, tag="conv", w_t=None, bias_t=None, bias_ld=0, n_mask=None, n_mask_gran=1,
             gap_partial=None, gap_tiles=0, n_expand=0) -> None:
    """Fill a laud_conv_desc and enqueue laud_conv_forward on the current stream."""
    d = ConvDesc()
    d.x, d.ldx = ptr(x), ldx if ldx is not None else x.shape[-1]
    d.w = ptr(w)
    d.y, d.ldy = ptr(y), ldy if ldy is not None else y.shape[-1]
    d.B, d.H_in, d.W_in, d.C_in = B, H_in, W_in, C_in
    d.H_out, d.W_out, d.C_out = H_out, W_out, C_out
    d.ksize, d.stride, d.pad = ksize, stride, pad
    d.scale, d.shift = ptr(scale), ptr(shift)
    d.relu_mode = relu
    d.residual, d.ldr = ptr(residual), ldr
    d.k_idx, d.k_cnt = ptr(k_idx), ptr(k_cnt)
    d.k_ld, d.k_gran = (k_idx.shape[-1] if k_idx is not None else 0), k_gran
    d.n_idx, d.n_cnt = ptr(n_idx), ptr(n_cnt)
    d.n_ld, d.n_gran = (n_idx.shape[-1] if n_idx is not None else 0), n_gran
    d.pre_bias, d.pre_bias_classes, d.pre_bias_ld = ptr(pre_bias), pre_bias_classes, pre_bias_ld
    d.out_mask, d.mask_groups = ptr(out_mask), mask_groups
    d.sample_idx, d.sample_cnt = ptr(sample_idx), ptr(sample_cnt)
    d.row_idx, d.row_cnt = ptr(row_idx), ptr(row_cnt)
    d.n_pad_align = n_pad_align
    d.gap_partial, d.gap_tiles = ptr(gap_partial), gap_tiles
    d.w_t = ptr(w_t)
    d.bias_t, d.bias_ld = ptr(bias_t), bias_ld
    d.n_mask, d.n_mask_gran = ptr(n_mask), n_mask_gran
    d.n_expand = n_expand
    if conv_tags is not None:
        conv_tags.append(tag)
    prof = conv_profile.active
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().laud_conv_forward(C.byref(d), impl, stream_ptr()), "laud_conv_forward")
    if prof is not None:
        e1.record()
        prof.records.append((tag, e0, e1))


@dataclass
class BlockPlan:
    """Static description + prepared parameters of one bottleneck."""
    index: int
    stage: int
    inplanes: int
    width: int
    outplanes: int
    stride: int
    H_in: int
    H_out: int
    mode: str
    gran: int
    G: int
    g_spatial: int
    mask_size: int
    w1: torch.Tensor = None
    w2: torch.Tensor = None
    w3: torch.Tensor = None
    wd: Optional[torch.Tensor] = None
    s1: torch.Tensor = None
    t1: torch.Tensor = None
    s2: torch.Tensor = None
    t2: torch.Tensor = None
    s3: torch.Tensor = None
    t3: torch.Tensor = None
    sd: Optional[torch.Tensor] = None
    td: Optional[torch.Tensor] = None
    w2t: Optional[torch.Tensor] = None      # transposed copies for the K-row-gather path
    w3t: Optional[torch.Tensor] = None
    cw: Optional[torch.Tensor] = None       # H1 constants: [9*width + outplanes, width] weights pre-scaled by relu(shift)
    module: nn.Module = None
    W_in: int = 0                           # feature-map widths (== heights for the classification networks; the detection
    W_out: int = 0                          # backbone adapter runs non-square inputs in channel / layer mode)

    def __post_init__(self):
        self.W_in = self.W_in or self.H_in
        self.W_out = self.W_out or self.H_out

    @property
    def masker_kind(self) -> Optional[str]:
        """'MLP' (pools the block input: its GAP can be fused into the producing conv3) or 'conv_linear' (pools
        relu(bn(conv1x1(x))), utils.py:150-169: needs the activations)."""
        mk = getattr(self.module, "masker_channel", None)
        if mk is None:
            return None
        return "MLP" if hasattr(mk, "gate_from_partials") else "conv_linear"

    @property
    def use_c(self) -> bool:
        return self.mode in ("channel", "both")

    @property
    def krows_ok(self) -> bool:
        """Channel granularities the 16-byte K-row-gather path takes (laud_conv_desc.w_t)."""
        return self.use_c and (self.gran in (2, 4) or self.gran % 8 == 0) and self.width <= 1024

    def pack_channel_mode(self, blk) -> None:
        """Per-model packing for channel skipping: transposed conv2/conv3 weights and the
        relu(shift)-scaled weights of the H1-constant GEMM (include/laud_b200.h)."""
        with torch.no_grad():
            self.w2t = pack_conv_weight_t(blk.conv2.weight)
            self.w3t = pack_conv_weight_t(blk.conv3.weight)
            c1, c2 = torch.relu(self.t1), torch.relu(self.t2)
            # rows ordered (tap, o): T2[b] comes out as [tap][o], the layout laud_conv_desc.bias_t wants
            w2s = (blk.conv2.weight.detach().float() * c1.view(1, -1, 1, 1)).permute(2, 3, 0, 1).reshape(9 * self.width, self.width)
            w3s = blk.conv3.weight.detach().float().view(self.outplanes, self.width) * c2.view(1, -1)
            self.cw = torch.cat([w2s, w3s]).to(torch.float16).contiguous()

    @property
    def use_s(self) -> bool:
        return self.mode in ("spatial", "layer", "both")


class BlockOutputs:
    """Optional per-block device tensors kept for parity tests."""
    __slots__ = ("channel_mask", "channel_idx", "channel_cnt", "channel_logits", "spatial_mask_small",
                 "spatial_logits", "mask_conv3", "mask_conv2", "mask_conv1", "a1", "a2", "out")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


class ResNetEngine:
    """Runs ResNet.forward (reference laud_resnet.py:316-363) as a kernel sequence."""

    def __init__(self, model: "nn.Module"):
        self.model = model
        self.plans: List[BlockPlan] = []
        self.prepared_for: Optional[torch.device] = None
        self.impl = _lib.CONV_AUTO
        # How a channel-gated block executes (both reproduce laud_resnet.py:115-126 exactly):
        #   "sparse": gathered GEMMs over the active channels only + H1 constants (compact a1 / a2);
        #   "dense" : masked-dense - weights shared by all samples, gated channels emitted as their BN constant;
        #   "nskip" : as "dense", but the 3x3 convolutions compute only the sample's ACTIVE output channels: their weight
        #             rows arrive by TMA gather4, the MMAs run over N = active columns, the epilogue expands them to dense
        #             rows with the BN constants of the gated channels (laud_conv_desc.n_expand) - the consumer still reads
        #             ordinary dense activations with shared weights.
        self.channel_exec = os.environ.get("LAUD_CHANNEL_EXEC", "nskip")     # measured (profiles/r02*): 34.6k vs 34.0k img/s
        # How a layer-gated block executes: "skip" = the convolutions run only on the device-side list of ACTIVE samples
        # and the block output is written in place over the block input (skipped samples are untouched: relu(identity)
        # == identity bit-exactly, laud_resnet.py:133-144); "mask" = masked-dense (compute all, zero the gated rows).
        self.layer_exec = os.environ.get("LAUD_LAYER_EXEC", "skip")
        # How a spatially gated block (dyn_mode='spatial', one mask group) executes:
        #   "mask": masked-dense - conv1/conv2/conv3 everywhere, the gate applied in conv3's epilogue (what the reference does);
        #   "skip": the scheme the reference only MODELS (DyNetSimulator multi_cores.py:67-337,470-511): conv1 runs on the
        #           pixels of mask_conv1 (the 3x3-dilated footprint), conv2 / conv3 on the pixels of mask_conv2 / mask_conv3,
        #           taken from device-side pixel lists; the block output is written IN PLACE over the block input, so a
        #           gated-off pixel keeps relu(identity) == identity bit-exactly (laud_resnet.py:133-144).  Lists with fewer
        #           than ~one MMA tile of pixels fall to the CUDA-core kernel (device-side dispatch, laud_conv_forward).
        self.spatial_exec = os.environ.get("LAUD_SPATIAL_EXEC", "mask")
        # conv3 leaves the global-average-pool partial sums of the block output for the next block's channel masker
        self.fuse_gap = os.environ.get("LAUD_NO_GAP_FUSE") is None
        self._ws: Dict[tuple, dict] = {}

    # ------------------------------------------------------------------ prepare
    def prepare(self) -> None:
        m = self.model
        dev = m.conv1.weight.device
        if dev.type != "cuda":
            raise LaudError("prepare(): move the model to a CUDA device first - there is no CPU path")
        self.stem_w = m.conv1.weight.detach().to(torch.float16).contiguous()
        self.stem_s, self.stem_t = fold_bn(m.bn1)
        self.fc_w = m.fc.weight.detach().to(torch.float16).contiguous()
        self.fc_b = m.fc.bias.detach().float().contiguous()
        self.plans = []
        idx = 0
        for s, layer in enumerate((m.layer1, m.layer2, m.layer3, m.layer4)):
            for blk in layer:
                p = BlockPlan(index=idx, stage=s, inplanes=blk.conv1.weight.shape[1], width=blk.conv1.weight.shape[0],
                              outplanes=blk.conv3.weight.shape[0], stride=blk.stride,
                              H_in=blk.output_size * blk.stride, H_out=blk.output_size, mode=blk.dyn_mode,
                              gran=blk.channel_dyn_granularity, G=blk.channel_dyn_group,
                              g_spatial=blk.spatial_mask_channel_group, mask_size=blk.mask_size, module=blk)
                p.W_out = getattr(blk, "output_w", None) or blk.output_size
                p.W_in = p.W_out * blk.stride
                if p.W_out != p.H_out and blk.dyn_mode in ("spatial", "both"):
                    raise LaudError("non-square feature maps are supported in channel / layer mode only (the spatial mask "
                                    "geometry kernels take square masks)")
                if blk.conv2.groups != 1:
                    raise LaudError("grouped conv2 (group_width>1) is not supported by the CUDA path")
                p.w1, p.w2, p.w3 = (pack_conv_weight(c.weight) for c in (blk.conv1, blk.conv2, blk.conv3))
                p.s1, p.t1 = fold_bn(blk.bn1)
                p.s2, p.t2 = fold_bn(blk.bn2)
                p.s3, p.t3 = fold_bn(blk.bn3)
                if blk.downsample is not None:
                    p.wd = pack_conv_weight(blk.downsample[0].weight)
                    p.sd, p.td = fold_bn(blk.downsample[1])
                if p.use_c:
                    p.pack_channel_mode(blk)
                for cdim in (p.inplanes, p.width, p.outplanes):
                    if cdim % 8:
                        raise LaudError(f"channel counts must be multiples of 8 for the fp16 kernels (got {cdim})")
                self.plans.append(p)
                idx += 1
        self.stats_consts = self._stats_consts(dev)
        self.prepared_for = dev
        self._ws.clear()
        for g in getattr(self, "_graphs", []):       # captured graphs hold raw pointers to the tensors just replaced
            gf = g()
            if gf is not None:
                gf.valid = False
        self._graphs = []

    def _stats_consts(self, dev) -> torch.Tensor:
        rows = []
        for p in self.plans:
            blk = p.module
            m_chan = m_spat = 0
            if p.use_c:
                mk = blk.masker_channel
                if hasattr(mk, "conv_flops"):
                    m_chan = p.inplanes * p.H_in * p.W_in + mk.conv_flops
                else:
                    cr = mk.conv[0].weight.shape[0]
                    m_chan = cr * p.H_in * p.W_in + mk.masker_flops
            S = min(p.mask_size, p.H_in)
            if p.use_s:
                m_spat = p.inplanes * S * S + blk.masker_spatial.conv_flops_pp * S * S
            rows.append([m_chan, m_spat,
                         blk.conv1_flops_per_pixel * p.H_in * p.W_in,
                         blk.conv2_flops_per_pixel * p.H_out * p.W_out,
                         blk.conv3_flops_per_pixel * p.H_out * p.W_out,
                         (blk.downsample_flops * p.H_out * p.W_out) if blk.downsample is not None else 0,
                         0, 0, 0, 0,      # denominators depend on the batch: filled per forward
                         (1 if p.use_c else 0) | (2 if p.use_s else 0), 0])
        return torch.tensor(rows, dtype=torch.int64)     # host; a device copy per batch size lives in the workspace

    # --------------------------------------------------------------- workspaces
    def _workspace(self, B: int, H: int, W: int, dev, slot: int = 0) -> dict:
        key = (B, H, W, dev, slot)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        m = self.model
        f16 = dict(dtype=torch.float16, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        C0 = m.conv1.weight.shape[0]
        act = max([B * (H // 4) * (W // 4) * C0] + [B * p.H_out * p.W_out * p.outplanes for p in self.plans])
        a1 = max(B * p.H_in * p.W_in * (p.width + 16) for p in self.plans)
        a2 = max(B * p.H_out * p.W_out * (p.width + 16) for p in self.plans)
        nb = len(self.plans)
        Gmax = max([p.G for p in self.plans if p.use_c] or [1])
        Cmax = max([p.inplanes for p in self.plans] + [m.fc.weight.shape[1]])
        wmax = max(p.width for p in self.plans)
        comax = max(p.outplanes for p in self.plans)
        hw_max = max(p.H_in * p.W_in for p in self.plans)
        g_max = max(p.g_spatial for p in self.plans)
        ws = dict(
            act=[torch.empty(act, **f16), torch.empty(act, **f16), torch.empty(act, **f16)],
            a1=torch.empty(a1, **f16), a2=torch.empty(a2, **f16),
            partial=torch.empty(B * (_lib.GAP_SPLITS + 1) * Cmax, dtype=torch.float32, device=dev),
            cmask=torch.empty((B, Gmax), dtype=torch.uint8, device=dev),
            cidx=torch.empty((B, Gmax), **i32), ccnt=torch.empty((B,), **i32),
            pb2=torch.empty(B * 16 * wmax, dtype=torch.float32, device=dev),
            pb3=torch.empty(B * comax, dtype=torch.float32, device=dev),
            inact=torch.empty(B * wmax, **f16), T=torch.empty(B * (9 * wmax + comax), **f16),
            smask=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m3=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m2=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m1=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            counts=torch.zeros((nb, 4), **i32),
            srows=torch.empty((B,), **i32), scnt=torch.zeros((1,), **i32),
            cws=torch.zeros((max(64, B * hw_max // 2048 + 2),), **i32),
            rows1=torch.empty((B * hw_max,), **i32), rows2=torch.empty((B * hw_max,), **i32),
            rcnt=torch.zeros((2,), **i32),
            lidx=torch.empty((B * g_max,), **i32), lcnt=torch.empty((B,), **i32),
            stats=torch.empty(nb * 5 + 1, dtype=torch.float32, device=dev),
            logits=None,
            gap=torch.empty(max(B * gap_tiles(p.H_out * p.W_out) * p.outplanes for p in self.plans), dtype=torch.float32,
                            device=dev),
        )
        cl = [p for p in self.plans if p.masker_kind == "conv_linear"]
        if cl:      # Masker_channel_conv_linear: reduced-width feature map and its pool
            cr = lambda p: (p.module.masker_channel.conv[0].weight.shape[0] + 7) // 8 * 8
            ws["mkz"] = torch.empty(max(B * p.H_in * p.W_in * cr(p) for p in cl), **f16)
            ws["mkpool"] = torch.empty(max(B * cr(p) for p in cl), dtype=torch.float32, device=dev)
        else:
            ws["mkz"] = ws["mkpool"] = None
        consts = self.stats_consts.clone()
        for i, p in enumerate(self.plans):
            S = min(p.mask_size, p.H_in)
            consts[i, 6] = B * p.G
            consts[i, 7] = B * p.g_spatial * S * S
            consts[i, 8] = B * p.g_spatial * p.H_out * p.W_out
            consts[i, 9] = B * p.g_spatial * p.H_in * p.W_in
        ws["consts"] = consts.to(dev)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ blocks
    def run_block(self, p: BlockPlan, x: torch.Tensor, out: torch.Tensor, idbuf: torch.Tensor, B: int, ws: dict,
                  keep: Optional[BlockOutputs] = None, forced_channel_mask: Optional[torch.Tensor] = None,
                  forced_spatial_mask: Optional[torch.Tensor] = None, gap_in: bool = False, gap_out: bool = False,
                  noise=None, tau: float = 1.0) -> None:
        """x: fp16 [B,H_in,H_in,inplanes] -> out: fp16 [B,H_out,H_out,outplanes].
        noise = (channel Gumbel sample [B,2G] | None, spatial [B,2g,S,S] | None): the gates take the reference's TRAINING
        branch (hard Gumbel-softmax at temperature tau, utils.py:56-58,123-125) with the given samples; BatchNorm stays
        in eval mode (the mmdet backbones' norm_eval configuration).
        gap_in: ws["gap"] holds the fused-GAP partial sums of x (left by the previous block's conv3): the channel
        masker decides from them instead of pooling x.  gap_out: conv3 leaves the partial sums of `out` there."""
        blk = p.module
        L = lib()
        st = stream_ptr()
        Hi, Ho, Wi, Wo = p.H_in, p.H_out, p.W_in, p.W_out
        counts = ws["counts"][p.index]
        gate = None
        m3 = None
        fast_layer = False
        wp = p.width + 16                       # channel pitch of the compact intermediates
        use_wt = False
        dense_gate = False
        T, Tn = None, 0
        if p.use_c:
            from .utils import _ChannelGate
            G = p.G
            gate = _ChannelGate(ws["cmask"].view(-1)[:B * G].view(B, G), ws["cidx"].view(-1)[:B * G].view(B, G),
                                ws["ccnt"], None, None)
            nz_c = noise[0] if noise is not None else None
            if keep is not None or nz_c is not None:
                gate.logits = torch.empty((B, 2 * G), dtype=torch.float32, device=x.device)
                gate.pooled = torch.empty((B, x.shape[-1]), dtype=torch.float32, device=x.device)
            ctot = counts[0:1] if nz_c is None else None         # (the Gumbel decision below does the counting)
            if forced_channel_mask is not None:
                self._force_channel_gate(gate, forced_channel_mask, counts[0:1])
            elif gap_in:
                blk.masker_channel.gate_from_partials(ws["gap"], B, Hi * Wi, p.inplanes, gap_tiles(Hi * Wi), ctot, gate)
            elif p.masker_kind == "conv_linear":
                # conv 1x1 + BN + ReLU at full resolution, then pool (utils.py:150-169): pre-packed weights, workspaces
                blk.masker_channel.gate_nhwc(x[:B * Hi * Wi * p.inplanes].view(B, Hi, Wi, p.inplanes), ctot,
                                             out=gate, partial_ws=ws["partial"], impl=self.impl, z_ws=ws["mkz"],
                                             pooled_ws=ws["mkpool"])
            else:
                blk.masker_channel.gate_nhwc(x[:B * Hi * Wi * p.inplanes].view(B, Hi, Wi, p.inplanes), ctot,
                                             out=gate, partial_ws=ws["partial"])
            if nz_c is not None and forced_channel_mask is None:
                from .utils import gate_from_logits
                gate_from_logits(gate.logits, nz_c, tau, G, 1, gate.mask, gate.idx, gate.cnt, counts[0:1])
            dense_gate = self.channel_exec in ("dense", "nskip")
            if not dense_gate:
                # H1 constants: 0/1 indicator of the masked channels -> one dense GEMM -> fold taps into border classes
                use_wt = p.krows_ok and self.impl in (_lib.CONV_AUTO, _lib.CONV_UMMA)
                inact = ws["inact"][:B * p.width].view(B, p.width)
                Tn = 9 * p.width + p.outplanes
                T = ws["T"][:B * Tn].view(B, Tn)
                check(L.laud_gate_inactive(ptr(gate.mask), B, G, p.gran, ptr(inact), st), "laud_gate_inactive")
                run_conv(inact, p.cw, T, 1, B, 1, p.width, B, 1, Tn, 1, 1, 0, ldx=p.width, ldy=Tn, impl=self.impl,
                         tag=f"s{p.stage + 1}.h1gemm")
                if not use_wt:      # fallback layouts: fold the taps into per-border-class pre-bias tables
                    check(L.laud_channel_consts_fold(ptr(T), B, p.width, p.outplanes, ptr(gate.idx), ptr(gate.cnt), G,
                                                     p.gran, 1, ptr(ws["pb2"]), ptr(ws["pb3"]), st),
                          "laud_channel_consts_fold")
        if p.use_s:
            g = p.g_spatial
            S = min(p.mask_size, Hi)
            small = ws["smask"][:B * g * S * S].view(B, g, S, S)
            if forced_spatial_mask is not None:
                small.copy_(forced_spatial_mask.to(torch.uint8))
                counts[1:2].copy_(small.sum().to(torch.int32).view(1))
                slog = None
            else:
                nz_s = noise[1] if noise is not None else None
                slog = (torch.empty((B, 2 * g, S, S), dtype=torch.float32, device=x.device)
                        if (keep is not None or nz_s is not None) else None)
                stot = counts[1:2] if nz_s is None else None
                wt, wbias = blk.masker_spatial._weights()
                if S == 1 and "lidx" in ws:
                    # one gate per sample (layer skip): global pool -> 2g-row linear -> keep>=drop is exactly the
                    # one-layer channel masker; its fused one-CTA-per-sample kernel pools at HBM speed
                    check(L.laud_masker_channel_mlp(ptr(x), B, Hi * Wi, p.inplanes, 1, ptr(wt),
                                                    ptr(wbias), 0, None, None, g,
                                                    ptr(ws["partial"]), None, ptr(slog), ptr(small), ptr(ws["lidx"]),
                                                    ptr(ws["lcnt"]), ptr(stot), st), "laud_masker_channel_mlp")
                else:
                    check(L.laud_masker_spatial(ptr(x), B, Hi, Wi, p.inplanes, ptr(wt),
                                                ptr(wbias), g, S, ptr(slog), ptr(small),
                                                ptr(stot), st), "laud_masker_spatial")
                if nz_s is not None:
                    from .utils import gate_from_logits
                    gate_from_logits(slog, nz_s, tau, g, S * S, small, total=counts[1:2])
            m3 = ws["m3"][:B * g * Ho * Wo].view(B, g, Ho, Wo)
            m2 = ws["m2"][:B * g * Ho * Wo].view(B, g, Ho, Wo)
            m1 = ws["m1"][:B * g * Hi * Wi].view(B, g, Hi, Wi)
            fast_layer = (p.mode == "layer" and self.layer_exec == "skip" and g == 1 and S == 1 and "srows" in ws
                          and keep is None)
            if fast_layer:
                # one launch: active-sample work list + the counts of the broadcast / dilated masks
                check(L.laud_layer_gate_lists(ptr(small), B, Ho * Wo, Hi * Wi, ptr(counts), ptr(ws["srows"]),
                                              ptr(ws["scnt"]), st), "laud_layer_gate_lists")
                if p.wd is not None:     # the downsample branch gates its ReLU per pixel
                    if Ho == Wo:
                        check(L.laud_resize_mask_nearest(ptr(small), B, g, S, Ho, ptr(m3), st), "laud_resize_mask_nearest")
                    else:
                        check(L.laud_broadcast_gate(ptr(small), B, g, Ho * Wo, ptr(m3), None, st), "laud_broadcast_gate")
            elif Ho != Wo:
                # non-square map (detection adapter, layer mode: S == 1): the gate broadcast to the three mask sizes; a
                # per-sample gate dilates to itself (ExpandMask of a constant map)
                if S != 1:
                    raise LaudError("non-square feature maps need a per-sample gate (mask_size 1)")
                check(L.laud_broadcast_gate(ptr(small), B, g, Ho * Wo, ptr(m3), None, st), "laud_broadcast_gate")
                check(L.laud_broadcast_gate(ptr(small), B, g, Ho * Wo, ptr(m2), ptr(counts[2:3]), st), "laud_broadcast_gate")
                check(L.laud_broadcast_gate(ptr(small), B, g, Hi * Wi, ptr(m1), ptr(counts[3:4]), st), "laud_broadcast_gate")
            else:
                check(L.laud_spatial_masks(ptr(small), B, g, S, Ho, p.stride, ptr(m3), ptr(m2), ptr(m1), ptr(counts[2:3]),
                                           ptr(counts[3:4]), st), "laud_spatial_masks")
            if keep is not None:
                keep.spatial_mask_small, keep.spatial_logits = small.clone(), slog
                keep.mask_conv3, keep.mask_conv2, keep.mask_conv1 = m3.clone(), m2.clone(), m1.clone()

        # layer skip: ordered list of the active samples; convolutions take it as their work list
        # (one gate per SAMPLE: with spatial_mask_channel_group > 1 the gate is per (sample, channel group) and the
        #  block runs masked-dense with the grouped out_mask, as laud_resnet.py:133 / utils.py:27-33 do)
        skip = (p.mode == "layer" and self.layer_exec == "skip" and gate is None and "srows" in ws and p.g_spatial == 1)
        sl = {}
        if skip:
            if not fast_layer:
                check(L.laud_compact_rows(ptr(small), B, 1, 1, ptr(ws["srows"]), ptr(ws["scnt"]), ptr(ws["cws"]), st),
                      "laud_compact_rows")
            sl = dict(sample_idx=ws["srows"], sample_cnt=ws["scnt"])
        a1, a2 = ws["a1"], ws["a2"]
        if (p.mode == "spatial" and self.spatial_exec == "skip" and p.g_spatial == 1 and gate is None and "rows1" in ws
                and self.impl == _lib.CONV_AUTO):
            return self._run_block_spatial_skip(p, x, out, B, ws, m3, m2, m1, keep)
        sparse_gate = gate is not None and not dense_gate
        ck = dict(k_idx=gate.idx, k_cnt=gate.cnt, k_gran=p.gran) if sparse_gate else {}
        cn = dict(n_idx=gate.idx, n_cnt=gate.cnt, n_gran=p.gran, n_pad_align=16) if sparse_gate else {}
        nm = dict(n_mask=gate.mask, n_mask_gran=p.gran) if dense_gate else {}
        ld12 = wp if sparse_gate else p.width
        # conv1 1x1 (+ mask) + bn1 + relu      laud_resnet.py:115-118
        run_conv(x, p.w1, a1, B, Hi, Wi, p.inplanes, Hi, Wi, p.width, 1, 1, 0, ldx=p.inplanes, ldy=ld12,
                 scale=p.s1, shift=p.t1, relu=_lib.RELU_ALL, impl=self.impl, tag=f"s{p.stage + 1}.conv1", **cn, **nm, **sl)
        # conv2 3x3/stride (+ mask) + bn2 + relu     laud_resnet.py:123-126
        if dense_gate and not sl and self.uses_nskip(p):
            # channel skipping with a dense result: active weight rows by TMA gather4, N = active columns, expanded rows
            run_conv(a1, p.w2, a2, B, Hi, Wi, p.width, Ho, Wo, p.width, 3, 1, 1, ldx=p.width, ldy=p.width,
                     scale=p.s2, shift=p.t2, relu=_lib.RELU_ALL, impl=self.impl, tag=f"s{p.stage + 1}.conv2",
                     n_idx=gate.idx, n_cnt=gate.cnt, n_gran=p.gran, n_pad_align=16, n_expand=1)
        else:
          run_conv(a1, p.w2, a2, B, Hi, Wi, p.width, Ho, Wo, p.width, 3, p.stride, 1, ldx=ld12, ldy=ld12,
                 scale=p.s2, shift=p.t2, relu=_lib.RELU_ALL, impl=self.impl, tag=f"s{p.stage + 1}.conv2",
                 pre_bias=ws["pb2"] if (sparse_gate and not use_wt) else None,
                 pre_bias_classes=16 if (sparse_gate and not use_wt) else 0,
                 pre_bias_ld=p.width if sparse_gate else 0, w_t=p.w2t if use_wt else None,
                 bias_t=T if use_wt else None, bias_ld=Tn if use_wt else 0, **ck, **cn, **nm, **sl)
        # identity branch      laud_resnet.py:138-141
        if skip:
            if p.wd is not None:
                # every sample gets its downsampled identity straight into `out`; a skipped sample's output is
                # relu(identity) (ReLU applied here, where the gate is 0), an active one is finished by conv3 in place
                run_conv(x, p.wd, out, B, Hi, Wi, p.inplanes, Ho, Wo, p.outplanes, 1, p.stride, 0, ldx=p.inplanes,
                         ldy=p.outplanes, scale=p.sd, shift=p.td, relu=_lib.RELU_WHERE_GATE0, out_mask=m3, mask_groups=1,
                         impl=self.impl, tag=f"s{p.stage + 1}.down")
                dst = out
            else:
                dst = x                            # in place: skipped samples keep relu(x) == x
            run_conv(a2, p.w3, dst, B, Ho, Wo, p.width, Ho, Wo, p.outplanes, 1, 1, 0, ldx=ld12, ldy=p.outplanes,
                     scale=p.s3, shift=p.t3, relu=_lib.RELU_ALL, residual=dst, ldr=p.outplanes, impl=self.impl,
                     tag=f"s{p.stage + 1}.conv3", **sl)
            out = dst
        else:
            if p.wd is not None:
                run_conv(x, p.wd, idbuf, B, Hi, Wi, p.inplanes, Ho, Wo, p.outplanes, 1, p.stride, 0, ldx=p.inplanes,
                         ldy=p.outplanes, scale=p.sd, shift=p.td, relu=_lib.RELU_NONE, impl=self.impl,
                         tag=f"s{p.stage + 1}.down")
                res = idbuf
            else:
                res = x
            # conv3 1x1 + bn3 (+ spatial mask) + identity + relu     laud_resnet.py:131-144
            run_conv(a2, p.w3, out, B, Ho, Wo, p.width, Ho, Wo, p.outplanes, 1, 1, 0, ldx=ld12, ldy=p.outplanes,
                     scale=p.s3, shift=p.t3, relu=_lib.RELU_ALL, residual=res, ldr=p.outplanes, impl=self.impl,
                     tag=f"s{p.stage + 1}.conv3", pre_bias=ws["pb3"] if (sparse_gate and not use_wt) else None,
                     pre_bias_classes=1 if (sparse_gate and not use_wt) else 0,
                     bias_t=T.view(-1)[9 * p.width:] if use_wt else None, bias_ld=Tn if use_wt else 0,
                     pre_bias_ld=p.outplanes if sparse_gate else 0, out_mask=m3, mask_groups=p.g_spatial if m3 is not None else 1,
                     w_t=p.w3t if (sparse_gate and use_wt) else None,
                     gap_partial=ws["gap"] if gap_out else None, gap_tiles=gap_tiles(Ho * Wo) if gap_out else 0, **ck)
        if keep is not None:
            if gate is not None:
                keep.channel_mask, keep.channel_idx, keep.channel_cnt = gate.mask.clone(), gate.idx.clone(), gate.cnt.clone()
                keep.channel_logits = gate.logits
            keep.a1 = a1[:B * Hi * Wi * ld12].view(B, Hi, Wi, ld12).clone()
            keep.a2 = a2[:B * Ho * Wo * ld12].view(B, Ho, Wo, ld12).clone()
            keep.out = out[:B * Ho * Wo * p.outplanes].view(B, Ho, Wo, p.outplanes).clone()
        return out          # the buffer that holds the block output (the input buffer for an in-place layer skip)

    def uses_nskip(self, p: BlockPlan) -> bool:
        """True if block p's 3x3 convolution computes only the sample's ACTIVE output channels (laud_conv_desc.n_expand).
        Where it pays, measured on B200 (profiles/r02*): width >= 256 with at least two m-tiles per sample (14x14 maps:
        1.36 ms vs 1.55 ms masked-dense over the 22 stage-3 layers of ResNet-101); narrower layers are not MMA-bound and
        at 7x7 the per-sample weight-row gather (4.7 MB of weights per sample) costs more than the skipped MMAs save."""
        return (self.channel_exec == "nskip" and p.use_c and p.stride == 1 and p.W_out + 2 <= 128 and p.gran % 2 == 0
                and p.width >= getattr(self, "nskip_min_width", 256)
                and p.H_out * p.W_out >= getattr(self, "nskip_min_pixels", 128)
                and self.impl in (_lib.CONV_AUTO, _lib.CONV_UMMA))

    def _run_block_spatial_skip(self, p: BlockPlan, x, out, B, ws, m3, m2, m1, keep):
        """Spatial skipping executed (see spatial_exec): pixel lists of mask_conv1 / mask_conv2 (= mask_conv3 for one mask
        group), conv1 -> conv2 -> conv3 over the listed pixels only, output in place."""
        L = lib()
        st = stream_ptr()
        Hi, Ho, Wi, Wo = p.H_in, p.H_out, p.W_in, p.W_out
        rows1, rows2, rc = ws["rows1"], ws["rows2"], ws["rcnt"]
        check(L.laud_compact_rows(ptr(m1), B, 1, Hi * Wi, ptr(rows1), ptr(rc[0:1]), ptr(ws["cws"]), st), "laud_compact_rows")
        check(L.laud_compact_rows(ptr(m2), B, 1, Ho * Wo, ptr(rows2), ptr(rc[1:2]), ptr(ws["cws"]), st), "laud_compact_rows")
        a1, a2 = ws["a1"], ws["a2"]
        tag = f"s{p.stage + 1}"
        # conv1 on the dilated footprint: everything conv2 will read (ExpandMask(stride, 1) covers its 3x3 windows)
        run_conv(x, p.w1, a1, B, Hi, Wi, p.inplanes, Hi, Wi, p.width, 1, 1, 0, ldx=p.inplanes, ldy=p.width,
                 scale=p.s1, shift=p.t1, relu=_lib.RELU_ALL, impl=self.impl, tag=tag + ".conv1", row_idx=rows1, row_cnt=rc[0:1])
        run_conv(a1, p.w2, a2, B, Hi, Wi, p.width, Ho, Wo, p.width, 3, p.stride, 1, ldx=p.width, ldy=p.width,
                 scale=p.s2, shift=p.t2, relu=_lib.RELU_ALL, impl=self.impl, tag=tag + ".conv2", row_idx=rows2, row_cnt=rc[1:2])
        if p.wd is not None:
            # every pixel gets its downsampled identity straight into `out`; a gated-off pixel's output is relu(identity)
            # (ReLU applied here, where the gate is 0), an active one is finished by conv3 in place
            run_conv(x, p.wd, out, B, Hi, Wi, p.inplanes, Ho, Wo, p.outplanes, 1, p.stride, 0, ldx=p.inplanes,
                     ldy=p.outplanes, scale=p.sd, shift=p.td, relu=_lib.RELU_WHERE_GATE0, out_mask=m3, mask_groups=1,
                     impl=self.impl, tag=tag + ".down")
            dst = out
        else:
            dst = x                                # in place: gated-off pixels keep relu(x) == x
        run_conv(a2, p.w3, dst, B, Ho, Wo, p.width, Ho, Wo, p.outplanes, 1, 1, 0, ldx=p.width, ldy=p.outplanes,
                 scale=p.s3, shift=p.t3, relu=_lib.RELU_ALL, residual=dst, ldr=p.outplanes, impl=self.impl,
                 tag=tag + ".conv3", row_idx=rows2, row_cnt=rc[1:2])
        if keep is not None:
            keep.out = dst[:B * Ho * Wo * p.outplanes].view(B, Ho, Wo, p.outplanes).clone()
        return dst

    def _gap_fusable(self, p: BlockPlan) -> bool:
        """True if block p's conv3 can leave the GAP partial sums of its output for the NEXT block's channel masker:
        the next block pools its input for a channel gate, and this conv3 is a flat GEMM (1x1, nothing per sample:
        masked-dense channel execution, no spatial mask, no layer skip) on the tcgen05 path."""
        if not self.fuse_gap:
            return False
        if p.index + 1 < len(self.plans):        # (the last block's pool feeds the head)
            nxt = self.plans[p.index + 1]
            if not nxt.use_c or nxt.masker_kind != "MLP":
                return False             # conv_linear pools relu(bn(conv(x))), not x: it must read the activations
        if os.environ.get("LAUD_CONV_V3") or os.environ.get("LAUD_NO_FLAT") or os.environ.get("LAUD_NO_DMA"):
            return False                 # A/B switches that take conv3 off the flat slab path of the TMA-staged kernel
        if p.use_s or (p.use_c and self.channel_exec not in ("dense", "nskip")) or self.impl not in (_lib.CONV_AUTO, _lib.CONV_UMMA):
            return False
        return p.outplanes % 64 == 0 and p.H_out * p.W_out >= 43

    @staticmethod
    def _force_channel_gate(gate, mask: torch.Tensor, total: torch.Tensor) -> None:
        """Teacher forcing (tests): install a given 0/1 mask [B,G] as the gate."""
        m = mask.to(torch.uint8)
        gate.mask.copy_(m)
        B, G = m.shape
        order = torch.argsort(1 - m.to(torch.int32), dim=1, stable=True).to(torch.int32)   # actives first, ascending
        gate.idx.copy_(order)
        gate.cnt.copy_(m.sum(dim=1).to(torch.int32))
        total.copy_(m.sum().to(torch.int32).view(1))

    # ----------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, keep: Optional[List[BlockOutputs]] = None, slot: int = 0,
                logits_out: Optional[torch.Tensor] = None, want_stats: bool = True, forced=None, gumbel_noise=None,
                temperature: float = 1.0, stage_outputs: Optional[list] = None):
        """forced (tests): per block a pair (channel mask [B,G] | None, spatial mask [B,g,S,S] | None) installed
        instead of the block's own gating decision - the teacher-forced network forward.
        gumbel_noise: per block a pair (channel sample [B,2G] | None, spatial sample [B,2g,S,S] | None): the gates take
        the training branch (hard Gumbel-softmax at `temperature`) with these samples, BatchNorm in eval mode."""
        if x.device.type != "cuda":
            raise LaudError("ResNet.forward: expected a CUDA tensor - there is no CPU path")
        with torch.cuda.device(x.device):        # launches go to the current stream of the INPUT's device
            return self._forward(x, keep, slot, logits_out, want_stats, forced, gumbel_noise, temperature, stage_outputs)

    def _forward(self, x, keep, slot, logits_out, want_stats, forced=None, gumbel_noise=None, temperature=1.0,
                 stage_outputs=None):
        """stage_outputs: a list -> BACKBONE mode (the mmdet adapter): the output of the last block of every stage is
        appended as fp32 NCHW, no classifier head is run and the returned flops exclude it (lad_mmdet_resnet.py:680-751)."""
        m = self.model
        if self.prepared_for != x.device:
            self.prepare()
        if x.dtype not in (torch.float16, torch.float32):
            raise LaudError(f"ResNet.forward: unsupported input dtype {x.dtype}")
        B, cin, H, W = x.shape
        w_in = getattr(m, "input_w", None) or m.input_size       # (set by the detection adapter for non-square inputs)
        if cin != 3 or H != m.input_size or W != w_in:
            raise LaudError(f"ResNet.forward: expected [B,3,{m.input_size},{w_in}], got {tuple(x.shape)}")
        xh = x.contiguous() if x.dtype == torch.float16 else x.contiguous().to(torch.float16)
        ws = self._workspace(B, H, W, x.device, slot)
        L = lib()
        st = stream_ptr()
        ws["counts"].zero_()
        C0 = m.conv1.weight.shape[0]
        bufs = ws["act"]
        cur = 0
        check(L.laud_stem_forward(ptr(xh), B, H, W, ptr(self.stem_w), C0, ptr(self.stem_s), ptr(self.stem_t),
                                  ptr(bufs[cur]), st), "laud_stem_forward")
        nvtx = bool(os.environ.get("LAUD_NVTX"))
        gap_in = False
        for p in self.plans:
            nxt = (cur + 1) % 3
            idb = (cur + 2) % 3
            ko = None
            if keep is not None:
                ko = BlockOutputs()
                keep.append(ko)
            if nvtx:
                torch.cuda.nvtx.range_push(f"blk{p.index}")
            gap_out = self._gap_fusable(p)
            fc, fs = forced[p.index] if forced is not None else (None, None)
            res_buf = self.run_block(p, bufs[cur], bufs[nxt], bufs[idb], B, ws, ko, gap_in=gap_in, gap_out=gap_out,
                                     forced_channel_mask=fc, forced_spatial_mask=fs,
                                     noise=gumbel_noise[p.index] if gumbel_noise is not None else None, tau=temperature)
            gap_in = gap_out
            if nvtx:
                torch.cuda.nvtx.range_pop()
            if res_buf is not bufs[cur]:
                cur = nxt
            if stage_outputs is not None and (p.index + 1 == len(self.plans) or self.plans[p.index + 1].stage != p.stage):
                from .utils import to_nchw_f32
                stage_outputs.append(to_nchw_f32(bufs[cur][:B * p.H_out * p.W_out * p.outplanes].view(B, p.H_out, p.W_out, p.outplanes)))
        if stage_outputs is not None:
            stats = torch.empty_like(ws["stats"])
            self._launch_stats(ws["counts"], ws["consts"], H, W, stats, head=False)
            return None, stats
        last = self.plans[-1]
        feat = last.outplanes
        ncls = m.fc.weight.shape[0]
        logits = logits_out if logits_out is not None else torch.empty((B, ncls), dtype=torch.float32, device=x.device)
        if gap_in:      # the last conv3 left the pool of its output
            check(L.laud_head_forward_from_partials(ptr(ws["gap"]), B, last.H_out * last.W_out, feat,
                                                    gap_tiles(last.H_out * last.W_out), ptr(self.fc_w), ptr(self.fc_b), ncls,
                                                    ptr(ws["partial"]), ptr(logits), st), "laud_head_forward_from_partials")
        else:
            check(L.laud_head_forward(ptr(bufs[cur]), B, last.H_out * last.W_out, feat, ptr(self.fc_w), ptr(self.fc_b),
                                      ncls, ptr(ws["partial"]), ptr(logits), st), "laud_head_forward")
        if not want_stats:
            return logits, None
        stats = torch.empty_like(ws["stats"])
        self._launch_stats(ws["counts"], ws["consts"], H, W, stats)
        return logits, stats

    def _launch_stats(self, counts, consts, H, W, stats, head: bool = True) -> None:
        m = self.model
        C0 = m.conv1.weight.shape[0]
        feat, ncls = self.plans[-1].outplanes, m.fc.weight.shape[0]
        stem_flops = 3 * C0 * (H // 2) * (W // 2) * 49 + C0 * (H // 4) * (W // 4) * 9
        check(lib().laud_forward_stats(ptr(counts), ptr(consts), len(self.plans), stem_flops, feat if head else 0,
                                       feat * ncls if head else 0, ptr(stats), stream_ptr()), "laud_forward_stats")

    def forward_split(self, x: torch.Tensor, splits: int):
        """The forward as `splits` independent chains over contiguous slices of the batch, each on its own stream
        (samples are independent in eval mode).  Inside a CUDA graph the chains become parallel branches: while one
        chain's kernel drains its last CTAs, the other chain's kernels fill the idle SMs - the per-sample work items
        of a 256-image batch otherwise leave ~14 % of the 148 SMs idle in every second wave.  Logits land in one
        tensor; the statistics are computed from the summed counts with the whole batch's denominators, so they equal
        the unsplit forward's."""
        with torch.cuda.device(x.device):
            return self._forward_split(x, splits)

    def _forward_split(self, x: torch.Tensor, splits: int):
        if self.prepared_for != x.device:
            self.prepare()
        B, _, H, W = x.shape
        dev = x.device
        xh = x.contiguous() if x.dtype == torch.float16 else x.contiguous().to(torch.float16)
        ncls = self.model.fc.weight.shape[0]
        logits = torch.empty((B, ncls), dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream()
        if not hasattr(self, "_streams") or len(self._streams) < splits:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(splits)]
        start = torch.cuda.Event()
        start.record(main)
        bounds = [(i * B // splits, (i + 1) * B // splits) for i in range(splits)]
        done = []
        for i, (lo, hi) in enumerate(bounds):
            st = self._streams[i]
            st.wait_event(start)
            with torch.cuda.stream(st):
                self.forward(xh[lo:hi], slot=i + 1, logits_out=logits[lo:hi], want_stats=False)
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        for ev in done:
            main.wait_event(ev)
        counts = None
        for i, (lo, hi) in enumerate(bounds):
            c = self._workspace(hi - lo, H, W, dev, i + 1)["counts"]
            counts = c.clone() if counts is None else counts + c
        consts = self.stats_consts.clone()
        for i, p in enumerate(self.plans):
            S = min(p.mask_size, p.H_in)
            consts[i, 6] = B * p.G
            consts[i, 7] = B * p.g_spatial * S * S
            consts[i, 8] = B * p.g_spatial * p.H_out * p.W_out
            consts[i, 9] = B * p.g_spatial * p.H_in * p.W_in
        key = ("split_consts", B, dev)
        if key not in self._ws:
            self._ws[key] = consts.to(dev)
        stats = torch.empty(len(self.plans) * 5 + 1, dtype=torch.float32, device=dev)
        self._launch_stats(counts, self._ws[key], H, W, stats)
        return logits, stats

    # ------------------------------------------------------------- CUDA graph
    def capture(self, x_example: torch.Tensor) -> "GraphedForward":
        """Capture one forward (all kernels of liblaud_b200.so for this batch
        shape) into a CUDA graph: the launch-bound host loop (~10 launches per
        block) becomes a single cudaGraphLaunch."""
        return GraphedForward(self, x_example)

    def split_stats(self, stats: torch.Tensor):
        """stats [n_blocks*5+1] -> the reference's (rho3[4], rho2[4], rho1[4], rho_c[4], flops_perc, flops)."""
        nb = len(self.plans)
        tab = stats[:nb * 5].view(nb, 5)
        bounds, s0 = [], 0
        for layer in (self.model.layer1, self.model.layer2, self.model.layer3, self.model.layer4):
            bounds.append((s0, s0 + len(layer)))
            s0 += len(layer)
        col = lambda c: [tab[a:b, c] for a, b in bounds]
        return col(0), col(1), col(2), col(3), tab[:, 4], stats[nb * 5]


class GraphedForward:
    """A captured forward.  `run(x)` copies x into the static input (device to
    device, or host to device when x is pinned host memory), replays the graph
    and returns the static logits / stats tensors (overwritten by the next run)."""

    def __init__(self, engine: ResNetEngine, x_example: torch.Tensor, splits: Optional[int] = None, post=None,
                 profile_convs: bool = False):
        """post: optional callable(logits) captured at the end of the graph (e.g. the logits all-gather of a sharded
        batch); its return value is kept as `self.post_out`.  profile_convs: capture the graph with every convolution
        kernel bracketed by event-record nodes (laud_conv_profile): `conv_times()` after a replay returns the device
        time of each conv launch INSIDE that graph execution, `conv_tags` their layer tags."""
        if x_example.device.type != "cuda":
            raise LaudError("capture(): expected a CUDA example input")
        if splits is None:       # measured on B200 at batch 256: 2 chains +2.5 %, 3 chains +0.5 %, 4 chains -2 %
            splits = int(os.environ.get("LAUD_SPLITS", "2" if x_example.shape[0] >= 128 else "1"))
        if not hasattr(engine, "forward_split") or x_example.shape[0] < 2 * splits:
            splits = 1
        self.splits = splits
        fwd = (lambda xx: engine.forward_split(xx, splits)) if splits > 1 else engine.forward
        self.engine = engine
        self.valid = True
        self.device = x_example.device
        if engine.prepared_for != x_example.device:
            engine.prepare()
        self.static_x = x_example.detach().to(torch.float16).contiguous().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):              # warm-up: workspaces, func attributes, lazy prepare
            for _ in range(2):
                fwd(self.static_x)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.conv_tags = None
        self.post_out = None
        global conv_tags
        if profile_convs:
            lib().laud_conv_profile(1)
            conv_tags = []
        # programmatic dependent launch of the conv kernels inside a SINGLE-chain graph (small batches): measured +4 % at batch 8,
        # +0.8 % at batch 256; with two chains it costs 2.5 % (csrc/conv_tma.cu) and stays off
        self.pdl = splits == 1 and os.environ.get("LAUD_NO_PDL") is None
        try:
            lib().laud_conv_set_pdl(1 if self.pdl else 0)
            with torch.cuda.graph(self.graph):
                self.logits, self.stats = fwd(self.static_x)
                if post is not None:
                    self.post_out = post(self.logits)
        finally:
            lib().laud_conv_set_pdl(0)
            if profile_convs:
                lib().laud_conv_profile(2)           # stop collecting, keep the (graph-owned) records
                self.conv_tags, conv_tags = conv_tags, None
        self.launches = _lib.launch_count() - n0     # kernels of ours inside one replay
        # the graph replays raw device pointers: keep the prepared tensors and workspaces alive with it, and let the
        # engine mark it stale when prepare() replaces them (load_state_dict, .to(), in-place edits + prepare())
        self._keepalive = (list(engine.plans), dict(engine._ws), getattr(engine, "stem_w", None), getattr(engine, "fc_w", None))
        import weakref
        if not hasattr(engine, "_graphs"):
            engine._graphs = []
        engine._graphs.append(weakref.ref(self))

    def replay(self):
        if not self.valid or self.engine.prepared_for != self.device:
            raise LaudError("GraphedForward: the model was re-prepared (load_state_dict / .to() / prepare()) after this "
                            "graph was captured; capture() it again")
        self.graph.replay()
        return self.logits, self.stats

    def run(self, x: torch.Tensor):
        self.static_x.copy_(x, non_blocking=True)
        return self.replay()

    def conv_times(self):
        """After a replay and a synchronize of a `profile_convs` graph: ms of every conv launch, in `conv_tags` order."""
        if self.conv_tags is None:
            raise LaudError("GraphedForward.conv_times(): capture with profile_convs=True")
        n = len(self.conv_tags)
        buf = (C.c_float * n)()
        got = lib().laud_conv_profile_read_all(buf, n)
        if got != n:
            raise LaudError(f"conv profile holds {got} records, the graph has {n} conv launches (another profile was started)")
        return list(buf)
