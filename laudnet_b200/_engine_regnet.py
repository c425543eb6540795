"""Host-side executor of the LAUD-RegNet-Y forward on top of the C ABI (reference laud_regnet.py:157-217, 281-295,
574-613).  Same conventions as `_engine.ResNetEngine`: prepared fp16 weights + folded BatchNorm, workspaces allocated
once per batch shape, every arithmetic kernel lives in liblaud_b200.so, the forward is CUDA-graph capturable."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import os

import torch
import torch.nn as nn

from . import _lib
from ._engine import BlockOutputs, fold_bn, pack_conv_weight, run_conv
from ._lib import LaudError, check, lib, ptr, stream_ptr


@dataclass
class RegPlan:
    index: int
    stage: int
    w_in: int
    w_b: int
    w_out: int
    stride: int
    H_in: int
    H_out: int
    gw: int
    mode: str
    gran: int
    G: int
    g_spatial: int
    mask_size: int
    se_width: int
    module: nn.Module = None
    wa: torch.Tensor = None
    wb: torch.Tensor = None
    wc: torch.Tensor = None
    wp: Optional[torch.Tensor] = None
    sa: torch.Tensor = None
    ta: torch.Tensor = None
    sb: torch.Tensor = None
    tb: torch.Tensor = None
    sc: torch.Tensor = None
    tc: torch.Tensor = None
    sp: Optional[torch.Tensor] = None
    tp: Optional[torch.Tensor] = None
    se_w1: torch.Tensor = None
    se_b1: torch.Tensor = None
    se_w2: torch.Tensor = None
    se_b2: torch.Tensor = None

    @property
    def use_c(self) -> bool:
        return self.mode in ("channel", "both")

    @property
    def use_s(self) -> bool:
        return self.mode in ("spatial", "both")


class RegNetEngine:
    def __init__(self, model: nn.Module):
        self.model = model
        self.plans: List[RegPlan] = []
        self.prepared_for: Optional[torch.device] = None
        self.impl = _lib.CONV_AUTO
        # "skip": conv c runs on the pixel list of mask_conv3, block output in place (see run_block); "mask": masked-dense
        self.spatial_exec = os.environ.get("LAUD_SPATIAL_EXEC", "mask")
        self._ws: Dict[tuple, dict] = {}

    # ------------------------------------------------------------------ prepare
    def prepare(self) -> None:
        m = self.model
        dev = m.stem[0].weight.device
        if dev.type != "cuda":
            raise LaudError("prepare(): move the model to a CUDA device first - there is no CPU path")
        with torch.no_grad():
            self.stem_w = m.stem[0].weight.detach().to(torch.float16).contiguous()
            self.stem_s, self.stem_t = fold_bn(m.stem[1])
            self.fc_w = m.fc.weight.detach().to(torch.float16).contiguous()
            self.fc_b = m.fc.bias.detach().float().contiguous()
            self.plans = []
            idx = 0
            for s, stage in enumerate(m.trunk_output):
                for blk in stage:
                    f = blk.f
                    w_b = f.a[0].weight.shape[0]
                    p = RegPlan(index=idx, stage=s, w_in=blk.width_in, w_b=w_b, w_out=blk.width_out, stride=blk.stride,
                                H_in=f.output_size * blk.stride, H_out=f.output_size, gw=f.group_width, mode=f.dyn_mode,
                                gran=f.channel_dyn_granularity, G=f.channel_dyn_group,
                                g_spatial=f.spatial_mask_channel_group, mask_size=f.mask_size,
                                se_width=f.se.fc1.weight.shape[0], module=blk)
                    for cdim in (p.w_in, p.w_b, p.w_out):
                        if cdim % 8:
                            raise LaudError(f"channel counts must be multiples of 8 for the fp16 kernels (got {cdim})")
                    if p.gw % 8 or (p.gw not in (8, 16, 24) and f.dyn_mode in ("channel", "both")):
                        raise LaudError(f"grouped 3x3 convolution: group width {p.gw} with dyn_mode '{f.dyn_mode}' is not supported by "
                                        "the CUDA path (multiples of 8; the channel gate of conv b is built for widths 8 / 16 / 24)")
                    p.wa = pack_conv_weight(f.a[0].weight)
                    p.sa, p.ta = fold_bn(f.a[1])
                    wb = f.b[0].weight.detach()                      # [w_b, gw, 3, 3] -> [w_b][tap][gw]
                    p.wb = wb.permute(0, 2, 3, 1).reshape(w_b, 9, p.gw).to(torch.float16).contiguous()
                    p.sb, p.tb = fold_bn(f.b[1])
                    p.wc = pack_conv_weight(f.c[0].weight)
                    p.sc, p.tc = fold_bn(f.c[1])
                    if blk.proj is not None:
                        p.wp = pack_conv_weight(blk.proj[0].weight)
                        p.sp, p.tp = fold_bn(blk.proj[1])
                    p.se_w1 = f.se.fc1.weight.detach().float().reshape(p.se_width, w_b).contiguous()
                    p.se_b1 = f.se.fc1.bias.detach().float().contiguous()
                    p.se_w2 = f.se.fc2.weight.detach().float().reshape(w_b, p.se_width).t().contiguous()   # [S][w_b]: coalesced gate kernel
                    p.se_b2 = f.se.fc2.bias.detach().float().contiguous()
                    self.plans.append(p)
                    idx += 1
        self.stats_consts = self._stats_consts()
        self.prepared_for = dev
        self._ws.clear()

    def _stats_consts(self) -> torch.Tensor:
        rows = []
        for p in self.plans:
            f = p.module.f
            m_chan = m_spat = 0
            if p.use_c:
                mk = f.masker_channel
                if hasattr(mk, "conv_flops"):
                    m_chan = p.w_in * p.H_in * p.H_in + mk.conv_flops
                else:
                    m_chan = mk.conv[0].weight.shape[0] * p.H_in * p.H_in + mk.masker_flops
            S = min(p.mask_size, p.H_in)
            if p.use_s:
                m_spat = p.w_in * S * S + f.masker_spatial.conv_flops_pp * S * S
            rows.append([m_chan, m_spat,
                         f.conv1_flops_per_pixel * p.H_in * p.H_in,
                         f.conv2_flops_per_pixel * p.H_out * p.H_out,
                         f.conv3_flops_per_pixel * p.H_out * p.H_out,
                         (p.module.downsample_flops * p.H_out * p.H_out) if p.module.proj is not None else 0,
                         0, 0, 0, 0,
                         (1 if p.use_c else 0) | (2 if p.use_s else 0) | 4, f.se_flops_per_pixel])
        return torch.tensor(rows, dtype=torch.int64)

    # --------------------------------------------------------------- workspaces
    def _workspace(self, B: int, H: int, W: int, dev) -> dict:
        key = (B, H, W, dev)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        m = self.model
        f16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        C0 = m.stem[0].weight.shape[0]
        act = max([B * (H // 2) * (W // 2) * C0] + [B * p.H_out * p.H_out * p.w_out for p in self.plans])
        a1 = max(B * p.H_in * p.H_in * p.w_b for p in self.plans)
        a2 = max(B * p.H_out * p.H_out * p.w_b for p in self.plans)
        nb = len(self.plans)
        Cmax = max([p.w_in for p in self.plans] + [p.w_b for p in self.plans] + [m.fc.weight.shape[1]])
        Gmax = max([p.G for p in self.plans if p.use_c] or [1])
        g_max = max(p.g_spatial for p in self.plans)
        hw_max = max(p.H_in * p.H_in for p in self.plans)
        ws = dict(
            act=[torch.empty(act, **f16), torch.empty(act, **f16), torch.empty(act, **f16)],
            a1=torch.empty(a1, **f16), a2=torch.empty(a2, **f16),
            partial=torch.empty(B * (_lib.GAP_SPLITS + 1) * Cmax, **f32),
            pooled=torch.empty(B * Cmax, **f32), segate=torch.empty(B * Cmax, **f32),
            cmask=torch.empty((B, Gmax), dtype=torch.uint8, device=dev), cidx=torch.empty((B, Gmax), **i32),
            ccnt=torch.empty((B,), **i32),
            smask=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m3=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m2=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            m1=torch.empty(B * g_max * hw_max, dtype=torch.uint8, device=dev),
            counts=torch.zeros((nb, 4), **i32), stats=torch.empty(nb * 5 + 1, **f32),
            rows2=torch.empty((B * hw_max,), **i32), rcnt=torch.zeros((2,), **i32),
            cws=torch.zeros((max(64, B * hw_max // 2048 + 2),), **i32))
        consts = self.stats_consts.clone()
        for i, p in enumerate(self.plans):
            S = min(p.mask_size, p.H_in)
            consts[i, 6] = B * p.G
            consts[i, 7] = B * p.g_spatial * S * S
            consts[i, 8] = B * p.g_spatial * p.H_out * p.H_out
            consts[i, 9] = B * p.g_spatial * p.H_in * p.H_in
        ws["consts"] = consts.to(dev)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ blocks
    def run_block(self, p: RegPlan, x: torch.Tensor, out: torch.Tensor, idbuf: torch.Tensor, B: int, ws: dict,
                  keep: Optional[BlockOutputs] = None, forced_channel_mask: Optional[torch.Tensor] = None,
                  forced_spatial_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: fp16 [B,H_in,H_in,w_in] -> out: fp16 [B,H_out,H_out,w_out]   (laud_regnet.py:157-217, 281-295)"""
        from ._engine import ResNetEngine
        from .utils import _ChannelGate
        blk, f = p.module, p.module.f
        L = lib()
        st = stream_ptr()
        Hi, Ho = p.H_in, p.H_out
        counts = ws["counts"][p.index]
        gate = None
        m3 = None
        if p.use_c:
            G = p.G
            gate = _ChannelGate(ws["cmask"].view(-1)[:B * G].view(B, G), ws["cidx"].view(-1)[:B * G].view(B, G),
                                ws["ccnt"], None, None)
            if keep is not None:
                gate.logits = torch.empty((B, 2 * G), dtype=torch.float32, device=x.device)
                gate.pooled = torch.empty((B, p.w_in), dtype=torch.float32, device=x.device)
            if forced_channel_mask is not None:
                ResNetEngine._force_channel_gate(gate, forced_channel_mask, counts[0:1])
            else:
                f.masker_channel.gate_nhwc(x[:B * Hi * Hi * p.w_in].view(B, Hi, Hi, p.w_in), counts[0:1], out=gate,
                                           partial_ws=ws["partial"])
        if p.use_s:
            g = p.g_spatial
            S = min(p.mask_size, Hi)
            small = ws["smask"][:B * g * S * S].view(B, g, S, S)
            if forced_spatial_mask is not None:
                small.copy_(forced_spatial_mask.to(torch.uint8))
                counts[1:2].copy_(small.sum().to(torch.int32).view(1))
                slog = None
            else:
                slog = torch.empty((B, 2 * g, S, S), dtype=torch.float32, device=x.device) if keep is not None else None
                wt, wbias = f.masker_spatial._weights()
                check(L.laud_masker_spatial(ptr(x), B, Hi, Hi, p.w_in, ptr(wt), ptr(wbias),
                                            g, S, ptr(slog), ptr(small), ptr(counts[1:2]), st), "laud_masker_spatial")
            m3 = ws["m3"][:B * g * Ho * Ho].view(B, g, Ho, Ho)
            m2 = ws["m2"][:B * g * Ho * Ho].view(B, g, Ho, Ho)
            m1 = ws["m1"][:B * g * Hi * Hi].view(B, g, Hi, Hi)
            check(L.laud_spatial_masks(ptr(small), B, g, S, Ho, p.stride, ptr(m3), ptr(m2), ptr(m1), ptr(counts[2:3]),
                                       ptr(counts[3:4]), st), "laud_spatial_masks")
            if keep is not None:
                keep.spatial_mask_small, keep.spatial_logits = small.clone(), slog
                keep.mask_conv3, keep.mask_conv2, keep.mask_conv1 = m3.clone(), m2.clone(), m1.clone()

        a1, a2 = ws["a1"], ws["a2"]
        tag = f"r{p.stage + 1}"
        # a: 1x1 + BN + ReLU (dense; the channel gate of :183 is applied where conv b reads it)
        run_conv(x, p.wa, a1, B, Hi, Hi, p.w_in, Hi, Hi, p.w_b, 1, 1, 0, ldx=p.w_in, ldy=p.w_b, scale=p.sa, shift=p.ta,
                 relu=_lib.RELU_ALL, impl=self.impl, tag=tag + ".a")
        # b: grouped 3x3 + BN + ReLU (+ channel gate on its input and output)
        cmask = gate.mask if gate is not None else None
        check(L.laud_grouped_conv3x3_forward(ptr(a1), B, Hi, Hi, p.w_b, p.stride, ptr(p.wb), p.gw, ptr(p.sb), ptr(p.tb),
                                             ptr(cmask), p.gran, ptr(a2), st), "laud_grouped_conv3x3_forward")
        # Squeeze-Excitation over the DENSE conv-b output (parity mode)
        check(L.laud_global_avg_pool(ptr(a2), B, Ho * Ho, p.w_b, p.w_b, ptr(ws["partial"]), ptr(ws["pooled"]), st),
              "laud_global_avg_pool")
        check(L.laud_se_gate(ptr(ws["pooled"]), B, p.w_b, ptr(p.se_w1), ptr(p.se_b1), p.se_width, ptr(p.se_w2),
                             ptr(p.se_b2), ptr(cmask), p.gran, ptr(ws["segate"]), st), "laud_se_gate")
        check(L.laud_scale_channels(ptr(a2), B, Ho * Ho, p.w_b, ptr(ws["segate"]), st), "laud_scale_channels")
        spatial_skip = (getattr(self, "spatial_exec", "mask") == "skip" and p.mode == "spatial" and p.g_spatial == 1
                        and "rows2" in ws and self.impl == _lib.CONV_AUTO)
        if spatial_skip:
            # Parity mode of SURVEY 7 H2: the Squeeze-Excitation pools the DENSE conv-b output, so a / b run everywhere;
            # conv c - the one layer the reference's semantics allow to skip - runs on the pixel list of mask_conv3 and
            # the block output is written in place (a gated-off pixel keeps relu(identity), laud_regnet.py:197-198,290-295)
            rc = ws["rcnt"]
            check(L.laud_compact_rows(ptr(m3), B, 1, Ho * Ho, ptr(ws["rows2"]), ptr(rc[1:2]), ptr(ws["cws"]), st),
                  "laud_compact_rows")
            if p.wp is not None:
                run_conv(x, p.wp, out, B, Hi, Hi, p.w_in, Ho, Ho, p.w_out, 1, p.stride, 0, ldx=p.w_in, ldy=p.w_out,
                         scale=p.sp, shift=p.tp, relu=_lib.RELU_WHERE_GATE0, out_mask=m3, mask_groups=1, impl=self.impl,
                         tag=tag + ".proj")
                dst = out
            else:
                dst = x
            run_conv(a2, p.wc, dst, B, Ho, Ho, p.w_b, Ho, Ho, p.w_out, 1, 1, 0, ldx=p.w_b, ldy=p.w_out, scale=p.sc,
                     shift=p.tc, relu=_lib.RELU_ALL, residual=dst, ldr=p.w_out, impl=self.impl, tag=tag + ".c",
                     row_idx=ws["rows2"], row_cnt=rc[1:2])
            if keep is not None:
                keep.out = dst[:B * Ho * Ho * p.w_out].view(B, Ho, Ho, p.w_out).clone()
            return dst
        # identity / projection     :284-290
        if p.wp is not None:
            run_conv(x, p.wp, idbuf, B, Hi, Hi, p.w_in, Ho, Ho, p.w_out, 1, p.stride, 0, ldx=p.w_in, ldy=p.w_out,
                     scale=p.sp, shift=p.tp, relu=_lib.RELU_NONE, impl=self.impl, tag=tag + ".proj")
            res = idbuf
        else:
            res = x
        # c: 1x1 + BN (+ spatial gate) + identity + ReLU     :197-198, :290-295
        run_conv(a2, p.wc, out, B, Ho, Ho, p.w_b, Ho, Ho, p.w_out, 1, 1, 0, ldx=p.w_b, ldy=p.w_out, scale=p.sc, shift=p.tc,
                 relu=_lib.RELU_ALL, residual=res, ldr=p.w_out, out_mask=m3,
                 mask_groups=p.g_spatial if m3 is not None else 1, impl=self.impl, tag=tag + ".c")
        if keep is not None:
            if gate is not None:
                keep.channel_mask, keep.channel_idx, keep.channel_cnt = gate.mask.clone(), gate.idx.clone(), gate.cnt.clone()
                keep.channel_logits = gate.logits
            keep.a1 = a1[:B * Hi * Hi * p.w_b].view(B, Hi, Hi, p.w_b).clone()
            keep.a2 = a2[:B * Ho * Ho * p.w_b].view(B, Ho, Ho, p.w_b).clone()
            keep.out = out[:B * Ho * Ho * p.w_out].view(B, Ho, Ho, p.w_out).clone()
        return out

    # ----------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, keep: Optional[List[BlockOutputs]] = None, forced=None):
        if x.device.type != "cuda":
            raise LaudError("LAD_RegNet.forward: expected a CUDA tensor - there is no CPU path")
        with torch.cuda.device(x.device):        # launches go to the current stream of the INPUT's device
            return self._forward(x, keep, forced)

    def _forward(self, x, keep, forced=None):
        m = self.model
        if self.prepared_for != x.device:
            self.prepare()
        if x.dtype not in (torch.float16, torch.float32):
            raise LaudError(f"LAD_RegNet.forward: unsupported input dtype {x.dtype}")
        B, cin, H, W = x.shape
        if cin != 3 or H != m.input_size or W != m.input_size:
            raise LaudError(f"LAD_RegNet.forward: expected [B,3,{m.input_size},{m.input_size}], got {tuple(x.shape)}")
        xh = x.contiguous() if x.dtype == torch.float16 else x.contiguous().to(torch.float16)
        ws = self._workspace(B, H, W, x.device)
        L = lib()
        st = stream_ptr()
        ws["counts"].zero_()
        C0 = m.stem[0].weight.shape[0]
        bufs = ws["act"]
        cur = 0
        check(L.laud_regnet_stem_forward(ptr(xh), B, H, W, ptr(self.stem_w), C0, ptr(self.stem_s), ptr(self.stem_t),
                                         ptr(bufs[cur]), st), "laud_regnet_stem_forward")
        for p in self.plans:
            nxt, idb = (cur + 1) % 3, (cur + 2) % 3
            ko = None
            if keep is not None:
                ko = BlockOutputs()
                keep.append(ko)
            fc, fs = forced[p.index] if forced is not None else (None, None)
            res_buf = self.run_block(p, bufs[cur], bufs[nxt], bufs[idb], B, ws, ko, fc, fs)
            if res_buf is not bufs[cur]:          # (an in-place spatial skip returns the input buffer)
                cur = nxt
        last = self.plans[-1]
        ncls = m.fc.weight.shape[0]
        logits = torch.empty((B, ncls), dtype=torch.float32, device=x.device)
        check(L.laud_head_forward(ptr(bufs[cur]), B, last.H_out * last.H_out, last.w_out, ptr(self.fc_w), ptr(self.fc_b),
                                  ncls, ptr(ws["partial"]), ptr(logits), st), "laud_head_forward")
        stem_flops = 3 * C0 * (H // 2) * (W // 2) * 9
        stats = torch.empty_like(ws["stats"])
        check(L.laud_forward_stats(ptr(ws["counts"]), ptr(ws["consts"]), len(self.plans), stem_flops, last.w_out,
                                   last.w_out * ncls, ptr(stats), st), "laud_forward_stats")
        return logits, stats

    def split_stats(self, stats: torch.Tensor):
        nb = len(self.plans)
        tab = stats[:nb * 5].view(nb, 5)
        bounds, s0 = [], 0
        for stage in self.model.trunk_output:
            bounds.append((s0, s0 + len(stage)))
            s0 += len(stage)
        col = lambda c: [tab[a:b, c] for a, b in bounds]
        return col(0), col(1), col(2), col(3), tab[:, 4], stats[nb * 5]
