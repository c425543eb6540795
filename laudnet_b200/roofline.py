"""Algorithmic work of one LAUD-ResNet forward, from the MEASURED densities.

Definitions follow SURVEY.md section 8(d) and are restated in DESIGN.md:

* FLOPs per image  F = 2 x the reference's own counter (laud_resnet.py:112-147,
  321-356):  stem + sum_blocks( M_b + c1*rho_c*rho_1 + c2*rho_c^2*rho_2
  + c3*rho_c*rho_3 + ds ) + head, with rho = this batch's measured means.
* Bytes per image: fp16 activations, every tensor touched the minimum number
  of times WITHOUT cross-layer fusion, per block
      2 B x [ X (masker read) + X*rho_1 (conv1 read) + 2*I1*rho_c*rho_1
              + 2*I2*rho_c*rho_2 + ID + O ]
  X = C_in*H_in^2, I1 = w*H_in^2, I2 = w*H^2, O = 4w*H^2, ID = X (identity
  re-read) or 2*O (downsample output written then read); layer mode scales
  everything but the masker read by the per-sample gate rate (skipped samples
  stay in place).  Plus the stem's tensors and the weights once per batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


@dataclass
class BlockWork:
    stage: int
    masker_macs: float
    conv_macs: float          # conv1+conv2+conv3(+downsample) MACs actually required at the measured densities
    conv_macs_dense: float
    conv_bytes: float         # activation bytes of the conv launches (no masker read)
    masker_bytes: float


@dataclass
class NetWork:
    flops_per_image: float
    dense_flops_per_image: float
    bytes_per_image: float
    conv_flops_per_image: float
    conv_bytes_per_image: float
    weight_bytes: float
    blocks: List[BlockWork]
    by_class: dict = None     # "s<stage>.<conv1|conv2|conv3|down>" -> {bytes, macs (credited, sparse), macs_dense} per image


def network_work(plans: Sequence, rho_c: Sequence[float], rho_3: Sequence[float], rho_2: Sequence[float],
                 rho_1: Sequence[float], input_size: int, stem_c: int, n_classes: int, batch: int) -> NetWork:
    """`plans` are the engine's BlockPlan objects; the rho_* are per-block floats."""
    stem_macs = 3 * stem_c * (input_size // 2) ** 2 * 49 + stem_c * (input_size // 4) ** 2 * 9
    stem_bytes = 2.0 * (3 * input_size ** 2 + 2 * stem_c * (input_size // 2) ** 2 + stem_c * (input_size // 4) ** 2)
    feat = plans[-1].outplanes
    head_macs = feat * plans[-1].H_out ** 2 + feat * n_classes
    macs = float(stem_macs + head_macs)
    dense = float(stem_macs + head_macs)
    byts = stem_bytes + 2.0 * feat * plans[-1].H_out ** 2
    conv_macs_total = conv_bytes_total = 0.0
    weight_bytes = 2.0 * (3 * stem_c * 49 + feat * n_classes)
    blocks = []
    by_class = {}

    def add(tag, byts_, macs_, dense_):
        d = by_class.setdefault(tag, {"bytes": 0.0, "macs": 0.0, "macs_dense": 0.0})
        d["bytes"] += byts_
        d["macs"] += macs_
        d["macs_dense"] += dense_

    for i, p in enumerate(plans):
        rc, r3, r2, r1 = rho_c[i], rho_3[i], rho_2[i], rho_1[i]
        hi, ho = p.H_in, p.H_out
        X = p.inplanes * hi * hi
        I1, I2, O = p.width * hi * hi, p.width * ho * ho, p.outplanes * ho * ho
        c1 = p.inplanes * p.width * hi * hi
        c2 = 9 * p.width * p.width * ho * ho
        c3 = p.width * p.outplanes * ho * ho
        ds = p.inplanes * p.outplanes * ho * ho if p.wd is not None else 0
        blk = p.module
        m_macs = 0.0
        if p.use_c:
            mk = blk.masker_channel
            m_macs += (X + mk.conv_flops) if hasattr(mk, "conv_flops") else \
                (mk.conv[0].weight.shape[0] * hi * hi + mk.masker_flops)
        if p.use_s:
            S = min(p.mask_size, hi)
            m_macs += p.inplanes * S * S + blk.masker_spatial.conv_flops_pp * S * S
        mac1, mac2, mac3 = c1 * rc * r1, c2 * rc * rc * r2, c3 * rc * r3
        conv = mac1 + mac2 + mac3 + ds
        conv_dense = c1 + c2 + c3 + ds
        gate = r3 if p.mode == "layer" else 1.0          # per-sample skip: nothing of the block is touched
        ident = 2 * O if p.wd is not None else X
        if p.mode == "layer":          # all-or-nothing per sample: the per-sample gate rate scales the block once
            r1 = r2 = 1.0
        conv_b = 2.0 * (X * r1 + 2 * I1 * rc * r1 + 2 * I2 * rc * r2 + O)
        if p.mode == "layer":
            conv_b = conv_b * gate + 2.0 * (ident if p.wd is not None else X * gate)
        else:
            conv_b += 2.0 * ident
        masker_b = 2.0 * X
        w_b = 2.0 * (c1 / (hi * hi) + c2 / (ho * ho) + c3 / (ho * ho) + (ds / (ho * ho) if ds else 0))
        # the same bytes / MACs split by launch (sums to conv_b / conv): conv1 reads X writes a1, conv2 reads a1 writes a2,
        # conv3 reads a2 + the identity and writes O, the downsample writes O (its read of X is the one already counted)
        st = f"s{p.stage + 1}."
        wsh = 2.0 / max(batch, 1)
        add(st + "conv1", 2.0 * gate * (X * r1 + I1 * rc * r1) + wsh * c1 / (hi * hi), mac1, c1)
        add(st + "conv2", 2.0 * gate * (I1 * rc * r1 + I2 * rc * r2) + wsh * c2 / (ho * ho), mac2, c2)
        add(st + "conv3", 2.0 * gate * (I2 * rc * r2 + O + (O if p.wd is not None else X)) + wsh * c3 / (ho * ho), mac3, c3)
        if p.wd is not None:
            add(st + "down", 2.0 * O + (2.0 * O * (1.0 - gate) if p.mode == "layer" else 0.0) + wsh * ds / (ho * ho), ds, ds)
        weight_bytes += w_b
        macs += m_macs + conv
        dense += m_macs + conv_dense
        byts += conv_b + masker_b
        conv_macs_total += conv
        conv_bytes_total += conv_b
        blocks.append(BlockWork(p.stage, m_macs, conv, conv_dense, conv_b, masker_b))
    byts += weight_bytes / max(batch, 1)
    return NetWork(flops_per_image=2.0 * macs, dense_flops_per_image=2.0 * dense, bytes_per_image=byts,
                   conv_flops_per_image=2.0 * conv_macs_total,
                   conv_bytes_per_image=conv_bytes_total + (weight_bytes / max(batch, 1)),
                   weight_bytes=weight_bytes, blocks=blocks, by_class=by_class)
