"""DECLARED SELF-ORACLE for the AdaViT token / head / layer-skip block (BASELINE.json configs[3]).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference tree holds NO AdaViT model code: `README.md:24-26` only links
the external repository github.com/MengLcool/AdaViT (no commit pinned, not vendored, no network here).  The only in-tree
description of the block is the operator list of the latency simulator, `DyNetSimulator/adavit/simulate_adavit.py:83-182`:
per block a layer policy `Linear(D, 2)` and a head policy `Linear(D, heads)` on the policy (class) token, a token score
`Linear(D, 1)` on the L-1 patch tokens, attention over `int(L * token_density)` tokens x `int(heads * head_density)` heads,
projection / MLP on the selected tokens, everything of a sub-layer scaled by the layer policy.  This file restates that
operator list as a plain torch-fp32 MASKED-DENSE forward on the DeiT-S geometry BASELINE.json names (D = 384, 6 heads,
12 blocks, MLP ratio 4, 16x16 patches, L = 197) following the published AdaViT algorithm (Meng et al., CVPR 2022, section 3):

  p        = LayerNorm_policy(x[:, 0])                              policy token = class token
  layer    = Linear(D, 2)(p)     >= 0     -> (run attention, run MLP)          per sample
  head     = Linear(D, H)(p)     >= 0     -> head h contributes                 per sample
  token    = Linear(D, 1)(LayerNorm1(x)[:, 1:]) >= 0 ; class token always kept  per token
  eval decision = sigmoid(logit) > 0.5  <=>  logit > 0; ties (logit == 0) keep, as the LAUD maskers do (utils.py:59-60)
  attention: keys AND queries restricted to kept tokens (dropped keys masked with -inf before the softmax), a dropped
             head's output is zero before the projection, a dropped token / sample receives no update (identity)
  MLP      : same token mask, scaled by the MLP layer decision

Because nothing here can be checked against reference outputs, results obtained against this file are labelled
"self-oracle" everywhere (the bench line's `parity.kind`).  What pins it instead (tests/test_adavit_oracle.py):
  * with every gate open the forward equals the stock ViT arithmetic of the `transformers` package (ViTForImageClassification
    with DeiT-S dimensions and the same weights), an independent implementation;
  * the masked-dense forward equals an independently written REALLY SPARSE torch forward (`sparse_forward`: gathers the kept
    tokens / heads / samples and computes only those) - the property the CUDA path relies on.

Who may use it: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs.  Nothing under `laudnet_b200/` imports it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6          # timm / DeiT LayerNorm epsilon


@dataclass(frozen=True)
class AdaViTCfg:
    img_size: int = 224
    patch_size: int = 16
    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    mlp_ratio: float = 4.0
    num_classes: int = 1000
    keep_layers: int = 1          # the first blocks run without policies (AdaViT keeps the earliest layers static)
    ada_token: bool = True
    ada_head: bool = True
    ada_layer: bool = True

    @property
    def num_patches(self) -> int:
        return (self.img_size // self.patch_size) ** 2

    @property
    def seq_len(self) -> int:
        return self.num_patches + 1

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    def has_policy(self, i: int) -> bool:
        return i >= self.keep_layers and (self.ada_token or self.ada_head or self.ada_layer)

    def kwargs(self) -> dict:
        return dict(img_size=self.img_size, patch_size=self.patch_size, embed_dim=self.embed_dim, depth=self.depth,
                    num_heads=self.num_heads, mlp_ratio=self.mlp_ratio, num_classes=self.num_classes,
                    keep_layers=self.keep_layers, ada_token=self.ada_token, ada_head=self.ada_head, ada_layer=self.ada_layer)


def state_dict_shapes(cfg: AdaViTCfg) -> Dict[str, Tuple[int, ...]]:
    """Key layout: timm's VisionTransformer (DeiT) names plus the three policy heads of an AdaViT block."""
    D, Hd, L = cfg.embed_dim, cfg.hidden, cfg.seq_len
    s: Dict[str, Tuple[int, ...]] = {
        "cls_token": (1, 1, D), "pos_embed": (1, L, D),
        "patch_embed.proj.weight": (D, 3, cfg.patch_size, cfg.patch_size), "patch_embed.proj.bias": (D,),
        "norm.weight": (D,), "norm.bias": (D,), "head.weight": (cfg.num_classes, D), "head.bias": (cfg.num_classes,),
    }
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        s.update({p + "norm1.weight": (D,), p + "norm1.bias": (D,), p + "attn.qkv.weight": (3 * D, D), p + "attn.qkv.bias": (3 * D,),
                  p + "attn.proj.weight": (D, D), p + "attn.proj.bias": (D,), p + "norm2.weight": (D,), p + "norm2.bias": (D,),
                  p + "mlp.fc1.weight": (Hd, D), p + "mlp.fc1.bias": (Hd,), p + "mlp.fc2.weight": (D, Hd), p + "mlp.fc2.bias": (D,)})
        if cfg.has_policy(i):
            s.update({p + "norm_policy.weight": (D,), p + "norm_policy.bias": (D,)})
            if cfg.ada_layer:
                s.update({p + "layer_select.weight": (2, D), p + "layer_select.bias": (2,)})
            if cfg.ada_head:
                s.update({p + "head_select.weight": (cfg.num_heads, D), p + "head_select.bias": (cfg.num_heads,)})
            if cfg.ada_token:
                s.update({p + "token_select.weight": (1, D), p + "token_select.bias": (1,)})
    return s


def _ln(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], LN_EPS)


def embed(sd: Dict[str, Tensor], cfg: AdaViTCfg, img: Tensor) -> Tensor:
    """patch_embed (conv PxP / P) -> [B, L, D] with the class token in front, + pos_embed."""
    t = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    t = t.flatten(2).transpose(1, 2)
    x = torch.cat([sd["cls_token"].expand(img.shape[0], -1, -1), t], dim=1)
    return x + sd["pos_embed"]


@dataclass
class BlockPolicy:
    token: Tensor      # bool [B, L]   (class token always True)
    head: Tensor       # bool [B, H]
    layer: Tensor      # bool [B, 2]   (attention, MLP)
    token_logits: Optional[Tensor] = None    # fp32 [B, L-1]
    head_logits: Optional[Tensor] = None     # fp32 [B, H]
    layer_logits: Optional[Tensor] = None    # fp32 [B, 2]


@dataclass
class BlockTrace:
    x_in: Tensor
    policy: BlockPolicy
    x_mid: Tensor      # after the attention sub-layer
    x_out: Tensor


def block_policy(x: Tensor, sd: Dict[str, Tensor], cfg: AdaViTCfg, i: int) -> BlockPolicy:
    """The three decisions of block i (all-open for the static blocks)."""
    B, L, _ = x.shape
    p = f"blocks.{i}."
    tok = torch.ones(B, L, dtype=torch.bool)
    head = torch.ones(B, cfg.num_heads, dtype=torch.bool)
    layer = torch.ones(B, 2, dtype=torch.bool)
    pol = BlockPolicy(tok, head, layer)
    if not cfg.has_policy(i):
        return pol
    pt = _ln(x[:, 0], sd, p + "norm_policy.")
    if cfg.ada_layer:
        pol.layer_logits = F.linear(pt, sd[p + "layer_select.weight"], sd[p + "layer_select.bias"])
        pol.layer = pol.layer_logits >= 0
    if cfg.ada_head:
        pol.head_logits = F.linear(pt, sd[p + "head_select.weight"], sd[p + "head_select.bias"])
        pol.head = pol.head_logits >= 0
    if cfg.ada_token:
        y = _ln(x[:, 1:], sd, p + "norm1.")
        pol.token_logits = F.linear(y, sd[p + "token_select.weight"], sd[p + "token_select.bias"]).squeeze(-1)
        pol.token = torch.cat([torch.ones(B, 1, dtype=torch.bool), pol.token_logits >= 0], dim=1)
    return pol


def block_forward(x: Tensor, sd: Dict[str, Tensor], cfg: AdaViTCfg, i: int, forced: Optional[BlockPolicy] = None,
                  trace: Optional[List[BlockTrace]] = None) -> Tensor:
    """One AdaViT block, masked-dense: everything is computed and multiplied by the 0/1 decisions."""
    B, L, D = x.shape
    H, d = cfg.num_heads, D // cfg.num_heads
    p = f"blocks.{i}."
    pol = forced if forced is not None else block_policy(x, sd, cfg, i)
    tok = pol.token.to(x.dtype)
    y = _ln(x, sd, p + "norm1.")
    qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).view(B, L, 3, H, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]                                      # [B, H, L, d]
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    s = s.masked_fill(~pol.token[:, None, None, :], float("-inf"))        # dropped keys
    o = torch.softmax(s, dim=-1) @ v
    o = o * pol.head[:, :, None, None].to(x.dtype)                        # dropped heads contribute nothing
    o = F.linear(o.transpose(1, 2).reshape(B, L, D), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x_mid = x + o * tok[:, :, None] * pol.layer[:, 0, None, None].to(x.dtype)
    z = _ln(x_mid, sd, p + "norm2.")
    m = F.linear(F.gelu(F.linear(z, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    x_out = x_mid + m * tok[:, :, None] * pol.layer[:, 1, None, None].to(x.dtype)
    if trace is not None:
        trace.append(BlockTrace(x, pol, x_mid, x_out))
    return x_out


def head_forward(x: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    return F.linear(_ln(x[:, 0], sd, "norm."), sd["head.weight"], sd["head.bias"])


def forward(sd: Dict[str, Tensor], cfg: AdaViTCfg, img: Tensor, traces: Optional[List[BlockTrace]] = None,
            forced: Optional[List[Optional[BlockPolicy]]] = None):
    """-> (logits [B, classes], token_select bool [B, depth, L], head_select bool [B, depth, H], layer_select bool [B, depth, 2])."""
    x = embed(sd, cfg, img)
    tr: List[BlockTrace] = [] if traces is None else traces
    for i in range(cfg.depth):
        x = block_forward(x, sd, cfg, i, forced[i] if forced is not None else None, tr)
    logits = head_forward(x, sd)
    return (logits, torch.stack([t.policy.token for t in tr], 1), torch.stack([t.policy.head for t in tr], 1),
            torch.stack([t.policy.layer for t in tr], 1))


# --------------------------------------------------------------------------
# The same block computed REALLY sparsely (per sample: only the kept tokens, heads and sub-layers).  Written
# independently of block_forward; tests assert both agree - the exactness argument of the CUDA execution.
# --------------------------------------------------------------------------
def sparse_block_forward(x: Tensor, sd: Dict[str, Tensor], cfg: AdaViTCfg, i: int, pol: BlockPolicy) -> Tensor:
    B, L, D = x.shape
    H, d = cfg.num_heads, D // cfg.num_heads
    p = f"blocks.{i}."
    out = x.clone()
    Wq, bq = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
    for b in range(B):
        rows = torch.nonzero(pol.token[b]).squeeze(1)
        if bool(pol.layer[b, 0]):
            y = _ln(out[b, rows], sd, p + "norm1.")                       # only the kept tokens are normalised
            o = torch.zeros(rows.numel(), D)
            for h in torch.nonzero(pol.head[b]).squeeze(1).tolist():      # only the kept heads
                sl = slice(h * d, (h + 1) * d)
                q = y @ Wq[0 * D:1 * D][sl].T + bq[0 * D:1 * D][sl]
                k = y @ Wq[1 * D:2 * D][sl].T + bq[1 * D:2 * D][sl]
                v = y @ Wq[2 * D:3 * D][sl].T + bq[2 * D:3 * D][sl]
                o[:, sl] = torch.softmax((q @ k.T) * (d ** -0.5), dim=-1) @ v
            out[b, rows] = out[b, rows] + o @ sd[p + "attn.proj.weight"].T + sd[p + "attn.proj.bias"]
        if bool(pol.layer[b, 1]):
            z = _ln(out[b, rows], sd, p + "norm2.")
            hdn = F.gelu(z @ sd[p + "mlp.fc1.weight"].T + sd[p + "mlp.fc1.bias"])
            out[b, rows] = out[b, rows] + hdn @ sd[p + "mlp.fc2.weight"].T + sd[p + "mlp.fc2.bias"]
    return out


# --------------------------------------------------------------------------
# FLOP accounting (multiply-accumulates x 2), following the operator list of simulate_adavit.py:83-182
# --------------------------------------------------------------------------
def block_macs(cfg: AdaViTCfg, n_tok: float, n_head: float, attn_on: float, mlp_on: float) -> float:
    """MACs of one block for a sample that keeps `n_tok` tokens and `n_head` heads (batch means work too)."""
    D, d, Hd = cfg.embed_dim, cfg.embed_dim // cfg.num_heads, cfg.hidden
    attn = n_tok * D * 3 * d * n_head + 2 * n_tok * n_tok * d * n_head + n_tok * d * n_head * D
    mlp = 2 * n_tok * D * Hd
    return attn_on * attn + mlp_on * mlp


def dense_macs(cfg: AdaViTCfg) -> float:
    L = cfg.seq_len
    pe = cfg.num_patches * cfg.embed_dim * 3 * cfg.patch_size ** 2
    return pe + cfg.depth * block_macs(cfg, L, cfg.num_heads, 1.0, 1.0) + cfg.embed_dim * cfg.num_classes


def sparse_macs(cfg: AdaViTCfg, token: Tensor, head: Tensor, layer: Tensor) -> Tensor:
    """Per-sample MACs [B] from the decisions (token [B,depth,L], head [B,depth,H], layer [B,depth,2])."""
    nt, nh = token.sum(-1).double(), head.sum(-1).double()
    D, d, Hd = cfg.embed_dim, cfg.embed_dim // cfg.num_heads, cfg.hidden
    attn = nt * D * 3 * d * nh + 2 * nt * nt * d * nh + nt * d * nh * D
    mlp = 2 * nt * D * Hd
    per_block = layer[..., 0].double() * attn + layer[..., 1].double() * mlp
    pe = cfg.num_patches * cfg.embed_dim * 3 * cfg.patch_size ** 2
    return pe + per_block.sum(1) + cfg.embed_dim * cfg.num_classes
