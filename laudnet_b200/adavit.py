"""AdaViT (DeiT backbone) with token / head / layer skipping EXECUTED on the CUDA path (BASELINE.json configs[3]).

Self-oracle scope: the reference tree contains no AdaViT model code (README.md:24-26 links the external
repository), only the operator list of `DyNetSimulator/adavit/simulate_adavit.py:83-182`.  The module keeps the layout a
user of AdaViT / timm expects - `state_dict` keys of timm's `VisionTransformer` (`cls_token`, `pos_embed`,
`patch_embed.proj`, `blocks.{i}.norm1 | attn.qkv | attn.proj | norm2 | mlp.fc1 | mlp.fc2`, `norm`, `head`) plus the three
policy heads of an AdaViT block (`blocks.{i}.norm_policy | layer_select | head_select | token_select`) - and its forward
returns `(logits, token_select, head_select, layer_select)`.  The arithmetic it must match is `oracle/adavit_oracle.py`
(declared self-oracle, "parity unpinned").

Execution (include/laud_adavit.h): the residual stream stays fp32 `[B, L, D]`; per block one policy pass decides, two
ordered scans place every kept token of every active sample in a COMPACT row list, LayerNorm writes only those rows
(fp16), the tcgen05 token GEMM runs over the compact rows with a device-side row count (QKV with whole head tiles
dropped, proj and fc2 adding straight into the residual stream at the tokens' own rows, fc1 with GELU), attention runs
per (sample, kept head) over the kept tokens.  No host synchronisation: the forward is CUDA-graph capturable.
There is no CPU path and no training path (`LaudError`).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import LaudError, check, ptr, stream_ptr

LN_EPS = 1e-6


class _Attn(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _PatchEmbed(nn.Module):
    def __init__(self, patch: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, stride=patch)


class AdaBlock(nn.Module):
    """Parameter container of one block (timm `Block` names + the AdaViT policy heads)."""

    def __init__(self, dim: int, heads: int, hidden: int, policy: bool, ada_token: bool, ada_head: bool, ada_layer: bool):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _Mlp(dim, hidden)
        self.has_policy = policy
        if policy:
            self.norm_policy = nn.LayerNorm(dim, eps=LN_EPS)
            if ada_layer:
                self.layer_select = nn.Linear(dim, 2)
            if ada_head:
                self.head_select = nn.Linear(dim, heads)
            if ada_token:
                self.token_select = nn.Linear(dim, 1)


@dataclass
class BlockKeep:
    """Per-block record of a forward run with `keep=[]` (tests / parity): decisions, logits and the stream after the block."""
    token: torch.Tensor
    head: torch.Tensor
    layer: torch.Tensor
    token_logits: torch.Tensor
    head_logits: torch.Tensor
    layer_logits: torch.Tensor
    x_out: torch.Tensor


class AdaViT(nn.Module):
    def __init__(self, img_size: int = 224, patch_size: int = 16, embed_dim: int = 384, depth: int = 12, num_heads: int = 6,
                 mlp_ratio: float = 4.0, num_classes: int = 1000, keep_layers: int = 1, ada_token: bool = True,
                 ada_head: bool = True, ada_layer: bool = True):
        super().__init__()
        if embed_dim != num_heads * 64:
            raise LaudError("AdaViT: the attention kernel is built for head dimension 64 (DeiT-Ti/S/B)")
        if embed_dim % 64 or int(embed_dim * mlp_ratio) % 64 or patch_size % 8 or img_size % patch_size:
            raise LaudError("AdaViT: embed_dim and the MLP width must be multiples of 64, patch_size of 8")
        self.img_size, self.patch_size, self.embed_dim, self.depth, self.num_heads = img_size, patch_size, embed_dim, depth, num_heads
        self.hidden, self.num_classes, self.keep_layers = int(embed_dim * mlp_ratio), num_classes, keep_layers
        self.ada_token, self.ada_head, self.ada_layer = ada_token, ada_head, ada_layer
        self.num_patches = (img_size // patch_size) ** 2
        self.seq_len = self.num_patches + 1
        if self.seq_len > 208:
            raise LaudError("AdaViT: at most 208 tokens per image (the attention kernel keeps a score row in registers)")
        if num_classes % 8:
            raise LaudError("AdaViT: num_classes must be a multiple of 8")
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.seq_len, embed_dim))
        self.patch_embed = _PatchEmbed(patch_size, embed_dim)
        any_policy = ada_token or ada_head or ada_layer
        self.blocks = nn.ModuleList([AdaBlock(embed_dim, num_heads, self.hidden, any_policy and i >= keep_layers, ada_token,
                                              ada_head, ada_layer) for i in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=LN_EPS)
        self.head = nn.Linear(embed_dim, num_classes)
        self._packed: Optional[dict] = None
        self._ws: Dict[int, dict] = {}
        self._graphs: Dict[int, "GraphedAdaViT"] = {}
        self.head_tile_skip = True          # drop whole per-head n-tiles of the QKV projection (A/B switch)
        # fc1 -> GELU -> fc2 in one kernel, hidden activations on chip (laud_adavit_mlp_fused); False: two token GEMMs
        self.fused_mlp = embed_dim in (128, 256, 384) and self.hidden % 128 == 0 and os.environ.get("LAUD_ADAVIT_MLP") != "split"
        self.profile: Optional[list] = None  # measurement aid: a list makes _run record (tag, CUDA event) before every launch

    # ------------------------------------------------------------------ parameters -> device layouts (once)
    def _invalidate(self):
        self._packed, self._ws = None, {}
        for g in self._graphs.values():
            g.valid = False
        self._graphs = {}

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out

    def prepare(self) -> dict:
        if self._packed is not None:
            return self._packed
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise LaudError("AdaViT.prepare: parameters must live on a CUDA device - there is no CPU path")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        D, H = self.embed_dim, self.num_heads
        # QKV rows head-major: head h holds q | k | v at rows h*192 (the attention kernel's layout; one n-tile per head)
        perm = torch.arange(3 * D, device=dev).view(3, H, 64).permute(1, 0, 2).reshape(-1)
        P: dict = {"patch_w": f16(self.patch_embed.proj.weight.reshape(D, -1)), "patch_b": f32(self.patch_embed.proj.bias),
                   "pos": f32(self.pos_embed[0]), "cls": f32(self.cls_token[0, 0]), "norm_w": f32(self.norm.weight),
                   "norm_b": f32(self.norm.bias), "head_w": f16(self.head.weight), "head_b": f32(self.head.bias), "blocks": []}
        for blk in self.blocks:
            q: dict = {"n1_w": f32(blk.norm1.weight), "n1_b": f32(blk.norm1.bias), "n2_w": f32(blk.norm2.weight),
                       "n2_b": f32(blk.norm2.bias), "qkv_w": f16(blk.attn.qkv.weight[perm]), "qkv_b": f32(blk.attn.qkv.bias[perm]),
                       "proj_w": f16(blk.attn.proj.weight), "proj_b": f32(blk.attn.proj.bias), "fc1_w": f16(blk.mlp.fc1.weight),
                       "fc1_b": f32(blk.mlp.fc1.bias), "fc2_w": f16(blk.mlp.fc2.weight), "fc2_b": f32(blk.mlp.fc2.bias)}
            for k in ("np_w", "np_b", "ls_w", "ls_b", "hs_w", "hs_b", "ts_w", "ts_b"):
                q[k] = None
            if blk.has_policy:
                q["np_w"], q["np_b"] = f32(blk.norm_policy.weight), f32(blk.norm_policy.bias)
                if self.ada_layer:
                    q["ls_w"], q["ls_b"] = f32(blk.layer_select.weight), f32(blk.layer_select.bias)
                if self.ada_head:
                    q["hs_w"], q["hs_b"] = f32(blk.head_select.weight), f32(blk.head_select.bias)
                if self.ada_token:
                    q["ts_w"], q["ts_b"] = f32(blk.token_select.weight.reshape(-1)), f32(blk.token_select.bias)
            P["blocks"].append(q)
        self._packed = P
        return P

    def workspace(self, B: int) -> dict:
        ws = self._ws.get(B)
        if ws is not None:
            return ws
        dev = self.cls_token.device
        L, D, H, Hd, n, NP = self.seq_len, self.embed_dim, self.num_heads, self.hidden, self.depth, self.num_patches
        z = lambda *s, dt=torch.float16: torch.zeros(*s, dtype=dt, device=dev)
        rows = B * L
        ws = {"x": z(B, L, D, dt=torch.float32), "patches": z(B * NP, 3 * self.patch_size ** 2),
              "patch_rows": (torch.arange(B, device=dev, dtype=torch.int32)[:, None] * L + 1 +
                             torch.arange(NP, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous(),
              "y": z(rows, D), "qkv": z(rows, 3 * D), "o": z(rows, D), "hdn": z(rows, Hd),
              "rows_a": z(rows, dt=torch.int32), "samp_a": z(rows, dt=torch.int32), "rows_m": z(rows, dt=torch.int32),
              "tok": z(n, B, L, dt=torch.uint8), "cnt": z(n, B, dt=torch.int32), "head": z(n, B, H, dt=torch.uint8),
              "layer": z(n, B, 2, dt=torch.uint8), "off_a": z(n, B + 1, dt=torch.int32), "off_m": z(n, B + 1, dt=torch.int32),
              "tok_lg": z(n, B, L, dt=torch.float32), "head_lg": z(n, B, H, dt=torch.float32), "layer_lg": z(n, B, 2, dt=torch.float32),
              "cls_mask": z(B, L, dt=torch.uint8), "cls_off": torch.arange(B + 1, device=dev, dtype=torch.int32),
              "cls_rows": torch.arange(B, device=dev, dtype=torch.int32), "ycls": z(B, D),
              "logits": z(B, self.num_classes, dt=torch.float32)}
        ws["cls_mask"][:, 0] = 1
        self._ws[B] = ws
        return ws

    # ------------------------------------------------------------------ launches
    @staticmethod
    def _gemm(a, w, bias, rows_max, K, N, st, row_cnt=None, act=_lib.ACT_NONE, out=None, resid=None, ldres=0, row_idx=None,
              col_gate=None, gate_ld=0, row_sample=None, bn=0, cta_pair=0):
        d = _lib.TokGemmDesc()
        d.a, d.lda, d.w, d.bias = ptr(a), K, ptr(w), ptr(bias)
        d.rows_max, d.K, d.N = rows_max, K, N
        d.row_cnt, d.act = ptr(row_cnt), act
        d.out, d.ldo = ptr(out), (N if out is not None else 0)
        d.resid, d.ldres, d.row_idx = ptr(resid), ldres, ptr(row_idx)
        d.col_gate, d.gate_ld, d.row_sample, d.bn, d.cta_pair = ptr(col_gate), gate_ld, ptr(row_sample), bn, cta_pair
        check(_lib.lib().laud_tok_gemm(C.byref(d), st), "laud_tok_gemm")

    def _run(self, x_img: torch.Tensor, keep: Optional[List[BlockKeep]] = None, forced: Optional[Sequence] = None,
             x_tokens: Optional[torch.Tensor] = None, only_block: Optional[int] = None) -> dict:
        """One forward on the current stream.  forced[i] = (token bool [B,L], head bool [B,H], layer bool [B,2]) or None
        installs given decisions (teacher forcing, tests); x_tokens / only_block run a single block on a given stream."""
        lib = _lib.lib()
        P, B = self.prepare(), x_img.shape[0] if x_tokens is None else x_tokens.shape[0]
        ws, st = self.workspace(B), stream_ptr()
        L, D, H, Hd = self.seq_len, self.embed_dim, self.num_heads, self.hidden
        rows, x = B * L, ws["x"]
        prof = self.profile

        def mark(tag):
            if prof is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                prof.append((tag, e))

        if x_tokens is None:
            mark("embed")
            check(lib.laud_vit_patchify(ptr(x_img), B, self.img_size, self.patch_size, ptr(ws["patches"]), st), "laud_vit_patchify")
            check(lib.laud_vit_init_tokens(ptr(x), B, L, D, ptr(P["pos"]), ptr(P["cls"]), st), "laud_vit_init_tokens")
            self._gemm(ws["patches"], P["patch_w"], P["patch_b"], B * self.num_patches, 3 * self.patch_size ** 2, D, st,
                       resid=x, ldres=D, row_idx=ws["patch_rows"])
        else:
            x.copy_(x_tokens)
        blocks = range(self.depth) if only_block is None else [only_block]
        for i in blocks:
            q = P["blocks"][i]
            tok, cnt, head, layer = ws["tok"][i], ws["cnt"][i], ws["head"][i], ws["layer"][i]
            off_a, off_m = ws["off_a"][i], ws["off_m"][i]
            mark("policy")
            check(lib.laud_adavit_policy(ptr(x), B, L, D, H, LN_EPS, ptr(q["n1_w"]), ptr(q["n1_b"]), ptr(q["ts_w"]), ptr(q["ts_b"]),
                                         ptr(q["np_w"]), ptr(q["np_b"]), ptr(q["ls_w"]), ptr(q["ls_b"]), ptr(q["hs_w"]),
                                         ptr(q["hs_b"]), ptr(tok), ptr(cnt), ptr(head), ptr(layer), ptr(ws["tok_lg"][i]),
                                         ptr(ws["head_lg"][i]), ptr(ws["layer_lg"][i]), st), "laud_adavit_policy")
            if forced is not None and forced[i] is not None:
                ft, fh, fl = forced[i]
                tok.copy_(ft.to(torch.uint8))
                head.copy_(fh.to(torch.uint8))
                layer.copy_(fl.to(torch.uint8))
                cnt.copy_(ft.sum(1).to(torch.int32))
            check(lib.laud_adavit_lists(ptr(cnt), ptr(layer), B, ptr(off_a), ptr(off_m), st), "laud_adavit_lists")
            # ---- attention sub-layer on the kept tokens of the samples that run it
            check(lib.laud_adavit_row_lists(ptr(tok), B, L, ptr(off_a), ptr(off_m), ptr(ws["rows_a"]), ptr(ws["samp_a"]),
                                            ptr(ws["rows_m"]), st), "laud_adavit_row_lists")
            mark("ln_gather")
            check(lib.laud_adavit_ln_rows(ptr(x), D, LN_EPS, ptr(q["n1_w"]), ptr(q["n1_b"]), ptr(ws["rows_a"]), ptr(off_a[B:]), rows,
                                          ptr(ws["y"]), st), "laud_adavit_ln_rows")
            gate = head if self.head_tile_skip else None
            mark("gemm_qkv")
            self._gemm(ws["y"], q["qkv_w"], q["qkv_b"], rows, D, 3 * D, st, row_cnt=off_a[B:], out=ws["qkv"], bn=192,
                       col_gate=gate, gate_ld=H, row_sample=ws["samp_a"] if gate is not None else None)
            mark("attention")
            check(lib.laud_adavit_attention(ptr(ws["qkv"]), 3 * D, ptr(off_a), ptr(head), B, H, L, ptr(ws["o"]), st),
                  "laud_adavit_attention")
            mark("gemm_proj")
            self._gemm(ws["o"], q["proj_w"], q["proj_b"], rows, D, D, st, row_cnt=off_a[B:], resid=x, ldres=D, row_idx=ws["rows_a"])
            # ---- MLP sub-layer
            mark("ln_gather")
            check(lib.laud_adavit_ln_rows(ptr(x), D, LN_EPS, ptr(q["n2_w"]), ptr(q["n2_b"]), ptr(ws["rows_m"]), ptr(off_m[B:]), rows,
                                          ptr(ws["y"]), st), "laud_adavit_ln_rows")
            if self.fused_mlp:
                mark("gemm_mlp_fused")
                check(lib.laud_adavit_mlp_fused(ptr(ws["y"]), rows, D, Hd, ptr(off_m[B:]), ptr(q["fc1_w"]), ptr(q["fc1_b"]),
                                                ptr(q["fc2_w"]), ptr(q["fc2_b"]), ptr(x), D, ptr(ws["rows_m"]), st),
                      "laud_adavit_mlp_fused")
            else:
                mark("gemm_fc1")
                self._gemm(ws["y"], q["fc1_w"], q["fc1_b"], rows, D, Hd, st, row_cnt=off_m[B:], act=_lib.ACT_GELU, out=ws["hdn"])
                mark("gemm_fc2")
                self._gemm(ws["hdn"], q["fc2_w"], q["fc2_b"], rows, Hd, D, st, row_cnt=off_m[B:], resid=x, ldres=D, row_idx=ws["rows_m"])
            if keep is not None:
                keep.append(BlockKeep(tok.bool().clone(), head.bool().clone(), layer.bool().clone(), ws["tok_lg"][i].clone(),
                                      ws["head_lg"][i].clone(), ws["layer_lg"][i].clone(), x.clone()))
        if only_block is None:
            # classifier: LayerNorm of the class tokens only, then the fc as a token GEMM into zeroed fp32 logits
            mark("head")
            check(lib.laud_adavit_ln_gather(ptr(x), B, L, D, LN_EPS, ptr(P["norm_w"]), ptr(P["norm_b"]), ptr(ws["cls_mask"]),
                                            ptr(ws["cls_off"]), ptr(ws["ycls"]), None, None, st), "laud_adavit_ln_gather")
            ws["logits"].zero_()
            self._gemm(ws["ycls"], P["head_w"], P["head_b"], B, D, self.num_classes, st, resid=ws["logits"],
                       ldres=self.num_classes, row_idx=ws["cls_rows"])
        mark("end")
        return ws

    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise LaudError("AdaViT: training mode (Gumbel policies) is not part of the inference hot path")
        if not x.is_cuda:
            raise LaudError("AdaViT.forward: expected a CUDA tensor - there is no CPU path")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_size or x.shape[3] != self.img_size:
            raise LaudError(f"AdaViT.forward: expected [B, 3, {self.img_size}, {self.img_size}], got {tuple(x.shape)}")
        return x.to(torch.float16).contiguous()

    def forward(self, x: torch.Tensor, keep: Optional[List[BlockKeep]] = None, forced: Optional[Sequence] = None):
        """-> (logits fp32 [B, classes], token_select bool [B, depth, L], head_select bool [B, depth, H],
        layer_select bool [B, depth, 2]) - fresh tensors."""
        x = self._check_input(x)
        with torch.cuda.device(x.device), torch.no_grad():
            ws = self._run(x, keep, forced)
            return (ws["logits"].clone(), ws["tok"].bool().transpose(0, 1).contiguous(), ws["head"].bool().transpose(0, 1).contiguous(),
                    ws["layer"].bool().transpose(0, 1).contiguous())

    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        """Logits only, in the workspace buffer (valid until the next forward of this batch size)."""
        x = self._check_input(x)
        with torch.cuda.device(x.device), torch.no_grad():
            return self._run(x)["logits"]

    def run_block(self, i: int, x_tokens: torch.Tensor, forced=None) -> torch.Tensor:
        """Block i alone on a given fp32 token stream [B, L, D] (teacher-forced block tests)."""
        if not x_tokens.is_cuda:
            raise LaudError("AdaViT.run_block: expected a CUDA tensor")
        with torch.cuda.device(x_tokens.device), torch.no_grad():
            f = [None] * self.depth
            f[i] = forced
            ws = self._run(None, None, f if forced is not None else None, x_tokens=x_tokens.float().contiguous(), only_block=i)
            return ws["x"].clone()

    def capture(self, batch: int) -> "GraphedAdaViT":
        g = self._graphs.get(batch)
        if g is None or not g.valid:
            g = GraphedAdaViT(self, batch)
            self._graphs[batch] = g
        return g

    def decisions(self, batch: int):
        """(token u8 [depth,B,L], head u8 [depth,B,H], layer u8 [depth,B,2]) of the last forward at this batch size."""
        ws = self.workspace(batch)
        return ws["tok"], ws["head"], ws["layer"]


class GraphedAdaViT:
    """The forward captured in one CUDA graph (static input buffer; replay() returns the workspace logits)."""

    def __init__(self, model: AdaViT, batch: int):
        self.model, self.batch, self.valid = model, batch, True
        dev = model.cls_token.device
        self.x = torch.zeros(batch, 3, model.img_size, model.img_size, dtype=torch.float16, device=dev)
        with torch.cuda.device(dev), torch.no_grad():
            model._run(self.x)                       # warm-up: first-call initialisation must not happen inside a capture
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.logits = model._run(self.x)["logits"]

    def replay(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.valid:
            raise LaudError("GraphedAdaViT: the model's parameters changed after the capture - call model.capture() again")
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.logits


def ada_deit_small_patch16_224(**kw) -> AdaViT:
    """AdaViT on DeiT-S (BASELINE.json configs[3]): D = 384, 6 heads, 12 blocks, MLP ratio 4."""
    return AdaViT(**{**dict(img_size=224, patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4.0), **kw})


def ada_deit_tiny_patch16_224(**kw) -> AdaViT:
    return AdaViT(**{**dict(img_size=224, patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4.0), **kw})


def ada_deit_base_patch16_224(**kw) -> AdaViT:
    return AdaViT(**{**dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0), **kw})
