"""GPU tests of the AdaViT token / head / layer-skip path (include/laud_adavit.h, laudnet_b200/adavit.py) against the
DECLARED self-oracle oracle/adavit_oracle.py ("parity unpinned": the reference holds no AdaViT code).

Bars: compact row lists / offsets / counts bit-exact; decisions bit-exact wherever the oracle's own |logit| exceeds the
fp32 summation noise (teacher-forced inputs) resp. the fp16 activation budget (free-running); activations within 1e-3 of
max|oracle| per block on identical inputs and decisions (fp16 operands, fp32 accumulation, fp32 residual stream)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from laudnet_b200 import _lib, synth
from laudnet_b200.adavit import AdaViT, ada_deit_small_patch16_224
from oracle import adavit_oracle as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ACT_TOL = 1e-3            # per block, teacher-forced
NET_TOL = 5e-3            # logits after the whole fp16-operand trunk (12 blocks), identical decisions
POLICY_MARGIN = 1e-4      # teacher-forced inputs: a decision may differ only if |oracle logit| <= this x max|logit|
FREE_MARGIN = 5e-3        # free-running: ... the fp16 operand budget


def _rel(ours, ref):
    return ((ours.double().cpu() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-12)).item()


def _gemm(a, w, bias, rows_max, K, N, **kw):
    AdaViT._gemm(a, w, bias, rows_max, K, N, _lib.stream_ptr(), **kw)
    torch.cuda.synchronize()


# --------------------------------------------------------------------------- token GEMM
@pytest.mark.parametrize("rows,K,N,bn", [(300, 128, 192, 0), (1000, 384, 1152, 192), (517, 1536, 384, 0), (130, 384, 1000, 0),
                                         (64, 64, 64, 0), (2600, 384, 1536, 256), (333, 768, 384, 128),
                                         # weight-resident mode (K <= 384, enough m-tiles per CTA): QKV / proj / fc1 shapes
                                         (12000, 384, 1152, 192), (21000, 384, 384, 0), (9000, 384, 1536, 0), (7000, 128, 192, 0),
                                         # CTA-pair mode (streaming, >= 16 m-tiles): fc2 / patch-projection shapes, odd tile counts
                                         (6000, 1536, 384, 0), (4225, 768, 384, 0), (2600, 1536, 1000, 0), (5000, 768, 64, 0)])
@pytest.mark.parametrize("pair", [0, 1])      # 1: CTA pairs (tcgen05 cta_group::2) where the shape allows - same results
def test_tok_gemm_store_vs_torch(cuda_lib, rows, K, N, bn, pair):
    g = torch.Generator().manual_seed(rows + K + N)
    a = (torch.randn(rows, K, generator=g) * 0.5).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    cnt = torch.tensor([rows - 37 if rows > 100 else rows], dtype=torch.int32, device=DEV)
    for act in (_lib.ACT_NONE, _lib.ACT_GELU):
        out = torch.full((rows, N), 7.0, dtype=torch.float16, device=DEV)
        _gemm(a, w, bias, rows, K, N, row_cnt=cnt, act=act, out=out, bn=bn, cta_pair=pair)
        ref = a.float() @ w.float().T + bias
        if act:
            ref = F.gelu(ref)
        n = int(cnt.item())
        assert _rel(out[:n], ref[:n].cpu()) <= 1e-3
        assert (out[n:] == 7.0).all()                              # rows past the device-side count are never written
    assert cuda_lib.laud_tok_gemm_launch_count() > 0


@pytest.mark.parametrize("rows,K,N,R", [(700, 384, 384, 1500), (5300, 1536, 384, 9000), (3000, 768, 384, 3000)])
@pytest.mark.parametrize("pair", [0, 1])
def test_tok_gemm_residual_scatter_and_no_row_count(cuda_lib, rows, K, N, R, pair):
    g = torch.Generator().manual_seed(5)
    a = (torch.randn(rows, K, generator=g) * 0.5).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    dest = torch.randperm(R, generator=g)[:rows].to(torch.int32).to(DEV)
    x0 = torch.randn(R, N, generator=g).to(DEV)
    x = x0.clone()
    _gemm(a, w, bias, rows, K, N, resid=x, ldres=N, row_idx=dest, cta_pair=pair)
    ref = x0.clone()
    ref[dest.long()] += a.float() @ w.float().T + bias
    assert _rel(x, ref.cpu()) <= 2e-4
    untouched = torch.ones(R, dtype=torch.bool, device=DEV)
    untouched[dest.long()] = False
    assert torch.equal(x[untouched], x0[untouched])


@pytest.mark.parametrize("big", [False, True])      # big: enough rows for the weight-resident mode
def test_tok_gemm_head_tile_skip(cuda_lib, big):
    """col_gate: an n-tile is computed iff one of the samples of the m-tile's rows keeps that head."""
    g = torch.Generator().manual_seed(9)
    B, H, K = (12, 6, 384) if not big else (96, 6, 384)
    N = H * 192
    cnts = torch.randint(20, 150, (B,), generator=g)
    rows = int(cnts.sum())
    samp = torch.repeat_interleave(torch.arange(B), cnts).to(torch.int32).to(DEV)
    gate = (torch.rand(B, H, generator=g) < 0.4).to(torch.uint8).to(DEV)
    a = (torch.randn(rows, K, generator=g) * 0.5).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(DEV)
    out = torch.full((rows, N), 7.0, dtype=torch.float16, device=DEV)
    _gemm(a, w, None, rows, K, N, out=out, bn=192, col_gate=gate, gate_ld=H, row_sample=samp)
    ref = (a.float() @ w.float().T).cpu()
    out, samp_c, gate_c = out.cpu(), samp.cpu().long(), gate.cpu().bool()
    skipped = 0
    for m0 in range(0, rows, 128):
        m1 = min(m0 + 128, rows)
        ss = range(int(samp_c[m0]), int(samp_c[m1 - 1]) + 1)
        for h in range(H):
            blk = out[m0:m1, h * 192:(h + 1) * 192]
            if any(bool(gate_c[s, h]) for s in ss):
                assert _rel(blk, ref[m0:m1, h * 192:(h + 1) * 192]) <= 1e-3
            else:
                assert (blk == 7.0).all()
                skipped += 1
    assert skipped > 0


# the last three are large enough for the CTA-pair form (>= 2 m-tiles per pair of SMs), one with an odd number of m-tiles
@pytest.mark.parametrize("rows,D,Hd", [(300, 128, 512), (1000, 384, 1536), (5000, 384, 1536), (129, 256, 1024),
                                       (19369, 384, 1536), (40000, 128, 512), (25000, 256, 1024)])
def test_mlp_fused_vs_torch(cuda_lib, rows, D, Hd):
    """fc1 -> GELU -> fc2 -> residual add in one kernel (hidden activations on chip) against fp32 torch on the same fp16
    operands; rows past the device-side count and rows that are not destinations stay untouched."""
    g = torch.Generator().manual_seed(rows + D)
    y = (torch.randn(rows, D, generator=g) * 0.7).half().to(DEV)
    w1 = (torch.randn(Hd, D, generator=g) / D ** 0.5).half().to(DEV)
    w2 = (torch.randn(D, Hd, generator=g) / Hd ** 0.5).half().to(DEV)
    b1, b2 = (torch.randn(Hd, generator=g) * 0.3).to(DEV), (torch.randn(D, generator=g) * 0.3).to(DEV)
    R = rows + 77
    dest = torch.randperm(R, generator=g)[:rows].to(torch.int32).to(DEV)
    n = rows - 41 if rows > 200 else rows
    cnt = torch.tensor([n], dtype=torch.int32, device=DEV)
    x0 = torch.randn(R, D, generator=g).to(DEV)
    x = x0.clone()
    _lib.check(cuda_lib.laud_adavit_mlp_fused(y.data_ptr(), rows, D, Hd, cnt.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                              b2.data_ptr(), x.data_ptr(), D, dest.data_ptr(), _lib.stream_ptr()), "laud_adavit_mlp_fused")
    torch.cuda.synchronize()
    hid = F.gelu(y.float() @ w1.float().T + b1).half().float()          # the hidden activations are fp16 on chip too
    ref = x0.clone()
    ref[dest[:n].long()] += (hid @ w2.float().T + b2)[:n]
    assert _rel(x, ref.cpu()) <= 5e-4
    untouched = torch.ones(R, dtype=torch.bool, device=DEV)
    untouched[dest[:n].long()] = False
    assert torch.equal(x[untouched], x0[untouched])


def test_tok_gemm_rejects_bad_shapes(cuda_lib):
    a = torch.zeros(64, 100, dtype=torch.float16, device=DEV)
    with pytest.raises(_lib.LaudError):
        _gemm(a, a, None, 64, 100, 64, out=torch.zeros(64, 64, dtype=torch.float16, device=DEV))     # K % 64


# --------------------------------------------------------------------------- models
TINY = A.AdaViTCfg(img_size=64, embed_dim=128, depth=4, num_heads=2, num_classes=16)
SMALL = A.AdaViTCfg()       # DeiT-S


def _build(cfg, seed, batch, **rates):
    sd = synth.synth_adavit_state_dict(A.state_dict_shapes(cfg), seed)
    x = synth.synth_images(batch, cfg.img_size, seed + 2)
    calib = x.to(DEV) if cfg.embed_dim > 128 else x
    sd = synth.calibrate_adavit(sd, cfg.kwargs(), calib, **rates)
    m = AdaViT(**cfg.kwargs())
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd, x


def _forced(pol):
    return (pol.token.to(DEV), pol.head.to(DEV), pol.layer.to(DEV))


@pytest.mark.parametrize("cfg,batch", [(TINY, 5), (SMALL, 3)])
def test_policy_lists_and_gather_vs_oracle(cuda_lib, cfg, batch):
    m, sd, x = _build(cfg, 21, batch, token_rate=0.6, head_rate=0.5, layer_rate=0.6)
    traces = []
    with torch.no_grad():
        A.forward(sd, cfg, x, traces)
    lib, st = cuda_lib, _lib.stream_ptr()
    B, L, D, H = batch, cfg.seq_len, cfg.embed_dim, cfg.num_heads
    P, ws = m.prepare(), m.workspace(B)
    for i in (1, cfg.depth - 1):
        t, q = traces[i], P["blocks"][i]
        xin = t.x_in.to(DEV).contiguous()
        _lib.check(lib.laud_adavit_policy(xin.data_ptr(), B, L, D, H, A.LN_EPS, *[_lib.ptr(q[k]) for k in
                   ("n1_w", "n1_b", "ts_w", "ts_b", "np_w", "np_b", "ls_w", "ls_b", "hs_w", "hs_b")], *[ws[k][i].data_ptr() for k in
                   ("tok", "cnt", "head", "layer", "tok_lg", "head_lg", "layer_lg")], st))
        _lib.check(lib.laud_adavit_lists(ws["cnt"][i].data_ptr(), ws["layer"][i].data_ptr(), B, ws["off_a"][i].data_ptr(),
                                         ws["off_m"][i].data_ptr(), st))
        torch.cuda.synchronize()
        pol = t.policy
        for ours, lg_ours, want, lg in ((ws["tok"][i][:, 1:], ws["tok_lg"][i][:, 1:], pol.token[:, 1:], pol.token_logits),
                                        (ws["head"][i], ws["head_lg"][i], pol.head, pol.head_logits),
                                        (ws["layer"][i], ws["layer_lg"][i], pol.layer, pol.layer_logits)):
            assert _rel(lg_ours, lg) <= 1e-5
            diff = ours.bool().cpu() != want
            assert (lg.abs()[diff] <= POLICY_MARGIN * lg.abs().max()).all()
        assert ws["tok"][i][:, 0].all()
        tok, layer = ws["tok"][i].bool().cpu(), ws["layer"][i].bool().cpu()
        cnt = tok.sum(1)
        assert torch.equal(ws["cnt"][i].cpu().long(), cnt)
        for col, key in ((0, "off_a"), (1, "off_m")):
            want_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt * layer[:, col], 0)])
            assert torch.equal(ws[key][i].cpu().long(), want_off)
        # LayerNorm + gather of the kept tokens of the samples that run the attention sub-layer
        y = torch.full((B * L, D), 9.0, dtype=torch.float16, device=DEV)
        rows = torch.full((B * L,), -1, dtype=torch.int32, device=DEV)
        samp = torch.full((B * L,), -1, dtype=torch.int32, device=DEV)
        _lib.check(lib.laud_adavit_ln_gather(xin.data_ptr(), B, L, D, A.LN_EPS, q["n1_w"].data_ptr(), q["n1_b"].data_ptr(),
                                             ws["tok"][i].data_ptr(), ws["off_a"][i].data_ptr(), y.data_ptr(), rows.data_ptr(),
                                             samp.data_ptr(), st))
        torch.cuda.synchronize()
        want_rows = torch.tensor([b * L + l for b in range(B) if layer[b, 0] for l in range(L) if tok[b, l]], dtype=torch.int32)
        n = want_rows.numel()
        assert n == int(ws["off_a"][i][B])
        assert torch.equal(rows[:n].cpu(), want_rows) and (rows[n:] == -1).all()
        assert torch.equal(samp[:n].cpu(), want_rows // L)
        ln = F.layer_norm(t.x_in, (D,), sd[f"blocks.{i}.norm1.weight"], sd[f"blocks.{i}.norm1.bias"], A.LN_EPS).reshape(B * L, D)
        assert _rel(y[:n], ln[want_rows.long()]) <= 1e-3
        assert (y[n:] == 9.0).all()
        # the two-step form AdaViT.forward uses: row lists of both sub-layers, then LayerNorm over a row list - identical output
        rows_a = torch.full((B * L,), -1, dtype=torch.int32, device=DEV)
        rows_m, samp_a = rows_a.clone(), rows_a.clone()
        _lib.check(lib.laud_adavit_row_lists(ws["tok"][i].data_ptr(), B, L, ws["off_a"][i].data_ptr(), ws["off_m"][i].data_ptr(),
                                             rows_a.data_ptr(), samp_a.data_ptr(), rows_m.data_ptr(), st))
        y2 = torch.full((B * L, D), 9.0, dtype=torch.float16, device=DEV)
        _lib.check(lib.laud_adavit_ln_rows(xin.data_ptr(), D, A.LN_EPS, q["n1_w"].data_ptr(), q["n1_b"].data_ptr(), rows_a.data_ptr(),
                                           ws["off_a"][i][B:].data_ptr(), B * L, y2.data_ptr(), st))
        torch.cuda.synchronize()
        assert torch.equal(rows_a, rows) and torch.equal(samp_a, samp)
        assert torch.equal(y2, y)
        want_m = torch.tensor([b * L + l for b in range(B) if layer[b, 1] for l in range(L) if tok[b, l]], dtype=torch.int32)
        assert int(ws["off_m"][i][B]) == want_m.numel()
        assert torch.equal(rows_m[:want_m.numel()].cpu(), want_m) and (rows_m[want_m.numel():] == -1).all()


@pytest.mark.parametrize("B", [1, 33, 1024, 2500])
def test_lists_offsets_any_batch(cuda_lib, B):
    """ordered exclusive scans of the per-sample row counts, also past one block of 1024 samples"""
    g = torch.Generator().manual_seed(B)
    cnt = torch.randint(1, 198, (B,), generator=g).to(torch.int32)
    layer = (torch.rand(B, 2, generator=g) < 0.7).to(torch.uint8)
    off_a = torch.full((B + 1,), -1, dtype=torch.int32, device=DEV)
    off_m = off_a.clone()
    cnt_d, layer_d = cnt.to(DEV), layer.to(DEV)
    _lib.check(cuda_lib.laud_adavit_lists(cnt_d.data_ptr(), layer_d.data_ptr(), B, off_a.data_ptr(), off_m.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    for col, off in ((0, off_a), (1, off_m)):
        want = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt.long() * layer[:, col].long(), 0)])
        assert torch.equal(off.cpu().long(), want)


@pytest.mark.parametrize("H,L", [(2, 17), (6, 197), (3, 208)])
def test_attention_over_kept_tokens_vs_torch(cuda_lib, H, L):
    g = torch.Generator().manual_seed(H * 100 + L)
    B, D = 7, H * 64
    cnt = torch.randint(1, L + 1, (B,), generator=g)
    cnt[0], cnt[1], cnt[2] = L, 1, 0                              # full, class token only, sample skipped
    off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt, 0)]).to(torch.int32)
    rows = int(off[-1])
    head = (torch.rand(B, H, generator=g) < 0.6).to(torch.uint8)
    head[0] = 1
    qkv = (torch.randn(rows, 3 * D, generator=g)).half()
    o = torch.full((rows + 4, D), 5.0, dtype=torch.float16, device=DEV)
    qkv_d, off_d, head_d = qkv.to(DEV), off.to(DEV), head.to(DEV)
    _lib.check(cuda_lib.laud_adavit_attention(qkv_d.data_ptr(), 3 * D, off_d.data_ptr(), head_d.data_ptr(), B, H, L,
                                              o.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    o = o.cpu()
    assert (o[rows:] == 5.0).all()
    for b in range(B):
        r0, r1 = int(off[b]), int(off[b + 1])
        for h in range(H if r1 > r0 else 0):
            got = o[r0:r1, h * 64:(h + 1) * 64].float()
            if not head[b, h]:
                assert (got == 0).all()
                continue
            q, k, v = (qkv[r0:r1, h * 192 + j * 64: h * 192 + (j + 1) * 64].float() for j in range(3))
            want = torch.softmax(q @ k.T * 0.125, -1) @ v
            assert (got - want).abs().max() <= 2e-3 * max(1.0, want.abs().max().item())
    with pytest.raises(_lib.LaudError):
        _lib.check(cuda_lib.laud_adavit_attention(qkv_d.data_ptr(), 3 * D, off_d.data_ptr(), head_d.data_ptr(), B, H, 400,
                                                  o.data_ptr(), _lib.stream_ptr()))


@pytest.mark.parametrize("cfg,batch", [(TINY, 6), (SMALL, 4)])
def test_blocks_teacher_forced_vs_oracle(cuda_lib, cfg, batch):
    """Every block on the oracle's own input stream with the oracle's decisions installed: activations within 1e-3."""
    m, sd, x = _build(cfg, 31, batch, token_rate=0.55, head_rate=0.5, layer_rate=0.7)
    traces = []
    with torch.no_grad():
        A.forward(sd, cfg, x, traces)
    worst = 0.0
    for i, t in enumerate(traces):
        y = m.run_block(i, t.x_in.to(DEV), forced=_forced(t.policy))
        e = _rel(y, t.x_out)
        worst = max(worst, e)
        assert e <= ACT_TOL, (i, e)
        # dropped tokens / samples are bit-exactly untouched
        keep = t.policy.token & (t.policy.layer[:, :1] | t.policy.layer[:, 1:])
        assert torch.equal(y.cpu()[~keep], t.x_in[~keep])
    # head-tile skipping of the QKV projection changes nothing
    m.head_tile_skip = False
    t = traces[cfg.depth - 1]
    y0 = m.run_block(cfg.depth - 1, t.x_in.to(DEV), forced=_forced(t.policy))
    m.head_tile_skip = True
    y1 = m.run_block(cfg.depth - 1, t.x_in.to(DEV), forced=_forced(t.policy))
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("dim,heads", [(192, 3), (768, 12)])
def test_other_deit_widths_teacher_forced(cuda_lib, dim, heads):
    """DeiT-Ti (D = 192: scalar LayerNorm loads, three 64-channel chunks) and DeiT-B (D = 768: the QKV / proj / fc1 weight tiles
    no longer fit in shared memory -> streaming mode with head-tile skipping) through the same kernels, two blocks deep."""
    cfg = A.AdaViTCfg(embed_dim=dim, num_heads=heads, depth=3, num_classes=40)
    m, sd, x = _build(cfg, 61, 3, token_rate=0.5, head_rate=0.5, layer_rate=0.7)
    traces = []
    with torch.no_grad():
        want = A.forward(sd, cfg, x, traces)[0]
    for i, t in enumerate(traces):
        y = m.run_block(i, t.x_in.to(DEV), forced=_forced(t.policy))
        assert _rel(y, t.x_out) <= ACT_TOL, (i, _rel(y, t.x_out))
    lf, *_ = m(x.to(DEV), forced=[_forced(t.policy) for t in traces])
    assert _rel(lf, want) <= NET_TOL


def _compare_free_running(keeps, traces, B):
    """Decisions in execution order per sample up to its first differing block; a difference is accepted only inside the
    margin.  Returns (agree mask [B], flips)."""
    agree = torch.ones(B, dtype=torch.bool)
    flips = 0
    for k, t in zip(keeps, traces):
        pol = t.policy
        for ours, want, lg in ((k.token[:, 1:], pol.token[:, 1:], pol.token_logits), (k.head, pol.head, pol.head_logits),
                               (k.layer, pol.layer, pol.layer_logits)):
            if lg is None:
                assert ours.all()
                continue
            diff = (ours.cpu() != want).reshape(B, -1)
            lgr = lg.reshape(B, -1)
            for b in torch.nonzero(agree & diff.any(1)).flatten().tolist():
                assert (lgr[b][diff[b]].abs() <= FREE_MARGIN * lgr.abs().max()).all(), "decision differs outside the fp16 margin"
                flips += int(diff[b].sum())
        for b in range(B):
            if agree[b] and not (torch.equal(k.token[b].cpu(), pol.token[b]) and torch.equal(k.head[b].cpu(), pol.head[b])
                                 and torch.equal(k.layer[b].cpu(), pol.layer[b])):
                agree[b] = False
    return agree, flips


@pytest.mark.parametrize("cfg,batch", [(TINY, 8), (SMALL, 8)])
def test_network_free_running_vs_oracle_and_graph(cuda_lib, cfg, batch):
    m, sd, x = _build(cfg, 41, batch)
    traces, keeps = [], []
    with torch.no_grad():
        want, wt, wh, wl = A.forward(sd, cfg, x, traces)
    logits, tok, head, layer = m(x.to(DEV), keep=keeps)
    assert tuple(tok.shape) == (batch, cfg.depth, cfg.seq_len) and tok.dtype == torch.bool
    agree, flips = _compare_free_running(keeps, traces, batch)
    # a kept/dropped decision whose logit sits inside the fp16 operand noise may differ (each one was checked against the
    # margin above); after its first such flip a sample follows another trajectory and leaves the comparison.  ~2000 token
    # decisions per image at DeiT-S size make a flip per image likely, so only a floor on the agreeing samples is asserted;
    # the teacher-forced run below compares the logits of ALL samples.
    assert agree.sum() >= max(1, batch // 4) and flips <= 2 * batch
    assert _rel(logits[agree.to(DEV)], want[agree]) <= NET_TOL
    assert torch.equal(tok.cpu()[agree], wt[agree]) and torch.equal(head.cpu()[agree], wh[agree]) and torch.equal(layer.cpu()[agree], wl[agree])
    # decisions are not degenerate: something is skipped, something is kept
    dyn = tok[:, cfg.keep_layers:, 1:].float().mean().item()
    assert 0.3 < dyn < 0.9 and 0.3 < head[:, cfg.keep_layers:].float().mean().item() < 0.95
    # the CUDA-graph replay reproduces the eager forward bit-exactly
    g = m.capture(batch)
    lg2 = g.replay(x.to(DEV).half()).clone()
    torch.cuda.synchronize()
    assert torch.equal(lg2, logits)
    # teacher-forced whole network: identical decisions for every sample -> logits of all samples within the budget
    forced = [_forced(t.policy) for t in traces]
    lf, *_ = m(x.to(DEV), forced=forced)
    assert _rel(lf, want) <= NET_TOL


def test_skipping_is_executed_not_masked(cuda_lib):
    """Launch / work accounting: with everything dropped but the class token the GEMMs see one row per sample."""
    cfg = TINY
    m, sd, x = _build(cfg, 51, 4)
    B, L, H = 4, cfg.seq_len, cfg.num_heads
    tok = torch.zeros(B, L, dtype=torch.bool); tok[:, 0] = True
    forced = [(tok.to(DEV), torch.ones(B, H, dtype=torch.bool, device=DEV), torch.ones(B, 2, dtype=torch.bool, device=DEV))] * cfg.depth
    m(x.to(DEV), forced=forced)
    ws = m.workspace(B)
    assert int(ws["off_a"][-1][B]) == B and int(ws["off_m"][-1][B]) == B
    pol = [A.BlockPolicy(tok, torch.ones(B, H, dtype=torch.bool), torch.ones(B, 2, dtype=torch.bool))] * cfg.depth
    with torch.no_grad():
        want = A.forward(sd, cfg, x, forced=pol)[0]
    assert _rel(ws["logits"], want) <= NET_TOL


def test_empty_and_ragged_edge_cases(cuda_lib):
    """Edge cases of the compact lists: a block in which EVERY sample skips both sub-layers (zero rows: nothing is
    scheduled, the stream is untouched), a block in which only one sample runs, batch 1, and the all-kept path."""
    cfg = TINY
    m, sd, x = _build(cfg, 71, 5)
    B, L, H = 5, cfg.seq_len, cfg.num_heads
    traces = []
    with torch.no_grad():
        A.forward(sd, cfg, x, traces)
    x_in = traces[2].x_in
    ones_t, ones_h = torch.ones(B, L, dtype=torch.bool), torch.ones(B, H, dtype=torch.bool)
    # (1) nobody runs anything
    pol = A.BlockPolicy(ones_t, ones_h, torch.zeros(B, 2, dtype=torch.bool))
    y = m.run_block(2, x_in.to(DEV), forced=_forced(pol))
    assert torch.equal(y.cpu(), x_in)
    ws = m.workspace(B)
    assert int(ws["off_a"][2][B]) == 0 and int(ws["off_m"][2][B]) == 0
    # (2) only sample 3 runs, attention only, with a ragged token set and one head
    tok = torch.zeros(B, L, dtype=torch.bool); tok[:, 0] = True; tok[3, 1::3] = True
    head = torch.zeros(B, H, dtype=torch.bool); head[3, 1] = True
    layer = torch.zeros(B, 2, dtype=torch.bool); layer[3, 0] = True
    pol = A.BlockPolicy(tok, head, layer)
    y = m.run_block(2, x_in.to(DEV), forced=_forced(pol))
    with torch.no_grad():
        want = A.block_forward(x_in, sd, cfg, 2, pol)
    assert _rel(y, want) <= ACT_TOL
    keep = torch.zeros(B, L, dtype=torch.bool); keep[3] = tok[3]
    assert torch.equal(y.cpu()[~keep], x_in[~keep])
    # (3) batch 1, free-running, against the oracle
    with torch.no_grad():
        w1 = A.forward(sd, cfg, x[:1])[0]
    l1 = m(x[:1].to(DEV))[0]
    assert torch.isfinite(l1).all() and tuple(l1.shape) == tuple(w1.shape)    # (a single free-running sample may take another
    #                                                                           trajectory at a tie: parity is checked above)
    # (4) every gate open (policies off): the dense DeiT path
    cfg_d = A.AdaViTCfg(img_size=64, embed_dim=128, depth=3, num_heads=2, num_classes=16, ada_token=False, ada_head=False, ada_layer=False)
    md = AdaViT(**cfg_d.kwargs())
    sdd = synth.synth_adavit_state_dict(A.state_dict_shapes(cfg_d), 72)
    md.load_state_dict(sdd, strict=True)
    md = md.to(DEV).eval()
    with torch.no_grad():
        wd = A.forward(sdd, cfg_d, x)[0]
    ld, tk, hd, ly = md(x.to(DEV))
    assert tk.all() and hd.all() and ly.all()
    assert _rel(ld, wd) <= NET_TOL
