"""CPU tests of the DECLARED self-oracle of the AdaViT block (oracle/adavit_oracle.py) and of the host-side module layout.

The reference tree has no AdaViT code ("parity unpinned"), so the oracle is pinned by what CAN be
checked: the stock ViT arithmetic of an independent implementation with every gate open, and the equivalence of the
masked-dense forward with a really sparse one - the property the CUDA execution relies on."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.adavit_oracle as A                                   # noqa: E402
from laudnet_b200 import synth                                      # noqa: E402
from laudnet_b200.adavit import AdaViT                              # noqa: E402

TINY = A.AdaViTCfg(img_size=64, embed_dim=128, depth=4, num_heads=2, num_classes=16)


def _tiny(seed=3, batch=6, **rates):
    sd = synth.synth_adavit_state_dict(A.state_dict_shapes(TINY), seed)
    x = synth.synth_images(batch, TINY.img_size, seed + 2)
    return synth.calibrate_adavit(sd, TINY.kwargs(), x, **rates), x


def test_all_gates_open_equals_stock_vit_of_transformers():
    """With the policies switched off the oracle is a plain DeiT: compare with transformers' ViTForImageClassification
    (an independent implementation) on the same weights."""
    tf = pytest.importorskip("transformers")
    cfg = A.AdaViTCfg(img_size=64, embed_dim=128, depth=3, num_heads=2, num_classes=16, ada_token=False, ada_head=False, ada_layer=False)
    sd = synth.synth_adavit_state_dict(A.state_dict_shapes(cfg), 11)
    hf = tf.ViTForImageClassification(tf.ViTConfig(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=512,
                                                   image_size=64, patch_size=16, layer_norm_eps=A.LN_EPS, hidden_act="gelu", qkv_bias=True,
                                                   num_labels=16, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)).eval()
    m = {"vit.embeddings.cls_token": sd["cls_token"], "vit.embeddings.position_embeddings": sd["pos_embed"],
         "vit.embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
         "vit.embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
         "vit.layernorm.weight": sd["norm.weight"], "vit.layernorm.bias": sd["norm.bias"],
         "classifier.weight": sd["head.weight"], "classifier.bias": sd["head.bias"]}
    D = cfg.embed_dim
    for i in range(cfg.depth):
        p, q = f"blocks.{i}.", f"vit.encoder.layer.{i}."
        for j, nm in enumerate(("query", "key", "value")):
            m[q + f"attention.attention.{nm}.weight"] = sd[p + "attn.qkv.weight"][j * D:(j + 1) * D]
            m[q + f"attention.attention.{nm}.bias"] = sd[p + "attn.qkv.bias"][j * D:(j + 1) * D]
        for a, b in (("attention.output.dense", "attn.proj"), ("layernorm_before", "norm1"), ("layernorm_after", "norm2"),
                     ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2")):
            m[q + a + ".weight"], m[q + a + ".bias"] = sd[p + b + ".weight"], sd[p + b + ".bias"]
    missing, unexpected = hf.load_state_dict(m, strict=False)
    assert not unexpected and not [k for k in missing if "pooler" not in k], (missing, unexpected)
    x = synth.synth_images(3, 64, 12)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
        got, tok, head, layer = A.forward(sd, cfg, x)
    assert tok.all() and head.all() and layer.all()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_masked_dense_equals_really_sparse_block_by_block():
    sd, x = _tiny(token_rate=0.6, head_rate=0.5, layer_rate=0.7)
    traces = []
    with torch.no_grad():
        A.forward(sd, TINY, x, traces)
        for i, t in enumerate(traces):
            y = A.sparse_block_forward(t.x_in, sd, TINY, i, t.policy)
            assert (y - t.x_out).abs().max() <= 2e-5 * t.x_out.abs().max()
    pol = traces[-1].policy
    assert pol.token[:, 0].all()                                   # the class token is always kept
    assert 0.2 < torch.stack([t.policy.token[:, 1:].float().mean() for t in traces[TINY.keep_layers:]]).mean() < 0.9
    assert not all(bool(t.policy.head.all()) for t in traces[TINY.keep_layers:])
    assert traces[0].policy.token.all() and traces[0].policy.layer.all()   # static first block (keep_layers = 1)


def test_dropped_token_and_sample_are_identity():
    sd, x = _tiny()
    with torch.no_grad():
        xs = A.embed(sd, TINY, x)
        B, L, _ = xs.shape
        tok = torch.ones(B, L, dtype=torch.bool); tok[:, 3] = False
        layer = torch.ones(B, 2, dtype=torch.bool); layer[1] = False
        pol = A.BlockPolicy(tok, torch.ones(B, TINY.num_heads, dtype=torch.bool), layer)
        y = A.block_forward(xs, sd, TINY, 1, pol)
    assert torch.equal(y[:, 3], xs[:, 3]) and torch.equal(y[1], xs[1])
    assert not torch.equal(y[0, 2], xs[0, 2])


def test_flop_accounting_matches_the_dense_count_when_everything_is_kept():
    cfg = A.AdaViTCfg()
    B = 2
    tok = torch.ones(B, cfg.depth, cfg.seq_len, dtype=torch.bool)
    head = torch.ones(B, cfg.depth, cfg.num_heads, dtype=torch.bool)
    layer = torch.ones(B, cfg.depth, 2, dtype=torch.bool)
    assert torch.allclose(A.sparse_macs(cfg, tok, head, layer), torch.full((B,), A.dense_macs(cfg), dtype=torch.float64))
    assert 4.5e9 < A.dense_macs(cfg) < 4.7e9                      # DeiT-S: 4.6 GMACs
    tok[:, 3:, 100:] = False
    assert (A.sparse_macs(cfg, tok, head, layer) < A.dense_macs(cfg)).all()


def test_module_state_dict_layout_matches_the_oracle_and_loads():
    for cfg in (TINY, A.AdaViTCfg(img_size=64, embed_dim=128, depth=3, num_heads=2, num_classes=16, keep_layers=0, ada_head=False)):
        m = AdaViT(**cfg.kwargs())
        shapes = A.state_dict_shapes(cfg)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == shapes
        m.load_state_dict(synth.synth_adavit_state_dict(shapes, 1), strict=True)


def test_module_has_no_cpu_path():
    from laudnet_b200._lib import LaudError
    m = AdaViT(**TINY.kwargs()).eval()
    with pytest.raises(LaudError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(LaudError):
        AdaViT(img_size=64, embed_dim=96, depth=1, num_heads=2)    # head dimension 48
