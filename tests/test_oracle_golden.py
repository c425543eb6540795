"""The oracle (oracle/laud_oracle.py) against outputs of the unmodified
reference, committed as tests/golden/*.npz by tests/golden/make_golden.py."""
import os

import numpy as np
import pytest
import torch

from oracle import laud_oracle as O
from tests.golden_cases import CASES, GOLDEN_DIR, load_case


@pytest.mark.parametrize("name", list(CASES))
def test_network_matches_reference(name):
    cfg, sd, x, z = load_case(name)
    traces = []
    with torch.no_grad():
        logits, r3, r2, r1, rc, perc, flops = O.resnet_forward(sd, cfg, x, traces)
    # same fp32 ops in the same order: agreement to float rounding
    np.testing.assert_allclose(logits.numpy(), z["logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(perc.numpy(), z["flops_perc"], rtol=1e-6)
    np.testing.assert_allclose(flops.item(), z["flops"], rtol=1e-6)
    for key, lst in (("rho3", r3), ("rho2", r2), ("rho1", r1), ("rhoc", rc)):
        for s in range(4):
            np.testing.assert_array_equal(lst[s].numpy(), z[f"{key}.{s}"])
    geoms = O.resnet_geometry(cfg)
    assert len(geoms) == len(traces)
    for g, tr in zip(geoms, traces):
        tag = "ref." + g.prefix[:-1]
        if tr.channel_mask is not None:      # gating decisions are bit-exact
            np.testing.assert_array_equal(tr.channel_mask.numpy().astype(np.uint8), z[tag + ".channel_mask"])
        if tr.spatial_mask_small is not None:
            np.testing.assert_array_equal(tr.spatial_mask_small.numpy().astype(np.uint8), z[tag + ".spatial_mask"])
        o = tr.out.double()
        stats = np.array([o.mean().item(), o.abs().mean().item(), o.abs().max().item()])
        np.testing.assert_allclose(stats, z[tag + ".out_stats"], rtol=1e-5)
        if tag + ".out" in z.files:
            np.testing.assert_allclose(tr.out.numpy(), z[tag + ".out"], rtol=1e-5, atol=1e-5)


def test_dense_flops_counter():
    # static ResNet-50/101 MAC counts quoted by the reference's figure (4.1 / 7.8 GMAC)
    assert abs(O.dense_flops(O.ResNetCfg(layers=(3, 4, 6, 3))) / 1e9 - 4.1) < 0.05
    assert abs(O.dense_flops(O.ResNetCfg(layers=(3, 4, 23, 3))) / 1e9 - 7.8) < 0.05


class TestOperatorKATs:
    z = np.load(os.path.join(GOLDEN_DIR, "kat.npz"))

    def test_expand_mask(self):
        i = 0
        while f"expand.{i}.in" in self.z.files:
            st, pad, g = (int(v) for v in self.z[f"expand.{i}.cfg"])
            out = O.expand_mask(torch.from_numpy(self.z[f"expand.{i}.in"].astype(np.float32)), st, pad)
            np.testing.assert_array_equal(out.numpy().astype(np.uint8), self.z[f"expand.{i}.out"])
            i += 1
        assert i == 6

    def test_expand_mask_single_pixel(self):
        # one pixel at (1,1) of 3x3, stride 2 pad 1 -> ones at rows/cols 1..3 of 6x6
        m = torch.zeros(1, 1, 3, 3)
        m[0, 0, 1, 1] = 1
        out = O.expand_mask(m, 2, 1)[0, 0]
        want = torch.zeros(6, 6, dtype=torch.bool)
        want[1:4, 1:4] = True
        assert torch.equal(out, want)
        assert torch.equal(O.expand_mask(m, 1, 0), m > 0.5)

    def test_apply_masks(self):
        x = torch.from_numpy(self.z["acm.x"])
        out = O.apply_channel_mask(x, torch.from_numpy(self.z["acm.mask"]))
        np.testing.assert_array_equal(out.numpy(), self.z["acm.out"])
        out = O.apply_spatial_mask(x, torch.from_numpy(self.z["asm.mask"]))
        np.testing.assert_array_equal(out.numpy(), self.z["asm.out"])

    def test_nearest_resize(self):
        i = 0
        while f"resize.{i}.in" in self.z.files:
            want = self.z[f"resize.{i}.out"]
            out = O.nearest_resize(torch.from_numpy(self.z[f"resize.{i}.in"].astype(np.float32)), want.shape[-1])
            np.testing.assert_array_equal(out.numpy().astype(np.uint8), want)
            i += 1
        assert i == 5

    def _sd(self, tag):
        pre = f"masker.{tag}.sd."
        return {k[len(pre):]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith(pre)}

    def test_maskers(self):
        x = torch.from_numpy(self.z["masker.x"])
        sd = self._sd("spatial")
        mask, rho, flops, _ = O.masker_spatial(x, sd["conv.weight"], sd["conv.bias"], 4)
        self._cmp("spatial", mask, rho, flops)
        self._cmp("mlp2", *O.masker_channel_mlp(x, self._sd("mlp2"), "", 2)[:3])
        self._cmp("mlp1", *O.masker_channel_mlp(x, self._sd("mlp1"), "", 1)[:3])
        self._cmp("convlin", *O.masker_channel_conv_linear(x, self._sd("convlin"), "")[:3])

    def _cmp(self, tag, mask, rho, flops):
        np.testing.assert_array_equal(mask.numpy().astype(np.uint8), self.z[f"masker.{tag}.mask"])
        assert abs(rho.item() - float(self.z[f"masker.{tag}.sparsity"])) < 1e-7
        assert int(flops) == int(self.z[f"masker.{tag}.flops"])
        assert 0 < mask.mean().item() < 1, "degenerate KAT: the gate must be data dependent"


def test_masked_channel_constant_and_layer_skip_identity():
    """a masked channel after BN+ReLU is relu(beta - gamma*mu/sqrt(var+eps));
    a layer-skipped sample leaves the block as relu(identity) bit-exactly."""
    cfg, sd, x, _ = load_case("tiny_channel")
    g = O.resnet_geometry(cfg)[1]
    feat, _ = O.stem_forward(x, sd)
    feat = O.bottleneck_forward(feat, sd, O.resnet_geometry(cfg)[0])[0]
    tr = O.BlockTrace()
    O.bottleneck_forward(feat, sd, g, tr)
    p = g.prefix + "bn1."
    const = torch.relu(sd[p + "bias"] - sd[p + "weight"] * sd[p + "running_mean"] / torch.sqrt(sd[p + "running_var"] + 1e-5))
    per_ch = tr.channel_mask.repeat_interleave(g.width // g.groups_channel, dim=1)
    b, k = (per_ch == 0).nonzero()[0].tolist()
    assert torch.allclose(tr.a1[b, k], const[k].expand_as(tr.a1[b, k]), atol=1e-6)

    cfg, sd, x, _ = load_case("tiny_layer")
    g = O.resnet_geometry(cfg)[1]          # a non-downsample block
    feat, _ = O.stem_forward(x, sd)
    feat = O.bottleneck_forward(feat, sd, O.resnet_geometry(cfg)[0])[0]
    tr = O.BlockTrace()
    O.bottleneck_forward(feat, sd, g, tr)
    skipped = (tr.spatial_mask_small.view(-1) == 0).nonzero().view(-1)
    assert len(skipped) > 0
    assert torch.equal(tr.out[skipped], torch.relu(feat[skipped]))


# --------------------------------------------------------------------------- LAUD-RegNet-Y (laud_regnet.py)
from tests.golden_cases import REGNET_CASES, load_regnet_case, regnet_model     # noqa: E402


@pytest.mark.parametrize("name", list(REGNET_CASES))
def test_regnet_matches_reference(name):
    cfg, sd, x, z = load_regnet_case(name)
    traces = []
    with torch.no_grad():
        logits, r3, r2, r1, rc, perc, flops = O.regnet_forward(sd, cfg, x, traces)
    np.testing.assert_allclose(logits.numpy(), z["logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(perc.numpy(), z["flops_perc"], rtol=1e-6)
    np.testing.assert_allclose(flops.item(), z["flops"], rtol=1e-6)
    for key, lst in (("rho3", r3), ("rho2", r2), ("rho1", r1), ("rhoc", rc)):
        for s in range(4):
            np.testing.assert_array_equal(lst[s].numpy(), z[f"{key}.{s}"])
    geoms = O.regnet_geometry(cfg)
    assert len(geoms) == len(traces)
    for g, tr in zip(geoms, traces):
        tag = "ref." + g.prefix.split(".")[2]
        if tr.channel_mask is not None:
            np.testing.assert_array_equal(tr.channel_mask.numpy().astype(np.uint8), z[tag + ".channel_mask"])
        if tr.spatial_mask_small is not None:
            np.testing.assert_array_equal(tr.spatial_mask_small.numpy().astype(np.uint8), z[tag + ".spatial_mask"])
        o = tr.out.double()
        stats = np.array([o.mean().item(), o.abs().mean().item(), o.abs().max().item()])
        np.testing.assert_allclose(stats, z[tag + ".out_stats"], rtol=1e-5)
        if tag + ".out" in z.files:
            np.testing.assert_allclose(tr.out.numpy(), z[tag + ".out"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", list(REGNET_CASES))
def test_regnet_stage_params_and_state_dict_layout(name):
    """The drop-in's design-space derivation and module tree against what the reference produced."""
    from laudnet_b200.laud_regnet import stage_params
    cfg, sd, x, z = load_regnet_case(name)
    widths, depths, gws = stage_params(**REGNET_CASES[name][0])
    assert widths == z["stage_widths"].tolist() and depths == z["stage_depths"].tolist()
    assert gws == z["stage_group_widths"].tolist()
    m = regnet_model(name, cfg)
    assert sorted(m.state_dict().keys()) == z["state_dict_keys"].tolist()
    m.load_state_dict(sd, strict=True)


def test_regnet_y_800mf_geometry():
    """SURVEY appendix B.2 (verified there against the instantiated reference model)."""
    from laudnet_b200.laud_regnet import stage_params
    assert stage_params(depth=14, w_0=56, w_a=38.84, w_m=2.4, group_width=16) == ([64, 144, 320, 784], [1, 3, 8, 2], [16] * 4)
    assert stage_params(depth=16, w_0=48, w_a=27.89, w_m=2.09, group_width=8) == ([48, 104, 208, 440], [1, 3, 6, 6], [8] * 4)


# --------------------------------------------------------------------------- full-size pins (224x224, BASELINE architectures)
from tests.golden_cases import FULL_CASES, load_full_case       # noqa: E402


def unpack_mask(z, key):
    shape = tuple(int(v) for v in z[key + ".shape"])
    return np.unpackbits(z[key + ".bits"])[:int(np.prod(shape))].reshape(shape)


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_full_size_network_matches_reference(name):
    """The oracle against the reference ITSELF at the real shapes of the BASELINE.json architectures (ResNet-101
    channel-2222 / layer, ResNet-50 spatial / conv_linear, RegNetY-800MF spatial; 2 images, 224x224): logits, every
    gating decision of every block, the sparsity lists, flops_perc and flops (tests/golden/make_golden_fullsize.py)."""
    kind, cfg, sd, x, z = load_full_case(name)
    traces = []
    with torch.no_grad():
        if kind == "resnet":
            out = O.resnet_forward(sd, cfg, x, traces)
            tags = ["ref." + g.prefix[:-1] for g in O.resnet_geometry(cfg)]
        else:
            out = O.regnet_forward(sd, cfg, x, traces)
            tags = ["ref." + g.prefix.split(".")[2] for g in O.regnet_geometry(cfg)]
    logits, r3, r2, r1, rc, perc, flops = out
    np.testing.assert_allclose(logits.numpy(), z["logits"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(perc.numpy(), z["flops_perc"], rtol=1e-6)
    np.testing.assert_allclose(flops.item(), z["flops"], rtol=1e-6)
    for key, lst in (("rho3", r3), ("rho2", r2), ("rho1", r1), ("rhoc", rc)):
        for s in range(4):
            np.testing.assert_array_equal(lst[s].numpy(), z[f"{key}.{s}"])
    assert len(tags) == len(traces)
    n_gates = 0
    for tag, tr in zip(tags, traces):
        if tr.channel_mask is not None:
            np.testing.assert_array_equal(tr.channel_mask.numpy().astype(np.uint8), unpack_mask(z, tag + ".channel_mask"))
            n_gates += tr.channel_mask.numel()
        if tr.spatial_mask_small is not None:
            np.testing.assert_array_equal(tr.spatial_mask_small.numpy().astype(np.uint8), unpack_mask(z, tag + ".spatial_mask"))
            n_gates += tr.spatial_mask_small.numel()
        o = tr.out.double()
        stats = np.array([o.mean().item(), o.abs().mean().item(), o.abs().max().item()])
        np.testing.assert_allclose(stats, z[tag + ".out_stats"], rtol=1e-4)
    assert n_gates > 0


# --------------------------------------------------------------------------- training-mode gate with supplied Gumbel noise
def gumbel_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "kat_gumbel.npz"))
    for tag in ("spatial", "layer", "mlp2", "mlp1", "convlin"):
        sd = {k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + ".sd.")}
        yield tag, z, sd


@pytest.mark.parametrize("tag", ["spatial", "layer", "mlp2", "mlp1", "convlin"])
def test_gumbel_gate_with_supplied_noise_matches_reference(tag):
    """The reference maskers in TRAIN mode (F.gumbel_softmax, hard) with the generator seeded so that the noise they drew
    is known (tests/golden/make_golden_gumbel.py): the oracle, given that noise, reproduces the hard masks bit for bit -
    and they differ from the eval-mode masks (the noise matters)."""
    for t, z, sd in gumbel_cases():
        if t != tag:
            continue
        x, noise, tau = torch.from_numpy(z[f"{t}.x"]), torch.from_numpy(z[f"{t}.noise"]), float(z[f"{t}.tau"])
        if t in ("spatial", "layer"):
            S = z[f"{t}.mask"].shape[-1]
            mask, rho, _, _ = O.masker_spatial(x, sd["conv.weight"], sd["conv.bias"], S, noise, tau)
            ev = O.masker_spatial(x, sd["conv.weight"], sd["conv.bias"], S)[0]
        elif t == "convlin":
            mask, rho, _, _ = O.masker_channel_conv_linear(x, sd, "", noise, tau)
            ev = O.masker_channel_conv_linear(x, sd, "")[0]
        else:
            layers = 2 if t == "mlp2" else 1
            mask, rho, _, _ = O.masker_channel_mlp(x, sd, "", layers, noise, tau)
            ev = O.masker_channel_mlp(x, sd, "", layers)[0]
        np.testing.assert_array_equal(mask.numpy().astype(np.uint8), z[f"{t}.mask"])
        np.testing.assert_array_equal(ev.numpy().astype(np.uint8), z[f"{t}.eval_mask"])
        assert abs(float(rho) - float(z[f"{t}.sparsity"])) < 1e-6
        assert (z[f"{t}.mask"] != z[f"{t}.eval_mask"]).any()
