"""CPU-only checks of the host side: drop-in module tree / state_dict layout,
the C-ABI library (loads, exports every declared symbol - no compute calls),
error behaviour without a GPU, and the N>1 sharding logic over gloo."""
import ctypes
import hashlib
import json
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import laudnet_b200 as L
from laudnet_b200 import _lib, build as lbuild, dist as ldist, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_state_dict.json")))


# ------------------------------------------------------------------ module tree
@pytest.mark.parametrize("name", list(FIX))
def test_state_dict_layout_matches_reference(name):
    """Key order and shapes equal the reference model's (fixture generated from
    /root/reference by tests/golden/make_state_dict_fixture.py)."""
    f = FIX[name]
    ctor = L.uni_resnet50 if f["arch"] == "50" else L.uni_resnet101
    m = ctor(**f["kwargs"])
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert len(shapes) == f["n_keys"]
    assert hashlib.sha256(json.dumps(shapes, sort_keys=False).encode()).hexdigest() == f["ordered_keys_shapes_sha256"]
    if "keys" in f:
        assert shapes == f["keys"]
    pol = [[g["name"], len(g["params"]), g["lr_mult"]] for g in m.get_optim_policies()]
    assert pol == [list(p) for p in f["policies"]]


def test_state_dict_roundtrip_and_strict_load():
    kw = FIX["r50_spatial4421"]["kwargs"]
    a, b = L.uni_resnet50(**kw), L.uni_resnet50(**kw)
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ckpt = {"state_dict": a.state_dict()}          # reference checkpoint wrapping (train/main.py:306,488)
    b.load_state_dict(ckpt["state_dict"])


def test_masker_init_matches_reference_quirk():
    # keep-biased init with the reference's off-by-one (utils.py:42-43,107-108)
    m = L.Masker_channel_MLP(64, 8, layers=2, reduction=16)
    bias = m.conv[2].bias.detach()
    assert torch.all(bias[:8] == 2.0) and torch.all(bias[9:] == -2.0) and bias[8] not in (2.0, -2.0)
    s = L.Masker_spatial(64, 1, 7)
    assert s.conv.bias[0].item() == 5.0
    assert s.conv_flops_pp == 2 * 64 + 64


# ------------------------------------------------------------------ no-GPU behaviour
def test_cpu_inputs_raise_instead_of_falling_back():
    kw = FIX["r50_spatial4421"]["kwargs"]
    m = L.uni_resnet50(**kw).eval()
    with pytest.raises(L.LaudError):
        m(torch.zeros(1, 3, 224, 224), 1.0)
    m.train()
    with pytest.raises(L.LaudError):
        m(torch.zeros(1, 3, 224, 224), 1.0)
    with pytest.raises(L.LaudError):
        L.Masker_channel_MLP(16, 4).eval()(torch.zeros(1, 16, 4, 4), 1.0)
    with pytest.raises(L.LaudError):
        L.apply_channel_mask(torch.zeros(1, 4, 2, 2), torch.ones(1, 2))
    with pytest.raises(L.LaudError):
        L.uni_resnet50(pretrained=True)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "laudnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports the oracle"
                assert "/root/reference" not in src, f"{fn} reads the reference at run time"


# ------------------------------------------------------------------ C ABI
def _declared_symbols():
    import glob
    names = set()
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):      # laud_b200.h, laud_adavit.h
        hdr = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names.update(re.findall(r"\b(laud_[a-z0-9_]+)\s*\(", hdr))
    return sorted(names)


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert declared, "no symbols parsed from include/*.h"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_builds_and_exports_every_symbol():
    path = lbuild.build()                      # no-op when up to date; nvcc cross-compiles without a GPU
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} not exported"
    handle.laud_abi_version.restype = ctypes.c_int
    assert handle.laud_abi_version() == 4       # pure host call, no device needed (LAUD_ABI_VERSION)
    lib = _lib.lib()                            # binding sets argtypes for every symbol
    assert lib.laud_launch_count() == 0


def test_sass_contains_tcgen05_and_no_legacy_only_path():
    """The product conv kernel must be tcgen05 (UTC*MMA + LDTM in SASS)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lbuild.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass, "no tcgen05 MMA / TMEM load in the built library"


# ------------------------------------------------------------------ synthetic data
def test_synth_is_deterministic_and_shardable():
    a = synth.synth_images(4, 32, seed=3)
    b = synth.synth_images(2, 32, seed=3, start=2)
    assert torch.equal(a[2:], b)
    assert torch.equal(a, a.half().float())
    t = synth.synth_tensor("layer1.0.conv1.weight", (8, 4, 1, 1), 7)
    assert torch.equal(t, synth.synth_tensor("layer1.0.conv1.weight", (8, 4, 1, 1), 7))
    m = torch.tensor([[0.3, -1.0], [0.1, 2.0], [-0.2, 3.0], [0.5, -4.0]])
    d = synth.calibrate_two_way_bias(m, 0.5, per_group=True)
    assert ((m - d) >= 0).float().mean(0).tolist() == [0.5, 0.5]


# ------------------------------------------------------------------ sharding
def test_shard_range_partitions_the_batch():
    for batch in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [ldist.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == (ldist.max_shard(batch, world) if batch else 0)
    with pytest.raises(ValueError):
        ldist.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, batch, ncls, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(batch * ncls, dtype=torch.float32).view(batch, ncls)
        lo, hi = ldist.shard_range(batch, rank, world)
        got = ldist.allgather_logits(full[lo:hi].clone(), batch)
        counts = torch.full((3, 4), rank + 1, dtype=torch.int32)
        ldist.allreduce_counts(counts)

        class _M:                                   # stands in for the CUDA model: host logic only
            def forward_logits(self, x):
                return x * 2.0
        sc = ldist.ShardedClassifier(_M())
        got2 = sc(full[lo:hi].clone(), batch)
        q.put((rank, torch.equal(got, full), int(counts[0, 0]), torch.equal(got2, full * 2.0), sc.my_range(batch) == (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_allgather_logits_gloo_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, batch, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 3, True, True), (1, True, 3, True, True)]


def test_allgather_single_process_is_identity():
    x = torch.randn(3, 4)
    assert ldist.allgather_logits(x, 3) is x


def test_gap_partial_slot_arithmetic():
    """Layout contract of laud_conv_desc::gap_partial (include/laud_b200.h): the flat [B*HW] pixel list is cut into
    128-pixel tiles; slot k of sample b is tile (b*HW)//128 + k.  For every feature-map size of the supported networks
    the tiles that hold pixels of a sample are contiguous, fit in gap_tiles(HW) slots, a tile spans at most 4 samples
    (the kernel's segment limit, HW >= 43), and every pixel is counted exactly once."""
    from laudnet_b200._engine import gap_tiles
    import numpy as np
    for hw in (43, 49, 64, 100, 196, 784, 3136, 12544):
        K = gap_tiles(hw)
        assert K == (hw - 1) // 128 + 2
        B = 37
        tile_of = np.arange(B * hw) // 128
        sample_of = np.arange(B * hw) // hw
        counted = np.zeros(B, dtype=np.int64)
        for b in range(B):
            tiles = np.unique(tile_of[sample_of == b])
            first = (b * hw) // 128
            assert tiles[0] == first and tiles[-1] == ((b + 1) * hw - 1) // 128
            assert np.array_equal(tiles, np.arange(tiles[0], tiles[-1] + 1)) and len(tiles) <= K
            for t in tiles:                                   # what one pooling task adds to slot t - first
                counted[b] += int(np.sum((tile_of == t) & (sample_of == b)))
        assert np.all(counted == hw)
        spans = [len(np.unique(sample_of[tile_of == t])) for t in np.unique(tile_of)]
        assert max(spans) <= 4


def test_engine_decides_where_the_gap_is_fused(monkeypatch):
    """ResNetEngine._gap_fusable (host logic, no device): conv3 leaves the pool of its output for the next block's channel
    masker (or, in the last block, for the head) only when conv3 is a flat GEMM - masked-dense channel execution, no
    spatial / layer gate in the producing block - and the feature map is large enough for the kernel's segment limit."""
    from laudnet_b200 import _lib
    from laudnet_b200._engine import BlockPlan, ResNetEngine
    for var in ("LAUD_CONV_V3", "LAUD_NO_FLAT", "LAUD_NO_DMA"):
        monkeypatch.delenv(var, raising=False)

    class _MLP:                     # stands in for Masker_channel_MLP: pools the block input itself
        def gate_from_partials(self):
            pass

    class _ConvLinear:              # Masker_channel_conv_linear: pools relu(bn(conv1x1(x))) - needs the activations
        pass

    class _Blk:
        def __init__(self, mk):
            self.masker_channel = mk

    def plans(modes, hw=(56, 28, 14, 7), maskers=None):
        maskers = maskers or [_MLP()] * len(modes)
        return [BlockPlan(index=i, stage=i, inplanes=64, width=64, outplanes=256, stride=1, H_in=h, H_out=h, mode=m, gran=2, G=32,
                          g_spatial=1, mask_size=1, module=_Blk(mk if m in ("channel", "both") else None))
                for i, (m, h, mk) in enumerate(zip(modes, hw, maskers))]

    eng = ResNetEngine.__new__(ResNetEngine)
    eng.fuse_gap, eng.channel_exec, eng.impl = True, "dense", _lib.CONV_AUTO
    eng.plans = plans(["channel"] * 4)
    assert [eng._gap_fusable(p) for p in eng.plans] == [True, True, True, True]          # the last one feeds the head
    eng.plans = plans(["channel", "spatial", "channel", "layer"])
    assert [eng._gap_fusable(p) for p in eng.plans] == [False, False, False, False]      # next has no channel gate / own spatial gate
    eng.plans = plans(["channel"] * 4, hw=(56, 28, 14, 6))
    assert [eng._gap_fusable(p) for p in eng.plans] == [True, True, True, False]         # 36 pixels: a tile could span > 4 samples
    # the Bottleneck default masker (reference laud_resnet.py:36) cannot decide from the pool of x: ADVICE r1 (high)
    eng.plans = plans(["channel"] * 4, maskers=[_MLP(), _ConvLinear(), _MLP(), _ConvLinear()])
    assert [eng._gap_fusable(p) for p in eng.plans] == [False, True, False, True]
    assert [p.masker_kind for p in eng.plans] == ["MLP", "conv_linear", "MLP", "conv_linear"]
    eng.plans = plans(["channel"] * 4)
    eng.channel_exec = "sparse"
    assert not any(eng._gap_fusable(p) for p in eng.plans)                               # gathered conv3 is not a flat GEMM
    eng.channel_exec = "dense"
    monkeypatch.setenv("LAUD_NO_FLAT", "1")
    assert not any(eng._gap_fusable(p) for p in eng.plans)


# --------------------------------------------------------------------------- validate()-compatible caller
class _FakeLaud(torch.nn.Module):
    """Stand-in for a LAUD model on CPU: deterministic 7-tuple from the images (host logic of validate() only)."""

    def __init__(self, n_blocks=(2, 3, 1, 2), ncls=7):
        super().__init__()
        self.n_blocks, self.ncls = n_blocks, ncls
        g = torch.Generator().manual_seed(5)
        self.w = torch.randn(12, ncls, generator=g)

    def forward(self, images, temperature=1.0):
        feat = images.reshape(images.shape[0], -1)[:, :12]
        logits = feat @ self.w
        m = feat.mean()
        lists = [[torch.sigmoid(m * (k + 1) + torch.arange(n, dtype=torch.float32) * 0.1 * (j + 1)) for j, n in enumerate(self.n_blocks)]
                 for k in range(4)]
        perc = torch.sigmoid(m + torch.arange(sum(self.n_blocks), dtype=torch.float32) * 0.05)
        flops = (2.0e9 + 1e8 * m).reshape(())
        return (logits, *lists, perc, flops)


def _reference_validate_restated(batches, model, criterion, sparsity_criterion, args, epoch, world):
    """Literal restatement of the reference loop (train/main.py:627-757): per-batch all_reduce / world_size of every
    scalar, AverageMeter updates weighted by the local batch size, rank-local densities (the dead list-vs-string guard)."""
    from laudnet_b200.validate import AverageMeter, accuracy
    meters = {k: AverageMeter(k) for k in ("cls", "flops_loss", "loss", "act", "flops", "top1", "top5")}
    dens_sum, n = None, 0
    for images, target in batches:
        bs = images.size(0)
        n += bs
        out, r3, r2, r1, rc, perc, flops = model(images, temperature=args.t_last)
        d = torch.stack([torch.cat(l) for l in (r3, r2, r1, rc)]) * bs
        dens_sum = d if dens_sum is None else dens_sum + d
        flops = flops / 1e9
        vals = dict(cls=criterion(out, target), act=perc.mean(), flops=flops)
        vals["flops_loss"] = sparsity_criterion(epoch, perc, flops)
        vals["loss"] = vals["cls"] + args.lambda_act * vals["flops_loss"]
        vals["top1"], vals["top5"] = (a.reshape(()) for a in accuracy(out, target, topk=(1, 5)))
        for k, v in vals.items():
            v = v.clone().float()
            dist.all_reduce(v)
            meters[k].update((v / world).item(), bs)
    return (meters["top1"].avg, meters["top5"].avg, meters["loss"].avg, meters["act"].avg, meters["flops"].avg,
            (dens_sum / n).numpy())


def _validate_worker(rank, world, port, q):
    import types
    from laudnet_b200.validate import SparsityCriterion_bounds, validate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        batches = [(torch.randn(bs, 3, 4, 4, generator=g), torch.randint(0, 7, (bs,), generator=g)) for bs in (6, 6, 3)]
        model = _FakeLaud()
        args = types.SimpleNamespace(device="cpu", t_last=0.01, lambda_act=0.1, sparse=True, use_cuda_graph=False)
        crit = torch.nn.CrossEntropyLoss()
        sc = SparsityCriterion_bounds(0.5, 100, 4.1)
        got = validate(batches, model, crit, sc, args, epoch=40)
        want = _reference_validate_restated(batches, model, crit, sc, args, 40, world)
        ok = all(abs(a - b) <= 1e-5 * max(1.0, abs(b)) for a, b in zip(got[:5], want[:5]))
        ok_d = bool(np.allclose(got[5], want[5], rtol=1e-6, atol=1e-7)) and got[5].shape == (4, 8)
        red = validate(batches, model, crit, sc, args, epoch=40, reduce_density=True)[5]
        q.put((rank, ok, ok_d, red.tolist()))
    finally:
        dist.destroy_process_group()


def test_validate_matches_reference_loop_gloo_world2():
    """laudnet_b200.validate.validate (one all-reduce at the end) == the reference's per-batch all-reduce loop, on two
    gloo ranks with different data; reduce_density=True gives both ranks the same rank-averaged densities."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_validate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[:3] for r in res] == [(0, True, True), (1, True, True)]
    assert np.allclose(res[0][3], res[1][3])


def test_mmdet_adapter_interface_and_state_dict_keys():
    """laudnet_b200.LAD_MMDet_ResNet: the reference backbone's constructor keywords, its parameter names
    (the reference classifier's keys minus `fc.*`: build_norm_layer(postfix) names the norms bn1..3), and the refusals."""
    import json
    import laudnet_b200 as L
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_state_dict.json")))
    kw = dict(depth=101, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1, norm_cfg=dict(type="BN", requires_grad=True),
              norm_eval=True, style="pytorch", sparsity_target=0.5, temperature_0=1.0, temperature_t=0.01,
              spatial_mask_channel_group=[1, 1, 1, 1], mask_spatial_granularity=[1, 1, 1, 1], channel_dyn_granularity=[2, 2, 2, 2],
              dyn_mode=["channel"] * 4, channel_masker=["MLP"] * 4, channel_masker_layers=[2] * 4, reduction_ratio=[16] * 4)
    m = L.LAD_MMDet_ResNet(**kw)          # the backbone dict of the reference's faster_rcnn ... channel_2222 config (:14-20)
    keys = set(m.state_dict().keys())
    fx = ref["r101_channel2222"]                     # keys of the REFERENCE classifier with the same LAUD kwargs
    names = fx["shapes"].keys() if "shapes" in fx else fx["keys"]
    assert keys == {k for k in names if not k.startswith("fc.")}
    assert not any(k.startswith(("fc.", "_net.")) for k in keys) and "layer3.22.masker_channel.conv.2.bias" in keys
    assert m.norm1 is m.bn1
    m.train()
    assert not m.bn1.training and m.layer1[0].masker_channel.training          # norm_eval: BN frozen, gates train
    with pytest.raises(_lib.LaudError):
        L.LAD_MMDet_ResNet(depth=50, dyn_mode=["spatial"] * 4)                 # the reference backbone has no such masker
    with pytest.raises(_lib.LaudError):
        L.LAD_MMDet_ResNet(depth=50, dyn_mode=["layer"] * 4, deep_stem=True)
    with pytest.raises(_lib.LaudError):
        m(torch.zeros(1, 3, 800, 1344))                                         # rectangular inputs: not this round
