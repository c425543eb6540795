"""Full-size golden fixtures: pin `oracle/laud_oracle.py` to the reference at the REAL shapes of the BASELINE.json
architectures (224x224 input; ResNet-101 / ResNet-50 / RegNetY-800MF), not only at the width x0.25 / 64-pixel nets of
make_golden.py.  Build container only (imports the unmodified reference from /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fullsize.py

For every case of tests/golden_cases.FULL_CASES: build the REFERENCE model, fill it with seeded synthetic weights
(laudnet_b200.synth), calibrate BatchNorm statistics and gate biases on a separate seeded batch, run the reference
forward (eval, CPU, fp32) on the golden batch and store logits, every block's gating masks (bit-packed), the
sparsity lists, flops_perc, flops and per-block output statistics.  Inputs and seeded weights are regenerated from
the seed by the tests; only the tensors calibration changed travel (sd.*), with a digest of the whole state_dict."""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/imagenet_classification")
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

from laudnet_b200 import synth                      # noqa: E402
from oracle import laud_oracle as O                 # noqa: E402
from tests.golden_cases import FULL_CASES, state_dict_digest      # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import models                                   # noqa: E402,F401  (the reference)
    from models.laud_resnet import ResNet as RefResNet, Bottleneck as RefBottleneck   # noqa: E402
    from models import laud_regnet as ref_regnet    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RATES = dict(channel_rate=0.6, spatial_rate=0.4, layer_rate=0.47)     # SURVEY 8d "target-0.5" densities


def run_case(name):
    kind, cfg, calib, batch, seed = FULL_CASES[name]
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "resnet":
            ref = RefResNet(RefBottleneck, list(cfg.layers), **cfg.kwargs()).eval()
            blocks = [(f"layer{s + 1}.{i}", blk, blk) for s in range(4) for i, blk in enumerate(getattr(ref, f"layer{s + 1}"))]
        else:
            ref = ref_regnet.lad_regnet_y_800mf(**cfg.kwargs()).eval()
            blocks = [(f"block{s + 1}-{i}", blk, blk.f) for s, stage in enumerate(ref.trunk_output)
                      for i, blk in enumerate(stage.children())]
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed)
    sd0 = {k: v.clone() for k, v in sd.items()}
    xc = synth.synth_images(calib, cfg.input_size, seed + 1000)
    if kind == "resnet":
        sd = synth.calibrate_resnet(sd, O.resnet_geometry(cfg), xc, seed, **RATES)
    else:
        sd = synth.calibrate_regnet(sd, O.regnet_geometry(cfg), xc, seed, channel_rate=0.6, spatial_rate=0.3)
    ref.load_state_dict(sd, strict=True)
    x = synth.synth_images(batch, cfg.input_size, seed)
    rec, hooks = {}, []
    for tag, blk, holder in blocks:
        if holder.masker_channel is not None:
            hooks.append(holder.masker_channel.register_forward_hook(
                lambda mod, inp, out, tag=tag: rec.__setitem__(tag + ".channel_mask", out[0].numpy().astype(np.uint8))))
        if holder.masker_spatial is not None:
            hooks.append(holder.masker_spatial.register_forward_hook(
                lambda mod, inp, out, tag=tag: rec.__setitem__(tag + ".spatial_mask", out[0].numpy().astype(np.uint8))))

        def grab(mod, inp, out, tag=tag):
            o = out[0].detach().double()
            rec[tag + ".out_stats"] = np.array([o.mean().item(), o.abs().mean().item(), o.abs().max().item()])
        hooks.append(blk.register_forward_hook(grab))
    with torch.no_grad():
        out = ref(x, 1.0)
    for h in hooks:
        h.remove()
    logits, r3, r2, r1, rc, perc, flops = out
    save = {"logits": logits.numpy(), "flops_perc": perc.numpy(), "flops": np.float32(flops.item()),
            "x_sha256": np.frombuffer(state_dict_digest({"x": x}), dtype=np.uint8)}
    for key, lst in (("rho3", r3), ("rho2", r2), ("rho1", r1), ("rhoc", rc)):
        for s in range(len(lst)):
            save[f"{key}.{s}"] = lst[s].numpy()
    for k, v in rec.items():
        if k.endswith("mask"):
            save["ref." + k + ".shape"] = np.array(v.shape)
            save["ref." + k + ".bits"] = np.packbits(v.reshape(-1))
        else:
            save["ref." + k] = v
    for k in sorted(sd):
        if not torch.equal(sd[k], sd0[k]):
            save["sd." + k] = sd[k].numpy()
    save["sd_sha256"] = np.frombuffer(state_dict_digest(sd), dtype=np.uint8)
    save["seed"] = np.int64(seed)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **save)
    dens = {k: float(np.mean(v)) for k, v in rec.items() if k.endswith("mask")}
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, flops_perc mean {perc.mean().item():.3f}, "
          f"mask densities {min(dens.values()):.2f}..{max(dens.values()):.2f}, |logits| max {logits.abs().max().item():.3f}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or FULL_CASES):
        run_case(name)
