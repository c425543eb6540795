"""Golden fixtures for LAUD-RegNet-Y (reference imagenet_classification/models/laud_regnet.py), the RegNet twin of
make_golden.py.  Build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_regnet.py

Builds the REFERENCE LAD_RegNet from the design-space parameters of tests/golden_cases.REGNET_CASES, fills it with
seeded synthetic weights calibrated on the case's own batch, runs the reference forward (eval, CPU, fp32) and stores
input, calibrated tensors, the 7-tuple, every block's masks and output statistics, and the stage parameters the
reference derived (BlockParams.from_init_params)."""
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
from tests.golden_cases import REGNET_CASES, regnet_cfg, state_dict_digest      # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from models.laud_regnet import BlockParams as RefBlockParams, LAD_RegNet as RefRegNet   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(name):
    params, over, batch, seed = REGNET_CASES[name]
    bp = RefBlockParams.from_init_params(se_ratio=0.25, **params)
    cfg = regnet_cfg(name, bp.widths, bp.depths, bp.group_widths)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefRegNet(bp, **cfg.kwargs()).eval()
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed)
    sd0 = {k: v.clone() for k, v in sd.items()}
    x = synth.synth_images(batch, cfg.input_size, seed)
    sd = synth.calibrate_regnet(sd, O.regnet_geometry(cfg), x, seed, channel_rate=0.6, spatial_rate=0.45)
    ref.load_state_dict(sd, strict=True)
    rec = {}
    hooks = []
    for s, stage in enumerate(ref.trunk_output):
        blocks = list(stage.children())
        for i, blk in enumerate(blocks):
            tag = f"block{s + 1}-{i}"
            if blk.f.masker_channel is not None:
                hooks.append(blk.f.masker_channel.register_forward_hook(
                    lambda mod, inp, out, tag=tag: rec.__setitem__(tag + ".channel_mask", out[0].numpy().astype(np.uint8))))
            if blk.f.masker_spatial is not None:
                hooks.append(blk.f.masker_spatial.register_forward_hook(
                    lambda mod, inp, out, tag=tag: rec.__setitem__(tag + ".spatial_mask", out[0].numpy().astype(np.uint8))))
            last = i == len(blocks) - 1

            def grab(mod, inp, out, tag=tag, last=last):
                o = out[0].detach().double()
                rec[tag + ".out_stats"] = np.array([o.mean().item(), o.abs().mean().item(), o.abs().max().item()])
                if last:
                    rec[tag + ".out"] = out[0].detach().numpy().copy()
            hooks.append(blk.register_forward_hook(grab))
    with torch.no_grad():
        out = ref(x, 1.0)
    for h in hooks:
        h.remove()
    logits, r3, r2, r1, rc, perc, flops = out
    save = {"x": x.numpy().astype(np.float16), "logits": logits.numpy(), "flops_perc": perc.numpy(),
            "flops": np.float32(flops.item()), "stage_widths": np.array(bp.widths), "stage_depths": np.array(bp.depths),
            "stage_group_widths": np.array(bp.group_widths)}
    for key, lst in (("rho3", r3), ("rho2", r2), ("rho1", r1), ("rhoc", rc)):
        for s in range(len(lst)):
            save[f"{key}.{s}"] = lst[s].numpy()
    for k, v in rec.items():
        save["ref." + k] = v
    for k in sorted(sd):
        if not torch.equal(sd[k], sd0[k]):
            save["sd." + k] = sd[k].numpy()
    save["sd_sha256"] = np.frombuffer(state_dict_digest(sd), dtype=np.uint8)
    save["seed"] = np.int64(seed)
    save["state_dict_keys"] = np.array(sorted(shapes))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **save)
    dens = {k: float(np.mean(v)) for k, v in rec.items() if k.endswith("mask")}
    print(f"{name}: widths {bp.widths} depths {bp.depths} groups {bp.group_widths}; {os.path.getsize(path) / 1e6:.2f} MB, "
          f"flops_perc mean {perc.mean().item():.3f}, mask densities {min(dens.values()):.2f}..{max(dens.values()):.2f}")


if __name__ == "__main__":
    for name in REGNET_CASES:
        run_case(name)
