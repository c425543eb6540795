"""Shared definition of the golden cases (used by tests/golden/make_golden.py,
which needs the reference, and by the tests, which do not)."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from laudnet_b200 import synth              # noqa: E402
from oracle import laud_oracle as O         # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (cfg, batch, seed)
    "tiny_channel": (O.ResNetCfg(layers=(2, 2, 2, 1), width_mult=0.25, input_size=64, num_classes=40,
                                 dyn_mode=("channel",) * 4, channel_dyn_granularity=(2, 2, 2, 2),
                                 channel_masker=("MLP",) * 4, channel_masker_layers=(2, 2, 2, 2)), 4, 11),
    "tiny_spatial": (O.ResNetCfg(layers=(2, 2, 2, 1), width_mult=0.25, input_size=64, num_classes=40,
                                 dyn_mode=("spatial",) * 4, mask_spatial_granularity=(4, 4, 2, 1)), 4, 12),
    "tiny_layer": (O.ResNetCfg(layers=(2, 2, 3, 2), width_mult=0.25, input_size=64, num_classes=40,
                               dyn_mode=("layer",) * 4, mask_spatial_granularity=(16, 8, 4, 2)), 6, 13),
    "tiny_both": (O.ResNetCfg(layers=(1, 2, 2, 1), width_mult=0.25, input_size=64, num_classes=40,
                              dyn_mode=("both", "both", "channel", "spatial"),
                              channel_dyn_granularity=(4, 2, 1, 2), spatial_mask_channel_group=(2, 1, 1, 2),
                              mask_spatial_granularity=(2, 2, 2, 1),
                              channel_masker=("MLP", "conv_linear", "MLP", "MLP"),
                              channel_masker_layers=(1, 2, 1, 2), reduction_ratio=(16, 2, 16, 16)), 3, 14),
}


# FULL-SIZE cases (224x224, the BASELINE.json architectures): name -> (kind, cfg, calibration batch, golden batch, seed).
# The golden images are synth_images(batch, 224, seed); calibration runs on synth_images(calib, 224, seed + 1000).
# Fixtures hold only what calibration changed in the state_dict plus the reference's outputs
# (tests/golden/make_golden_fullsize.py): the input and the seeded weights are regenerated from the seed.
_R50 = (3, 4, 6, 3)
FULL_CASES = {
    # BASELINE configs[1]: LAUD-ResNet101 channel-2222, MLP masker (2 layers, reduction 16)
    "full_r101_channel": ("resnet", O.ResNetCfg(), 8, 2, 31),
    # BASELINE configs[0]: LAUD-ResNet50 spatial 4-4-2-1
    "full_r50_spatial": ("resnet", O.ResNetCfg(layers=_R50, dyn_mode=("spatial",) * 4, channel_dyn_granularity=(1,) * 4,
                                               mask_spatial_granularity=(4, 4, 2, 1)), 8, 2, 32),
    # BASELINE configs[2]: LAUD-ResNet101 layer skip
    "full_r101_layer": ("resnet", O.ResNetCfg(dyn_mode=("layer",) * 4, channel_dyn_granularity=(1,) * 4,
                                              mask_spatial_granularity=(56, 28, 14, 7)), 8, 2, 33),
    # the reference Bottleneck's DEFAULT channel masker (laud_resnet.py:36): conv_linear, channel mode, ResNet-50
    "full_r50_convlinear": ("resnet", O.ResNetCfg(layers=_R50, channel_masker=("conv_linear",) * 4), 8, 2, 34),
    # BASELINE configs[4]: LAUD-RegNetY-800MF spatial 4-4-2-1 (target 0.3)
    "full_regnety800_spatial": ("regnet", O.RegNetCfg(), 8, 2, 35),
}


def full_case_state_dict(name, z=None):
    """-> (kind, cfg, state_dict, golden npz) of a full-size case: seeded weights + the calibrated tensors of the fixture."""
    kind, cfg, calib, batch, seed = FULL_CASES[name]
    if z is None:
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    if kind == "resnet":
        shapes = model_shapes(cfg)
    else:
        from laudnet_b200.laud_regnet import lad_regnet_y_800mf
        shapes = {k: tuple(v.shape) for k, v in lad_regnet_y_800mf(**cfg.kwargs()).state_dict().items()}
    sd = synth.synth_state_dict(shapes, int(z["seed"]))
    for k in z.files:
        if k.startswith("sd."):
            sd[k[3:]] = torch.from_numpy(z[k])
    assert state_dict_digest(sd) == z["sd_sha256"].tobytes(), f"{name}: regenerated state_dict differs from the fixture's"
    return kind, cfg, sd, z


def load_full_case(name):
    """-> (kind, cfg, state_dict, x[fp32, fp16-representable], golden npz dict)."""
    kind, cfg, sd, z = full_case_state_dict(name)
    _, _, calib, batch, seed = FULL_CASES[name]
    return kind, cfg, sd, synth.synth_images(batch, cfg.input_size, seed), z


# LAUD-RegNet-Y cases: (design-space parameters of BlockParams.from_init_params, cfg overrides, batch, seed).  The stage
# widths / depths the reference derives from the parameters are stored in the fixture and compared with
# laudnet_b200.laud_regnet.stage_params by the tests.
REGNET_CASES = {
    "tiny_regnet_spatial": (dict(depth=6, w_0=16, w_a=24.0, w_m=2.0, group_width=8),
                            dict(input_size=64, num_classes=24, stem_width=16, dyn_mode=("spatial",) * 4,
                                 mask_spatial_granularity=(4, 2, 2, 1)), 4, 21),
    "tiny_regnet_both": (dict(depth=6, w_0=16, w_a=24.0, w_m=2.0, group_width=8),
                         dict(input_size=64, num_classes=24, stem_width=16,
                              dyn_mode=("both", "channel", "spatial", "both"), channel_dyn_granularity=(2, 4, 1, 8),
                              spatial_mask_channel_group=(1, 1, 2, 1), mask_spatial_granularity=(2, 2, 2, 1),
                              channel_masker=("MLP", "conv_linear", "MLP", "MLP"), channel_masker_layers=(2, 2, 1, 1),
                              reduction_ratio=(16, 2, 16, 16)), 3, 22),
}


def regnet_cfg(name, widths, depths, group_widths):
    params, over, batch, seed = REGNET_CASES[name]
    return O.RegNetCfg(widths=tuple(widths), depths=tuple(depths), group_widths=tuple(group_widths),
                       strides=(2,) * len(widths), se_ratio=0.25, **over)


def regnet_model(name, cfg):
    from laudnet_b200.laud_regnet import BlockParams, LAD_RegNet
    params = REGNET_CASES[name][0]
    bp = BlockParams.from_init_params(se_ratio=0.25, **params)
    return LAD_RegNet(bp, **cfg.kwargs())


def load_regnet_case(name):
    """-> (cfg, state_dict, x, golden npz dict) for a LAUD-RegNet-Y fixture."""
    params, over, batch, seed = REGNET_CASES[name]
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = regnet_cfg(name, z["stage_widths"].tolist(), z["stage_depths"].tolist(), z["stage_group_widths"].tolist())
    m = regnet_model(name, cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.synth_state_dict(shapes, int(z["seed"]))
    for k in z.files:
        if k.startswith("sd."):
            sd[k[3:]] = torch.from_numpy(z[k])
    assert state_dict_digest(sd) == z["sd_sha256"].tobytes(), f"{name}: regenerated state_dict differs from the fixture's"
    x = torch.from_numpy(z["x"].astype(np.float32))
    return cfg, sd, x, z


def state_dict_digest(sd) -> bytes:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    return h.digest()


def model_shapes(cfg):
    """state_dict key -> shape of the network, taken from the drop-in module tree
    (identical to the reference's; tests/test_host_logic.py checks the key list)."""
    from laudnet_b200.laud_resnet import Bottleneck, ResNet
    m = ResNet(Bottleneck, list(cfg.layers), **cfg.kwargs())
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def load_case(name):
    """-> (cfg, state_dict, x[fp32, fp16-representable], golden npz dict)."""
    cfg, batch, seed = CASES[name]
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    sd = synth.synth_state_dict(model_shapes(cfg), int(z["seed"]))
    for k in z.files:
        if k.startswith("sd."):
            sd[k[3:]] = torch.from_numpy(z[k])
    assert state_dict_digest(sd) == z["sd_sha256"].tobytes(), f"{name}: regenerated state_dict differs from the fixture's"
    x = torch.from_numpy(z["x"].astype(np.float32))
    return cfg, sd, x, z
