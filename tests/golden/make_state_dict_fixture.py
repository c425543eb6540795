"""Generate tests/golden/ref_state_dict.json from the unmodified reference.

Build-container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_state_dict_fixture.py

For four constructor configurations it records the reference model's
`state_dict` key order + shapes (as a sha256, and in full for the headline
ResNet-101 channel-2222 config) and the parameter-group sizes returned by
`get_optim_policies()` (laud_resnet.py:365-401).  tests/test_host_logic.py
checks the drop-in module tree against it.
"""
import contextlib
import hashlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/imagenet_classification")
sys.dont_write_bytecode = True

with contextlib.redirect_stdout(io.StringIO()):
    import models  # noqa: F401,E402  (the reference)
    from models.laud_resnet import uni_resnet50 as r50, uni_resnet101 as r101  # noqa: E402

COMMON = dict(input_size=224, channel_masker_layers=[2] * 4, reduction_ratio=[16] * 4)
CONFIGS = {
    "r101_channel2222": ("101", dict(COMMON, dyn_mode=["channel"] * 4, channel_dyn_granularity=[2] * 4,
                                     channel_masker=["MLP"] * 4, spatial_mask_channel_group=[1] * 4,
                                     mask_spatial_granularity=[4, 4, 2, 1], lr_mult=1.0)),
    "r101_layer": ("101", dict(COMMON, dyn_mode=["layer"] * 4, channel_dyn_granularity=[1] * 4,
                               channel_masker=["MLP"] * 4, spatial_mask_channel_group=[1] * 4,
                               mask_spatial_granularity=[56, 28, 14, 7], lr_mult=1.0)),
    "r50_spatial4421": ("50", dict(COMMON, dyn_mode=["spatial"] * 4, channel_dyn_granularity=[1] * 4,
                                   channel_masker=["MLP"] * 4, spatial_mask_channel_group=[1] * 4,
                                   mask_spatial_granularity=[4, 4, 2, 1], lr_mult=1.0)),
    "r50_both_convlinear": ("50", dict(COMMON, dyn_mode=["both"] * 4, channel_dyn_granularity=[4, 2, 2, 1],
                                       channel_masker=["conv_linear"] * 4, spatial_mask_channel_group=[2, 1, 1, 1],
                                       mask_spatial_granularity=[4, 4, 2, 1], lr_mult=0.5)),
}


def digest(shapes: dict) -> str:
    return hashlib.sha256(json.dumps(shapes, sort_keys=False).encode()).hexdigest()


if __name__ == "__main__":
    out = {}
    for name, (arch, kw) in CONFIGS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            ref = (r50 if arch == "50" else r101)(**kw)
        shapes = {k: list(v.shape) for k, v in ref.state_dict().items()}
        out[name] = dict(arch=arch, kwargs=kw, n_keys=len(shapes), ordered_keys_shapes_sha256=digest(shapes),
                         policies=[(g["name"], len(g["params"]), g["lr_mult"]) for g in ref.get_optim_policies()])
        if name == "r101_channel2222":
            out[name]["keys"] = shapes
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_state_dict.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: v["n_keys"] for k, v in out.items()})
