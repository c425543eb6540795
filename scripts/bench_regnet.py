"""Throughput of BASELINE configs[4] per GPU (LAUD-RegNetY-800MF spatial-skip target-0.3, batch 256 = 2048 / 8 GPUs).
Not the driver's bench line.  Optional: `--profile` brackets one eager forward with cudaProfilerStart/Stop for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laudnet_b200 as L                       # noqa: E402
from laudnet_b200 import synth                  # noqa: E402
from oracle import laud_oracle as O             # noqa: E402  (geometry for the calibration only)

B = int(os.environ.get("LAUD_PROFILE_BATCH", "256"))
dev = torch.device("cuda:0")
kw = dict(input_size=224, dyn_mode=["spatial"] * 4, mask_spatial_granularity=[4, 4, 2, 1],
          spatial_mask_channel_group=[1] * 4, channel_dyn_granularity=[1] * 4, channel_masker=["MLP"] * 4,
          channel_masker_layers=[2] * 4, reduction_ratio=[16] * 4)
model = L.lad_regnet_y_800mf(**kw)
cfg = O.RegNetCfg(**{k: tuple(v) if isinstance(v, list) else v for k, v in kw.items()})
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
calib = synth.synth_images(16, 224, 101).to(dev)
sd = synth.calibrate_regnet(synth.synth_state_dict(shapes, 1), O.regnet_geometry(cfg), calib, 1, spatial_rate=0.22)
model.load_state_dict(sd)
model = model.to(dev).eval()
x = synth.synth_images(B, 224, 1).to(torch.float16).to(dev)
with torch.no_grad():
    out = model(x, 1.0)
    if "--profile" in sys.argv:
        model.forward_logits(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        model.forward_logits(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        sys.exit(0)
    g = model.capture(x)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
dense = 0.835e9
print(f"RegNetY-800MF spatial: {B / ms * 1e3:.0f} img/s per GPU, {ms:.2f} ms/step (batch {B}), conv3 density "
      f"{float(torch.cat(out[1]).mean()):.3f}, reference flops counter / dense {float(out[6]) / dense:.3f}, "
      f"{g.launches} launches/step")
