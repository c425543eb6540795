"""Throughput of BASELINE configs[2] (LAUD-ResNet101 layer-skip target-0.5, batch 256) with the layer gate executed
as a real skip (active-sample work lists, in-place output) vs masked-dense.  Not the driver's bench line."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laudnet_b200 as L                       # noqa: E402
from laudnet_b200 import synth                  # noqa: E402

B = int(os.environ.get("LAUD_PROFILE_BATCH", "256"))
dev = torch.device("cuda:0")
kw = dict(synth.HEADLINE_KWARGS)
kw["dyn_mode"] = ["layer"] * 4
kw["mask_spatial_granularity"] = [56, 28, 14, 7]
for mode in ("skip", "mask"):
    os.environ["LAUD_LAYER_EXEC"] = mode
    model = L.uni_resnet101(**kw)
    calib = synth.synth_images(32, 224, 101).to(dev)
    sd = synth.synth_calibrated_state_dict(model, 1, calib, layer_rate=0.47)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    x = synth.synth_images(B, 224, 1).to(torch.float16).to(dev)
    with torch.no_grad():
        out = model(x, 1.0)
        rate = float(torch.cat(out[1]).mean())
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
    print(f"layer_exec={mode}: {B / ms * 1e3:.0f} img/s, {ms:.2f} ms/step, mean block activation rate {rate:.3f}, "
          f"flops ratio {float(out[6]) / 7.81e9:.3f}, logits checksum {float(g.logits.float().abs().sum()):.4e}")
