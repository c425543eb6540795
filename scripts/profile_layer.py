"""One eager forward of configs[2] (layer skip) for ncu: LAUD_LAYER_EXEC=skip|mask."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laudnet_b200 as L
from laudnet_b200 import synth
dev = torch.device("cuda:0")
kw = dict(synth.HEADLINE_KWARGS); kw["dyn_mode"] = ["layer"] * 4; kw["mask_spatial_granularity"] = [56, 28, 14, 7]
model = L.uni_resnet101(**kw)
calib = synth.synth_images(32, 224, 101).to(dev)
model.load_state_dict(synth.synth_calibrated_state_dict(model, 1, calib, layer_rate=0.47))
model = model.to(dev).eval()
x = synth.synth_images(256, 224, 1).to(torch.float16).to(dev)
with torch.no_grad():
    for _ in range(2):
        model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
