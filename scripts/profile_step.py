"""One eager forward of the bench workload for ncu (run under
`ncu --profile-from-start off ...`): warm-up forwards, then ONE forward between
cudaProfilerStart/Stop with every bottleneck in an NVTX range `blk<i>`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from laudnet_b200 import synth                  # noqa: E402

B = int(os.environ.get("LAUD_PROFILE_BATCH", "256"))
dev = torch.device("cuda:0")
model, sd = bench.build_model(dev)
model = model.to(dev).eval()
x = synth.synth_images(B, 224, bench.SEED).to(torch.float16).to(dev)
os.environ["LAUD_NVTX"] = "1"
with torch.no_grad():
    for _ in range(2):
        model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward at batch", B)
