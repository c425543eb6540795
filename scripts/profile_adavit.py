"""One eager forward of BASELINE configs[3] (AdaViT on DeiT-S, batch 512) between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
        --log-file gpurun_out/launches_c3.csv python scripts/profile_adavit.py
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from laudnet_b200 import synth                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
conf = bench.CONFIGS[3]
B = args.batch or conf["batch"]
dev = torch.device("cuda:0")
model, sd, kw = bench.build_adavit(conf, dev)
model = model.to(dev).eval()
x = synth.synth_images(B, 224, bench.SEED).to(torch.float16).to(dev)
with torch.no_grad():
    for _ in range(2):
        model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.forward_logits(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward of", conf["workload"], "batch", B)
