"""One eager forward of a BASELINE configuration between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_forward.py --config 1
    python scripts/ncu_launch_summary.py gpurun_out/launches.csv profiles/<name>_launch_summary.txt profiles/conv_traffic.json
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
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
conf = bench.CONFIGS[args.config]
B = args.batch or conf["batch"]
dev = torch.device("cuda:0")
model, sd, kw = bench.build_model(conf, dev)
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
