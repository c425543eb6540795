"""Per-layer device time of the conv launches of one eager forward of the bench workload
(CUDA events around every laud_conv_forward) + the CUDA-graphed step time.  For A/B runs of
kernel variants behind environment switches: `python scripts/layer_times.py [tag]`."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from laudnet_b200 import _engine, synth         # noqa: E402

B = int(os.environ.get("LAUD_PROFILE_BATCH", "256"))
dev = torch.device("cuda:0")
model, sd, _kw = bench.build_model(bench.CONFIGS[int(os.environ.get("LAUD_BENCH_CONFIG", "1"))], dev)
model = model.to(dev).eval()
x = synth.synth_images(B, 224, bench.SEED).to(torch.float16).to(dev)
runs = []
with torch.no_grad():
    for _ in range(2):
        model.forward_logits(x)
    torch.cuda.synchronize()
    for _ in range(5):
        with _engine.conv_profile() as prof:
            model.forward_logits(x)
            torch.cuda.synchronize()
        runs.append({k: t for k, (n, t) in prof.by_tag().items()})
    med = {k: statistics.median(r[k] for r in runs) for k in runs[0]}
    g = model.capture(x)
    for _ in range(3):
        g.run(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.run(x)
    e1.record()
    torch.cuda.synchronize()
tag = sys.argv[1] if len(sys.argv) > 1 else "run"
print(tag, "conv by layer (ms):", " ".join(f"{k}={v:.3f}" for k, v in sorted(med.items())))
print(tag, "conv total %.3f ms; graphed step %.3f ms = %.0f img/s" % (sum(med.values()), e0.elapsed_time(e1) / 20, B * 20e3 / e0.elapsed_time(e1)))
