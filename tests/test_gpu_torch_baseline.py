"""The stock-PyTorch GPU baseline SURVEY.md section 8(d) asks for: the reference's masked-dense arithmetic
(oracle/laud_oracle.py, a restatement of laud_resnet.py:88-165,316-363 - the reference itself cannot travel to
the GPU box) executed by PyTorch/cuDNN on the same B200, fp16 under autocast and fp32, next to the engine's
CUDA-graphed forward on the same weights and images.  A measurement with a loose assertion; the numbers are
printed (pytest -s) and written to gpurun_out/torch_cuda_baseline.json when that directory exists."""
import json
import os

import pytest
import torch

import bench
from laudnet_b200 import synth
from oracle import laud_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _time(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_engine_beats_stock_pytorch_cuda_on_the_headline_workload(cuda_lib):
    B = 256
    model, sd = bench.build_model(torch.device(DEV))
    model = model.to(DEV).eval()
    x = synth.synth_images(B, 224, bench.SEED)
    cfg = O.ResNetCfg()
    sd32 = {k: v.to(DEV).float() if v.is_floating_point() else v.to(DEV) for k, v in sd.items()}
    x32 = x.to(DEV).float()
    out = {}
    with torch.no_grad():
        ms32 = _time(lambda: O.resnet_forward(sd32, cfg, x32), 3)
        with torch.autocast("cuda", dtype=torch.float16):
            ms16 = _time(lambda: O.resnet_forward(sd32, cfg, x32), 3)
        xh = x.to(torch.float16).to(DEV)
        g = model.capture(xh)
        ms_ours = _time(lambda: g.run(xh), 10)
        ref_logits = O.resnet_forward(sd32, cfg, x32)[0]
        ours = g.run(xh)[0].float()
    out = {"batch": B, "torch_cuda_fp32_img_s": B * 1e3 / ms32, "torch_cuda_fp16_autocast_img_s": B * 1e3 / ms16,
           "engine_img_s": B * 1e3 / ms_ours, "note": "stock PyTorch/cuDNN executing the reference's masked-dense arithmetic "
           "(oracle restatement) vs the CUDA-graphed engine, same weights and images, one B200"}
    print(json.dumps(out))
    if os.path.isdir("gpurun_out"):
        json.dump(out, open("gpurun_out/torch_cuda_baseline.json", "w"), indent=1)
    # same network: the two forwards agree (fp16 activations vs fp32: loose tolerance; gates may differ at the margin)
    assert torch.isfinite(ours).all() and ours.shape == ref_logits.shape
    assert ms_ours < ms16, "the engine must beat stock PyTorch fp16 on its own hot path"
