"""Per-role lap timers of the TMA-staged conv kernel (diagnostic build, -DLAUD_KPROF).

    python -m laudnet_b200.build --prof
    LAUD_LIB=laudnet_b200/lib/liblaud_b200_prof.so python scripts/kprof.py [block_index ...]

Runs the bench model eagerly; for the chosen bottlenecks prints, per conv launch, the average over CTAs of the
cycles each warp role spent in each phase (see KP_LAP sites in csrc/conv_tma.cu)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from laudnet_b200 import _engine, _lib, synth   # noqa: E402

SITES = {
    "tma": ["decode", "wait_empty", "issue"],
    "mma": ["decode", "wait_tempty", "wait_full", "issue", "fences", "commit"],
    "gather": ["decode", "tables", "wait_empty", "issue", "h1step"],
    "epi0": ["loop_exit", "tables", "wait_tfull", "wait_rfull", "compute", "store", "drain", "release"],
    "epi1": ["loop_exit", "tables", "wait_tfull", "wait_rfull", "compute", "store", "drain", "release"],
}

blocks = [int(v) for v in sys.argv[1:]] or [1, 8]
B = int(os.environ.get("LAUD_PROFILE_BATCH", "256"))
dev = torch.device("cuda:0")
model, sd, _kw = bench.build_model(bench.CONFIGS[int(os.environ.get("LAUD_BENCH_CONFIG", "1"))], dev)
model = model.to(dev).eval()
x = synth.synth_images(B, 224, bench.SEED).to(torch.float16).to(dev)
L = _lib.lib()
L.laud_debug_kprof.argtypes = [C.c_void_p, C.c_int]
L.laud_debug_kprof.restype = C.c_int
buf = np.zeros((160, 5, 8), dtype=np.int64)

orig = _engine.run_conv
state = {"blk": -1}


def traced(*args, **kw):
    tag = kw.get("tag", "conv")
    if state["blk"] not in blocks:
        return orig(*args, **kw)
    torch.cuda.synchronize()
    L.laud_debug_kprof(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(*args, **kw)
    e1.record()
    torch.cuda.synchronize()
    L.laud_debug_kprof(buf.ctypes.data, 0)
    us = e0.elapsed_time(e1) * 1e3
    act = buf[:148]
    print(f"--- blk{state['blk']} {tag}: {us:.1f} us")
    for r, (role, names) in enumerate(SITES.items()):
        tot = act[:, r, :].sum(axis=1).astype(float)
        if tot.max() == 0:
            continue
        parts = "  ".join(f"{n}={act[:, r, i].mean() / 1e3:.1f}k" for i, n in enumerate(names))
        print(f"    {role:7s} total={tot.mean() / 1e3:7.1f}k (max {tot.max() / 1e3:.1f}k)  {parts}")


_engine.run_conv = traced
eng = model._engine
orig_block = eng.run_block


def run_block(p, *a, **k):
    state["blk"] = p.index
    return orig_block(p, *a, **k)


eng.run_block = run_block
with torch.no_grad():
    _engine.run_conv = orig
    model.forward_logits(x)
    torch.cuda.synchronize()
    _engine.run_conv = traced
    model.forward_logits(x)
    torch.cuda.synchronize()
