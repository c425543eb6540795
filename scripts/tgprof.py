"""Lap timers of the token GEMM (diagnostic build, -DLAUD_KPROF):

    python -m laudnet_b200.build --prof
    LAUD_LIB=laudnet_b200/lib/liblaud_b200_prof.so python scripts/tgprof.py

Runs the four GEMM shapes of a dynamic AdaViT-DeiT-S block at batch 512 (54k compact rows) and prints, averaged over the
CTAs, the kilo-cycles each warp role spent per phase."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laudnet_b200 import _lib                     # noqa: E402
from laudnet_b200.adavit import AdaViT            # noqa: E402

SITES = {"tma0": ["other", "wait_empty", "issue"], "mma": ["other", "wait_tempty", "wait_full", "issue+commit", "wait_bfull"],
         "epi": ["other", "wait_tfull", "ld+math+store"], "tma1": ["other", "wait_empty", "issue"]}
L = _lib.lib()
L.laud_debug_tgprof.argtypes = [C.c_void_p, C.c_int]
L.laud_debug_tgprof.restype = C.c_int
dev = "cuda:0"
rows = int(os.environ.get("ROWS", "54000"))
g = torch.Generator().manual_seed(0)
for name, K, N, mode in (("qkv", 384, 1152, "store"), ("proj", 384, 384, "resid"), ("fc1", 384, 1536, "gelu"), ("fc2", 1536, 384, "resid")):
    a = (torch.randn(rows, K, generator=g) * 0.5).half().to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.zeros(rows, N, dtype=torch.float16, device=dev)
    x = torch.zeros(rows, N, device=dev)
    idx = torch.arange(rows, dtype=torch.int32, device=dev)
    kw = dict(out=out, bn=192 if name == "qkv" else 0) if mode != "resid" else dict(resid=x, ldres=N, row_idx=idx)
    if mode == "gelu":
        kw["act"] = _lib.ACT_GELU
    buf = np.zeros((160, 4, 8), dtype=np.int64)
    for rep in range(3):
        torch.cuda.synchronize()
        L.laud_debug_tgprof(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        AdaViT._gemm(a, w, bias, rows, K, N, _lib.stream_ptr(), **kw)
        e1.record()
        torch.cuda.synchronize()
    L.laud_debug_tgprof(buf.ctypes.data, 0)
    us = e0.elapsed_time(e1) * 1e3
    tiles = ((rows + 127) // 128) * ((N + 191) // 192)
    print(f"--- {name}: rows {rows} K {K} N {N}: {us:.1f} us, {2e-6 * rows * K * N / us:.0f} TFLOP/s, ~{tiles / 144:.1f} tiles per CTA")
    act = buf[:148]
    for r, (role, names) in enumerate(SITES.items()):
        tot = act[:, r, :].sum(axis=1).astype(float)
        if tot.max() == 0:
            continue
        m = act[:, r, :].mean(axis=0) / 1e3
        print(f"    {role:5s} total={tot.mean() / 1e3:7.1f}k  " + "  ".join(f"{n}={m[i]:.1f}k" for i, n in enumerate(names)))
