"""Microbenchmark of the fused MLP kernel (laud_adavit_mlp_fused): rows = 148 x 128 x passes, or 128 x tiles.

    python scripts/fm_bench.py [passes [tiles]]

Switches read by the library: LAUD_FM_PAIR=0 (single CTAs instead of CTA pairs), LAUD_FM_STAGGER=<cycles per chunk> (0 = off).
Timing experiments (WRONG results) need the diagnostic build - `python -m laudnet_b200.build --prof`, then
LAUD_LIB=laudnet_b200/lib/liblaud_b200_prof.so LAUD_FM_DBG=<bits>: 1 no GELU arithmetic, 2 no weight loads, 4 no residual
reductions, 8 no GELU epilogue body (profiles/README.md has the table these produced)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laudnet_b200 import _lib

def run(D=384, Hd=1536, passes=3, iters=20, tiles=None):
    L = _lib.lib()
    dev = torch.device("cuda:0")
    rows = 128 * tiles if tiles else 148 * 128 * passes
    g = torch.Generator(device="cpu").manual_seed(1)
    y = (torch.randn(rows, D, generator=g) * 0.5).half().to(dev)
    w1 = (torch.randn(Hd, D, generator=g) * 0.05).half().to(dev)
    w2 = (torch.randn(D, Hd, generator=g) * 0.03).half().to(dev)
    b1 = torch.zeros(Hd, device=dev); b2 = torch.zeros(D, device=dev)
    x = torch.zeros(rows, D, device=dev)
    idx = torch.arange(rows, device=dev, dtype=torch.int32)
    cnt = torch.tensor([rows], device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        rc = L.laud_adavit_mlp_fused(y.data_ptr(), rows, D, Hd, cnt.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                     x.data_ptr(), D, idx.data_ptr(), ctypes.c_void_p(st))
        assert rc == 0, _lib.last_error()
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * rows * D * Hd
    print(f"dbg={os.environ.get('LAUD_FM_DBG','0')} D={D} Hd={Hd} passes={passes} rows={rows}: {ms*1e3:.1f} us, {ms*1e3/passes:.1f} us/pass, {fl/ms/1e9:.0f} TFLOP/s")

if __name__ == "__main__":
    run(passes=int(sys.argv[1]) if len(sys.argv) > 1 else 3, tiles=int(sys.argv[2]) if len(sys.argv) > 2 else None)
