"""Bring-up diagnostics for the conv kernels on a GPU box: every case of
tests/test_gpu_parity.CONV_CASES through each implementation, with an error
breakdown by row / column when a case is off (does not stop at the first failure)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laudnet_b200 import _lib            # noqa: E402
from tests import test_gpu_parity as T   # noqa: E402

impls = {"umma": _lib.CONV_UMMA, "umma_wt": "wt"}
if os.environ.get("DIAG_HMMA"):
    impls["hmma"] = _lib.CONV_HMMA
only = sys.argv[1:] or list(T.CONV_CASES_ALL)
bad = 0
for name in only:
    d = T._conv_case(sum(map(ord, name)), **T.CONV_CASES_ALL[name])
    for iname, impl in impls.items():
        if impl == "wt" and name not in T.WT_CASES and name not in T.TAPBIAS_CASES:
            continue
        if impl != "wt" and name in T.TAPBIAS_CASES:
            continue
        try:
            err = T._run_conv_case(d, _lib.CONV_UMMA, use_wt=True) if impl == "wt" else T._run_conv_case(d, impl)
            flag = "OK " if err <= T.ACT_TOL else "BAD"
            print(f"{flag} {name:24s} {iname}: normalised max err {err:.3e}", flush=True)
            bad += err > T.ACT_TOL
        except AssertionError as e:
            bad += 1
            print(f"BAD {name:24s} {iname}: assertion: {e}", flush=True)
        except Exception as e:       # launch failure: context is gone, stop
            print(f"FATAL {name} {iname}: {type(e).__name__}: {e}", flush=True)
            sys.exit(2)
print("failures:", bad, "conv paths:", _lib.conv_path_counts())
sys.exit(1 if bad else 0)
