"""Aggregate an ncu launch-list CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...) per kernel:  python scripts/ncu_launch_summary.py in.csv [out.txt] [traffic.json]"""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laudnet_b200.build import source_hash      # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, mi, vi, ui, ii = (h.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    name = r[ki].split("(")[0]
    name = name.replace("laud::<unnamed>::", "").replace("laud::", "").replace("void ", "")
    per.setdefault(r[ii], {"k": name[:48]})[r[mi]] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0, 0.0, 0.0])
    t = d.get("gpu__time_duration.sum", 0.0)
    a[0] += 1
    a[1] += t
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
    a[4] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
tot = sum(a[1] for a in agg.values())
lines = ["# per-kernel totals of one eager forward (ncu launch list, cold-cache serialised): " + sys.argv[1]]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{k:50s} n={a[0]:4d} time={a[1] / 1e3:7.3f} ms share={100 * a[1] / tot:5.1f}% dram_rd={a[2] / 1e9:6.2f} GB "
                 f"dram_wr={a[3] / 1e9:6.2f} GB tensor_active={a[4] / max(a[1], 1e-9):5.1f}%")
lines.append(f"total {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
print("\n".join(lines))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    conv = [a for k, a in agg.items() if "conv_tma_kernel" in k]
    n = sum(a[0] for a in conv)
    rd, wr, tm = sum(a[2] for a in conv), sum(a[3] for a in conv), sum(a[1] for a in conv)
    json.dump({"source": sys.argv[1] + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum; one eager "
               "forward, batch 256)", "kernel": "laud::conv_tma_kernel<*>", "build_source_hash": source_hash(), "launches_per_step": n,
               "dram_bytes_read_per_step": rd, "dram_bytes_write_per_step": wr, "dram_bytes_per_launch_avg": (rd + wr) / n,
               "tensor_pipe_active_pct_time_weighted": sum(a[4] for a in conv) / tm, "kernel_time_ms_per_step_ncu": tm / 1e3},
              open(sys.argv[3], "w"), indent=1)
