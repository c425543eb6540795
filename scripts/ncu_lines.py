"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: python scripts/ncu_lines.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iS = [i for i, c in enumerate(hdr) if c == "# Samples"][0]
        iI = [i for i, c in enumerate(hdr) if c == "Instructions Executed"][0]
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    line, src = r[0], r[1]
    if line.strip():
        cur = (int(line), src.strip())
        agg.setdefault(cur, [0, 0])
        # a cuda line row carries aggregated values already
        try:
            agg[cur][0] = int(r[iS]); agg[cur][1] = int(r[iI])
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (ln, src), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} smp {100*s/tot_s:5.1f}% inst {100*i/tot_i:5.1f}%  {src[:110]}")
