#!/bin/bash
# default bench (with CPU baseline) + smoke(), summarised; run under gpurun from the repo root
mkdir -p gpurun_out
python bench.py > gpurun_out/${1:-run}_bench.json 2> gpurun_out/${1:-run}_bench.err
tail -3 gpurun_out/${1:-run}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${1:-run}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"])
print(json.dumps(d["roofline"])[:1600])
print(d["cpu_baseline"]); print(d["net"]); print(d["clocks"])
PY
python -c "import __graft_entry__ as g; g.smoke()"
