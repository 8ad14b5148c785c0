#!/bin/bash
# usage (on the GPU box): scratch/gpu_quick.sh <tag> [bench args]   -> gpurun_out/<tag>_*.log
tag=$1; shift
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest.log
python bench.py --no-cpu --no-e2e "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_pytest.log
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "ms", d["ms_per_step"])
for v in d["config"].get("variants", []): print(v["name"][:40], v["value"], v["roofline_frac"])
print("single", d["config"].get("single_scan_latency_ms"))
PY
