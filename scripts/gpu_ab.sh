#!/bin/bash
# A/B helper: GPU tests, then the bench line (stage table included) for the current build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ab.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"])
for k,v in d["roofline"]["other_stages"].items(): print(k, v["ms_per_launch"], v["gbs"])
PY
