#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench10.json")); print(d["ms_per_step"], d["e2e"], d["cpu_baseline"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"], r["traffic"])
PY
tail -3 gpurun_out/bench10.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench10_ref.json 2> gpurun_out/bench10_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench10_ref.json
