#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/pair_dbg.py > gpurun_out/pair_dbg8.log 2>&1; head -8 gpurun_out/pair_dbg8.log | tail -4
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench8.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"])
PY
tail -3 gpurun_out/bench8.err
