#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/pair_dbg.py > gpurun_out/pair_dbg.log 2>&1; tail -12 gpurun_out/pair_dbg.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench7.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"])
for k,v in r["other_stages"].items(): print(k, v["ms_per_launch"], v["tflops"], v["gbs"])
PY
tail -3 gpurun_out/bench7.err
