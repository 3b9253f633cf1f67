#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench11.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench11.err
