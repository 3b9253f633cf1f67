#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c1new.json 2> gpurun_out/bench_c1new.err; echo "bench rc=$?"
B200ASR_CONV1_LEGACY=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c1old.json 2> gpurun_out/bench_c1old.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("c1new","c1old"):
    d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["other_stages"]["conv1"])
PY
