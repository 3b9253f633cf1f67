#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pair_direct" > gpurun_out/pytest_direct.log 2>&1; echo "direct rc=$?"; tail -5 gpurun_out/pytest_direct.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench13.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench13.err
