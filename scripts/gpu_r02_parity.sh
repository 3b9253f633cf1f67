#!/bin/bash
# Round-2 GPU pass #1 (run under gpurun from the repo root): parity tests (incl. the reference-sized ones), per-stage error table,
# one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "max \||rc=|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python scripts/stage_errors.py gpurun_out/stage_errors.md > gpurun_out/stage_errors.log 2>&1; echo "stage rc=$?"
tail -25 gpurun_out/stage_errors.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(d["ms_per_step"], d["e2e"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"])
for k,v in r["other_stages"].items(): print(k, v["ms_per_launch"], v["tflops"], v["gbs"])
PY
