#!/bin/bash
# what the driver runs at round end, in its order: GPU tests (-x), smoke(), reference arm, own arm with default flags
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/bench_ref.json")); print("reference", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
d=json.load(open("gpurun_out/bench.json")); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
print("ratio e2e/ref", d["e2e"]["value"]/r["value"])
PY
