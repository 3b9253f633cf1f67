#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/ubench_mma > gpurun_out/ubench_mma.log 2>&1; echo "ubench_mma rc=$?"
cat gpurun_out/ubench_mma.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"
B200ASR_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_nopdl.json 2> gpurun_out/bench4_nopdl.err; echo "bench nopdl rc=$?"
python - <<'PY'
import json
for n in ("bench4","bench4_nopdl"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json")); print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("cpu_baseline"))
    except Exception as e: print(n, "ERR", e)
PY
timeout 120 python scripts/chain_dbg.py > gpurun_out/chain_dbg4.log 2>&1; cat gpurun_out/chain_dbg4.log | tail -3
