#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chained" > gpurun_out/pytest_chain.log 2>&1; echo "pytest chain rc=$?"
tail -15 gpurun_out/pytest_chain.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
B200ASR_NO_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_nopair.json 2> gpurun_out/bench5_nopair.err; echo "bench nopair rc=$?"
python - <<'PY'
import json
for n in ("bench5","bench5_nopair"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json")); print(n, d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/bench5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench5.log 2>&1; echo "ncu rc=$?"
