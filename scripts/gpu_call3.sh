#!/bin/bash
# new frontend kernels (warp FFT, fast dB+mel) A/B vs legacy, TMA/MMA micro-benchmarks, PDL edge check, launch list
mkdir -p gpurun_out
timeout 120 ./scripts/ubench_tma_mma > gpurun_out/ubench.log 2>&1; echo "ubench rc=$?"
cat gpurun_out/ubench.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
B200ASR_GRAPH_DBG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench new rc=$?"
B200ASR_STFT_LEGACY=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_legacy.json 2> gpurun_out/bench_legacy.err; echo "bench legacy rc=$?"
python - <<'PY'
import json
for n in ("new","legacy"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["other_stages"]["stft"])
    except Exception as e: print(n, "ERR", e)
PY
cat gpurun_out/bench_new.err | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1; echo "ncu rc=$?"
