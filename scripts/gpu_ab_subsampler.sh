#!/bin/bash
# A/B of a kernel change: GEMM + kernel + parity tests, then bench config 2 with one and two batches in flight (and the switch off); prints the per-stage table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_reference_sized.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_j.log | cut -c1-400
timeout 600 python bench.py --inflight 1 --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_j1.json 2> gpurun_out/bench_j1.err; echo "inflight 1 rc=$?"
timeout 600 python bench.py --inflight 2 --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_j2.json 2> gpurun_out/bench_j2.err; echo "inflight 2 rc=$?"
B200ASR_NO_CONV_F16=2 timeout 600 python bench.py --inflight 1 --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_j3.json 2> gpurun_out/bench_j3.err; echo "inflight 1, fp32 conv1 map rc=$?"
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_j4.json 2> gpurun_out/bench_j4.err; echo "config 3 rc=$?"
python - <<'PY'
import json
for n in (1,2,3,4):
    try:
        d=json.load(open(f"gpurun_out/bench_j{n}.json")); r=d["roofline"]
        print(n, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "ffn_chain", round(r["ms_per_launch"]*1e3,2), {k: round(v.get("ms_per_launch",0)*1e3,2) for k,v in r["other_stages"].items()})
    except Exception as e: print(n, "ERR", e); print(open(f"gpurun_out/bench_j{n}.err").read()[-1500:])
PY
