#!/bin/bash
# A/B of the eight-channels-per-thread fp16 conv1 kernel against the four-channel one
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_sized.py -m gpu -q -x > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_m.log | cut -c1-400
timeout 600 python bench.py --inflight 2 --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err; echo "h8 rc=$?"
B200ASR_CONV1_H4=1 timeout 600 python bench.py --inflight 2 --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_m2.json 2> gpurun_out/bench_m2.err; echo "h4 rc=$?"
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.load(open(f"gpurun_out/bench_m{n}.json")); r=d["roofline"]
        print(n, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "single", d.get("single_batch_in_flight",{}).get("ms_per_step"), {k: round(v.get("ms_per_launch",0)*1e3,2) for k,v in r["other_stages"].items()})
    except Exception as e: print(n, "ERR", e); print(open(f"gpurun_out/bench_m{n}.err").read()[-1500:])
PY
