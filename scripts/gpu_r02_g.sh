#!/bin/bash
# session-layer GPU tests + batches-in-flight sweep (config 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py -m gpu -q -x > gpurun_out/pytest_session.log 2>&1; echo "pytest session rc=$?"
tail -25 gpurun_out/pytest_session.log | cut -c1-400
for n in 1 2 3 4; do
  timeout 600 python bench.py --inflight $n --steps 24 --warmup 3 --no-cpu-baseline --sustain 1 > gpurun_out/bench_if$n.json 2> gpurun_out/bench_if$n.err; echo "inflight $n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,3,4):
    try:
        d=json.load(open(f"gpurun_out/bench_if{n}.json")); print(n, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "sust", d.get("sustained",{}).get("ms_per_step"), "single", d.get("single_batch_in_flight"))
    except Exception as e: print(n, "ERR", e); print(open(f"gpurun_out/bench_if{n}.err").read()[-1500:])
PY
