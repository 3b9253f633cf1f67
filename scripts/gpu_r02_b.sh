#!/bin/bash
# Round-2 GPU pass: all GPU tests (no -x), bench line (fused subsampler), A/B against the unfused subsampler.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "precision|flipped|rc=|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200ASR_NO_FUSED_SUB=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused rc=$?"
python - <<'PY'
import json
for f in ("bench.json","bench_unfused.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"])
        for k,v in r["other_stages"].items(): print("  ",k, v.get("ms_per_launch"), v.get("tflops"), v.get("gbs"))
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-2000:])
PY
