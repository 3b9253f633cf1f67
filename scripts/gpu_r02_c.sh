#!/bin/bash
# Round-2 GPU pass: all GPU tests, bench config 2 (fused / unfused subsampler A/B), new bench configs 3, 5, 4 (short).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "precision|flipped|worst|rc=|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200ASR_NO_FUSED_SUB=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused rc=$?"
for c in 3 5 4; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "bench config $c rc=$?"; done
python - <<'PY'
import json
for f in ("bench.json","bench_unfused.json","bench_c3.json","bench_c5.json","bench_c4.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "value", round(d["value"]), "sustained", d.get("sustained",{}).get("ms_per_step"), "launches", d["gpu_launches"]); r=d["roofline"]; print("   roof", r.get("achieved"), r.get("frac"), r.get("ms_per_launch"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        for k,v in (r.get("other_stages") or {}).items(): print("      ",k, v.get("ms_per_launch"), v.get("tflops"), v.get("gbs"))
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-1500:])
PY
