#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -s -k "beam or translator or asr_surface or subsampling or gemm" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_d.log
grep -E "precision|rc=|passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_d.log | tail -30
python scripts/prof_subsample.py
B200ASR_NO_FUSED_SUB=1 python scripts/prof_subsample.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_subsample -s 2 -c 1 -o gpurun_out/r02_convsub python scripts/prof_subsample.py > gpurun_out/ncu_convsub.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200ASR_NO_FUSED_SUB=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused rc=$?"
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for f in ("bench.json","bench_unfused.json","bench_c4.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "launches", d["gpu_launches"]); r=d["roofline"]; print("   roof", r.get("achieved"), r.get("frac"), r.get("ms_per_launch"))
        for k,v in (r.get("other_stages") or {}).items(): print("      ",k, v.get("ms_per_launch"), v.get("tflops"), v.get("gbs"))
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-1500:])
PY
