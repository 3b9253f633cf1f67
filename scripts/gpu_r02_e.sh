#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py tests/test_gpu_reference_sized.py -m gpu -q -s -k "beam or translator or subsampling or attention or config2 or reference_wav or stage_taps" > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e.log
grep -E "beam mismatch|precision|rc=|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_e.log | tail -30
python scripts/prof_subsample.py
B200ASR_NO_FUSED_SUB=1 python scripts/prof_subsample.py
for v in "" "B200ASR_NO_FUSED_SUB=1" "B200ASR_NO_ATTN_ASYNC=1" "B200ASR_NO_FUSED_SUB=1 B200ASR_NO_ATTN_ASYNC=1"; do
  env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "[$v] rc=$?"
  python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_ab.json")); r=d["roofline"]; o=r["other_stages"]
    print("   ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "ffn", round(r["ms_per_launch"]*1e3,2), "conv2", o["conv2"].get("ms_per_launch"), "attn", o["attention"].get("ms_per_launch"), "qkv", o["qkv"].get("ms_per_launch"))
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_ab.err").read()[-1500:])
PY
done
