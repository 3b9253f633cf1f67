#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -s -k "attention or beam or reference_wav or noise_batch or batch_consistency or subsampling" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log
grep -E "beam, prob|precision|rc=|passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_f.log | tail -20
for v in "" "B200ASR_ATTN_ASYNC=1"; do
  env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "[$v] rc=$?"
  python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_ab.json")); r=d["roofline"]; o=r["other_stages"]
    print("   ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "ffn", round(r["ms_per_launch"]*1e3,2), "conv2", o["conv2"].get("ms_per_launch"), "attn", o["attention"].get("ms_per_launch"), "qkv", o["qkv"].get("ms_per_launch"))
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_ab.err").read()[-1500:])
PY
done
