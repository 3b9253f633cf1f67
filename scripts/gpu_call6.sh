#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench6.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["cpu_baseline"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"])
for k,v in r["other_stages"].items(): print(k, v["ms_per_launch"], v["tflops"], v["gbs"])
PY
tail -3 gpurun_out/bench6.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:gemm_chain_pair_kernel|attention_tc_kernel|stft_power_warp|conv1_kernel|db_mel_fast|gemm_tc_kernel<1, 144|dwconv_reg|gemm_tc_kernel<5, 224" \
  -s 219 -c 12 -o gpurun_out/r01_top12 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full6.log 2>&1; echo "ncu rc=$?"
grep -c "Profiling" gpurun_out/ncu_full6.log; grep "Profiling" gpurun_out/ncu_full6.log | head -14
