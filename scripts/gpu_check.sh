#!/bin/bash
# Standard GPU pass (run under gpurun from the repo root): parity tests, bench line with CPU baseline, reference arm,
# ncu launch list, one ncu --set full capture of the first forward pass's top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(d["ms_per_step"], d["e2e"], d["cpu_baseline"]); r=d["roofline"]; print(r["achieved"], r["frac"], r["ms_per_launch"], r["traffic"])
for k,v in r["other_stages"].items(): print(k, v["ms_per_launch"], v["tflops"], v["gbs"])
print(json.load(open("gpurun_out/bench_ref.json"))["value"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:stft_power_warp|db_mel_fast|conv1_kernel|gemm_tc_kernel|gemm_chain_pair_kernel|attention_tc_kernel|dwconv_reg" \
  -s 238 -c 14 -o gpurun_out/r01_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
grep "Profiling" gpurun_out/ncu_full.log | cut -c1-120
