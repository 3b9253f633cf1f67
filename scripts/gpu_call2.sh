#!/bin/bash
# PDL validation + timelines + ncu --set full captures of the top kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench pdl rc=$?"
B200ASR_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench nopdl rc=$?"
python - <<'PY'
import json
for n in ("pdl","nopdl"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("cpu_baseline"))
    except Exception as e: print(n, "ERR", e)
PY
timeout 120 python scripts/chain_dbg.py > gpurun_out/chain_dbg.log 2>&1
timeout 120 python scripts/attn_dbg.py > gpurun_out/attn_dbg.log 2>&1
timeout 120 python scripts/chain_prof.py > gpurun_out/chain_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:gemm_chain_kernel|attention_tc_kernel|stft_power_kernel|conv1_kernel|db_mel_kernel|gemm_tc_kernel<1, 144|dwconv_reg" \
  -s 148 -c 9 -o gpurun_out/r01_top9 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
