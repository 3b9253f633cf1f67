#!/bin/bash
# Round-2 full GPU pass on one B200: every GPU test, the per-stage error table, bench lines for configs 2 / 3 / 5 / 4 with CPU baselines,
# the reference arm of config 2, launch list + ncu --set full of the top kernels (config 2).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "rc=|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -12
timeout 600 python scripts/stage_errors.py gpurun_out/stage_errors.md > gpurun_out/stage_errors.log 2>&1; echo "stage rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --inflight 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_if1.json 2> gpurun_out/bench_c2_if1.err; echo "bench c2 one in flight rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "ref c2 rc=$?"
for c in 3 5 4; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "bench config $c rc=$?"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:stft_power_warp|db_mel_fast|conv1_f32x2|gemm_tc_kernel|gemm_chain_pair_kernel|attention_tc_kernel|dwconv_reg" \
  -s 238 -c 16 -o gpurun_out/r02_final python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)9|argmax_combine|ctc_collapse" -s 3 -c 3 -o gpurun_out/r02_ctcfc python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/ncu_ctcfc.log 2>&1; echo "ncu ctc head rc=$?"
python - <<'PY'
import json
for f in ("bench_c2.json","bench_c2_if1.json","bench_c2_ref.json","bench_c3.json","bench_c5.json","bench_c4.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", d["e2e"].get("ms_per_step"), "sust", d.get("sustained",{}).get("ms_per_step"), "launches", d["gpu_launches"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-1200:])
PY
