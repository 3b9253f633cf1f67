#!/bin/bash
# N = 2 launch exactly as the driver does it (torchrun, one rank per GPU, NCCL): own arm + reference arm
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
head -c 200 gpurun_out/bench_2gpu.json; echo; python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_2gpu.json')]
print("stdout lines:", len(lines))
d=json.loads(lines[-1]); print(d['ms_per_step'], d['value'], d['e2e'])
PY
tail -3 gpurun_out/bench_2gpu.err
