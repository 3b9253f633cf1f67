#!/bin/bash
# Two-GPU check (run under `gpurun --gpus 2`): bench config 2 (weak), 3 and 5 (strong) through torchrun + NCCL, and the reference arm's
# multi-rank behaviour (rank 0 prints, the others exit 0).
mkdir -p gpurun_out
for c in 2 3 5; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 10 --warmup 3 > gpurun_out/bench2_c$c.json 2> gpurun_out/bench2_c$c.err; echo "2-GPU config $c rc=$?"
done
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --sustain 0 2>&1 | grep -c "AllGather" > gpurun_out/nccl_allgather_count.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench2_ref.json 2> gpurun_out/bench2_ref.err; echo "2-GPU reference arm rc=$?"
python - <<'PY'
import json
for f in ("bench2_c2.json","bench2_c3.json","bench2_c5.json","bench2_ref.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", d["e2e"].get("ms_per_step"), "scaling", d["scaling"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-1500:])
print("AllGather lines in NCCL_DEBUG=INFO:", open("gpurun_out/nccl_allgather_count.txt").read().strip())
PY
