#!/bin/bash
# N-GPU check (run under `gpurun --gpus N`, N passed as $1): the driver's own command lines for config 2 (weak) plus configs 3 and 5
# (strong) through torchrun + NCCL; NCCL collective count per step; the reference arm's multi-rank behaviour.
N=${1:-2}
mkdir -p gpurun_out
for c in 2 3 5; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 20 --warmup 3 > gpurun_out/bench${N}_c$c.json 2> gpurun_out/bench${N}_c$c.err; echo "$N-GPU config $c rc=$?"
done
if [ "$N" = "2" ]; then
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --sustain 0 > gpurun_out/nccl_coll_$N.log 2>&1
grep -c "AllGather" gpurun_out/nccl_coll_$N.log > gpurun_out/nccl_allgather_count_$N.txt
grep "AllGather" gpurun_out/nccl_coll_$N.log | head -3 | cut -c1-300
else echo "n/a" > gpurun_out/nccl_allgather_count_$N.txt; fi
if [ "$N" = "2" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench${N}_ref.json 2> gpurun_out/bench${N}_ref.err; echo "$N-GPU reference arm rc=$?"
fi
python - $N <<'PY'
import json, sys
N=sys.argv[1]
for f in (f"bench{N}_c2.json",f"bench{N}_c3.json",f"bench{N}_c5.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", d["e2e"].get("ms_per_step"), "scaling", d["scaling"], "inflight", d["config"].get("batches_in_flight"), "single", (d.get("single_batch_in_flight") or {}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/"+f.replace("json","err")).read()[-1500:])
print("AllGather lines:", open(f"gpurun_out/nccl_allgather_count_{N}.txt").read().strip())
PY
