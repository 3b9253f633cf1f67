#!/bin/bash
# N = 2 launch exactly as the driver does it (torchrun, one rank per GPU, NCCL): own arm + reference arm
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
cat gpurun_out/bench_2gpu.json | cut -c1-700; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "2gpu ref rc=$?"
cut -c1-300 gpurun_out/bench_2gpu_ref.json; tail -3 gpurun_out/bench_2gpu_ref.err
