#!/bin/bash
# One gpurun call of round 2 (8 GPUs):  gpurun --gpus 8 --timeout 700 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
cat gpurun_out/bench_n$N.json; grep -v "^\s*$" gpurun_out/bench_n$N.err | grep -v Warning | tail -15
