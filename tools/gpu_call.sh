#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
cat gpurun_out/bench_n$N.json | cut -c1-1200; grep -v "^\s*$" gpurun_out/bench_n$N.err | grep -v Warning | tail -12
