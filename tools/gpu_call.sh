#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
IPM_PEER_POTRF=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --sections none --no-e2e > gpurun_out/bench_n${N}_peerpotrf.json 2> gpurun_out/bench_n${N}_peerpotrf.err; echo "bench n$N rc=$?"
cat gpurun_out/bench_n${N}_peerpotrf.json | cut -c1-600; grep -v "^\s*$" gpurun_out/bench_n${N}_peerpotrf.err | grep -v Warning | tail -8
