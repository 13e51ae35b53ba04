#!/bin/bash
# One gpurun call of round 2 (2 GPUs):  gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_sharded_gpu.py -q -x > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"
tail -40 gpurun_out/pytest_sharded.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --sections socp,lasso,replicas > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; grep -v "^\s*$" gpurun_out/bench_n2.err | grep -v Warning | tail -25
