#!/bin/bash
# One gpurun call of round 2 (2 GPUs):  gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_default.log 2>&1; echo "pytest default rc=$?"
tail -40 gpurun_out/pytest_default.log
IPM_POTRF_DAG=1 timeout 600 python -m pytest tests -q -m gpu -k "not sharded" > gpurun_out/pytest_dag.log 2>&1; echo "pytest dag rc=$?"
tail -15 gpurun_out/pytest_dag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -30 gpurun_out/bench_n2.err
