#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu25.log
timeout 900 python bench.py > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "bench rc=$?"
cat gpurun_out/bench25.json | cut -c1-1500; tail -5 gpurun_out/bench25.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench25_ref.json 2> gpurun_out/bench25_ref.err; echo "ref rc=$?"
cat gpurun_out/bench25_ref.json | cut -c1-800
else
if [ "$N" = "2" ]; then
timeout 500 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x --timeout 300 -k "int8 or n1024_warm or qp_dense_n512 or duals" > gpurun_out/pytest_sharded26.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_sharded26.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
cat gpurun_out/bench_n$N.json | cut -c1-1500; grep -v "^\s*$" gpurun_out/bench_n$N.err | grep -v Warning | tail -12
fi
