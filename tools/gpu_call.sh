#!/bin/bash
# One gpurun call of round 2:  gpurun --timeout 1500 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "potrf or lasso" 2>&1 | tail -15
IPM_B200_LIB=$PWD/interiorpoint-gpu_b200/lib/variants/liblassotiming.so timeout -s KILL 300 python tools/lasso_timing.py 4096 2048 512 > gpurun_out/lasso_timing.log 2>&1; echo "lasso_timing rc=$?"
cat gpurun_out/lasso_timing.log
timeout -s KILL 400 python tools/lasso_bench.py 4096 1024 512 > gpurun_out/lasso_bench.jsonl 2> gpurun_out/lasso_bench.err; echo "lasso_bench rc=$?"
cat gpurun_out/lasso_bench.jsonl; tail -5 gpurun_out/lasso_bench.err
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_default.log 2>&1; echo "pytest default rc=$?"
tail -30 gpurun_out/pytest_default.log
timeout 600 python bench.py --steps 1 --warmup 1 --sections qp --no-e2e --no-cpu-baseline > gpurun_out/bench_call4.json 2> gpurun_out/bench_call4.err; echo "bench rc=$?"
cat gpurun_out/bench_call4.json; tail -30 gpurun_out/bench_call4.err
