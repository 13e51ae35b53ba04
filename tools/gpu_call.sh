#!/bin/bash
# One gpurun call of round 2:  gpurun --timeout 1500 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "lasso or watchdog or potrf" 2>&1 | tail -15
timeout -s KILL 300 python -m pytest tests/test_lasso_gpu.py -q 2>&1 | tail -15
timeout -s KILL 400 python tools/lasso_bench.py 4096 2048 1024 512 > gpurun_out/lasso_bench.jsonl 2> gpurun_out/lasso_bench.err; echo "lasso_bench rc=$?"
cat gpurun_out/lasso_bench.jsonl; tail -5 gpurun_out/lasso_bench.err
timeout -s KILL 200 python tools/potrf_ab.py 512 1024 1536 2048 3072 > gpurun_out/potrf_ab_small.log 2>&1; echo "potrf_ab rc=$?"
tail -20 gpurun_out/potrf_ab_small.log
timeout 600 python bench.py --steps 2 --warmup 3 --factorisation > gpurun_out/bench_call2.json 2> gpurun_out/bench_call2.err; echo "bench rc=$?"
cat gpurun_out/bench_call2.json; tail -5 gpurun_out/bench_call2.err
