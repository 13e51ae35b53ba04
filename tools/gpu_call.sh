#!/bin/bash
# One gpurun call of round 2:  gpurun --timeout 1500 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_lasso_gpu.py -q -x -k "lasso" 2>&1 | tail -15
timeout -s KILL 400 python tools/lasso_bench.py 4096 2048 1024 512 > gpurun_out/lasso_bench.jsonl 2> gpurun_out/lasso_bench.err; echo "lasso_bench rc=$?"
cat gpurun_out/lasso_bench.jsonl; tail -5 gpurun_out/lasso_bench.err
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_call3.json 2> gpurun_out/bench_call3.err; echo "bench rc=$?"
cat gpurun_out/bench_call3.json; tail -30 gpurun_out/bench_call3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_call3.json 2> gpurun_out/bench_ref_call3.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_call3.json; tail -5 gpurun_out/bench_ref_call3.err
