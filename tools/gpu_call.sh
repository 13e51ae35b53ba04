#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -k "int8 or group_lasso or large" --timeout 300 > gpurun_out/pytest_i8b.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_i8b.log
timeout 600 python bench.py --sections socp --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/bench_socp_i8.json 2> gpurun_out/bench_socp_i8.err; echo "bench rc=$?"
cat gpurun_out/bench_socp_i8.json; tail -5 gpurun_out/bench_socp_i8.err
