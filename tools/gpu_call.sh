#!/bin/bash
# One gpurun call of round 2:  gpurun --timeout 1500 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_default.log 2>&1; echo "pytest default rc=$?"
tail -30 gpurun_out/pytest_default.log
IPM_POTRF_DAG=1 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_dag.log 2>&1; echo "pytest dag rc=$?"
tail -15 gpurun_out/pytest_dag.log
timeout 600 python bench.py --steps 1 --warmup 1 --sections qp --no-e2e --no-cpu-baseline > gpurun_out/bench_call5.json 2> gpurun_out/bench_call5.err; echo "bench rc=$?"
cat gpurun_out/bench_call5.json; tail -30 gpurun_out/bench_call5.err
timeout 600 python tools/lib_context.py > gpurun_out/lib_context.json 2> gpurun_out/lib_context.err; echo "lib_context rc=$?"
cat gpurun_out/lib_context.json; tail -5 gpurun_out/lib_context.err
