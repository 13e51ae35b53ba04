#!/bin/bash
# One gpurun call of round 2 (validation first): usage  gpurun --timeout 1500 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# 1. the pipelined / plain tile-DAG kernels and the stream-ordered code: correctness at ragged sizes, then timing
timeout -s KILL 200 python tools/potrf_ab.py 2048 4096 8192 16384 > gpurun_out/potrf_ab.log 2>&1; echo "potrf_ab rc=$?"
tail -20 gpurun_out/potrf_ab.log
# 2. step-by-step traces of the three cases that failed on the device in round 1
timeout 300 python tools/debug_parity.py option_cases.json lp_dense_n64_warm__update_slacks_every_3 lp_dense_n64_cold__update_slacks_every_2 > gpurun_out/debug_options.log 2>&1
timeout 300 python tools/debug_parity.py dual_cases.json lp_dense_n64_warm_duals > gpurun_out/debug_duals.log 2>&1
cat gpurun_out/debug_options.log gpurun_out/debug_duals.log
# 3. the whole GPU suite, default factorisation policy, then with the tile-DAG kernel forced for every admissible size
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_default.log 2>&1; echo "pytest default rc=$?"
tail -30 gpurun_out/pytest_default.log
IPM_POTRF_DAG=1 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_dag.log 2>&1; echo "pytest dag rc=$?"
tail -15 gpurun_out/pytest_dag.log
# 4. bench (short) with the factorisation section
timeout 600 python bench.py --steps 2 --warmup 3 --factorisation > gpurun_out/bench_call1.json 2> gpurun_out/bench_call1.err; echo "bench rc=$?"
cat gpurun_out/bench_call1.json; tail -5 gpurun_out/bench_call1.err
