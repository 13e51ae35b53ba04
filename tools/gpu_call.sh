#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py -m gpu -q -x -k "hess_i8 or int8" --timeout 300 > gpurun_out/pytest_i8c.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_i8c.log
timeout 300 python tools/hess_i8_sizes.py > gpurun_out/hess_i8_sizes2.jsonl 2> gpurun_out/hess_i8_sizes2.err; echo "sizes rc=$?"
cat gpurun_out/hess_i8_sizes2.jsonl; tail -5 gpurun_out/hess_i8_sizes2.err
