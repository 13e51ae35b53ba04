#!/bin/bash
# Experiment: the whole GPU suite with the INT8 Hessian forced everywhere at 7 digits
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IPM_HESSIAN_I8=7 timeout 600 python -m pytest tests/test_solvers_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_forced_s7.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_forced_s7.log | cut -c1-250
