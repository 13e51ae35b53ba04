#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IPM_HESSIAN_I8=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu30_forced_i8.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu30_forced_i8.log
