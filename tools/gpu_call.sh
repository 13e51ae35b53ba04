#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu32.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu32.log
