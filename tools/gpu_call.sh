#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu22.log
timeout 600 python bench.py --sections qp,socp --no-cpu-baseline > gpurun_out/bench22.json 2> gpurun_out/bench22.err; echo "bench rc=$?"
cat gpurun_out/bench22.json | cut -c1-3000; tail -5 gpurun_out/bench22.err
