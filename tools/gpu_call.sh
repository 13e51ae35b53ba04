#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu29.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu29.log
timeout 300 python tools/hess_i8_sizes.py 8192 16384 > gpurun_out/hess_i8_one.jsonl 2>&1; cat gpurun_out/hess_i8_one.jsonl
timeout 600 python bench.py --sections none --no-cpu-baseline > gpurun_out/bench29.json 2> gpurun_out/bench29.err; echo "bench rc=$?"
cat gpurun_out/bench29.json | cut -c1-700; tail -3 gpurun_out/bench29.err
