#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench38.json 2> gpurun_out/bench38.err; echo "bench rc=$?"
cat gpurun_out/bench38.json | cut -c1-900; tail -3 gpurun_out/bench38.err
