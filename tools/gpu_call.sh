#!/bin/bash
# One gpurun call of round 2 (1 GPU):  gpurun --timeout 900 -- 'bash tools/gpu_call.sh > gpurun_out/call.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 800 python tools/ozaki_syrk_test.py auto > gpurun_out/ozaki_syrk.jsonl 2> gpurun_out/ozaki_syrk.err; echo "rc=$?"
cat gpurun_out/ozaki_syrk.jsonl; tail -20 gpurun_out/ozaki_syrk.err
nvidia-smi --query-gpu=name,clocks.sm --format=csv
