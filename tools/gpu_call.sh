#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python tools/ozaki_syrk_test.py tile 0; timeout 120 python tools/ozaki_syrk_test.py full 0 1000 2100 8; timeout 200 python tools/ozaki_syrk_test.py time 0 8192 16384 8; timeout 200 python tools/ozaki_syrk_test.py time 0 8192 16384 7 ) > gpurun_out/ozaki_v3.jsonl 2> gpurun_out/ozaki_v3.err; echo "rc=$?"
cut -c1-1500 gpurun_out/ozaki_v3.jsonl; tail -5 gpurun_out/ozaki_v3.err
