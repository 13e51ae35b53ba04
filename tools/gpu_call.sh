#!/bin/bash
# One gpurun call of round 2 (1 GPU): tests, Ozaki probe, ncu captures
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "potrf" 2>&1 | tail -8
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_default.log 2>&1; echo "pytest default rc=$?"
tail -25 gpurun_out/pytest_default.log
timeout 300 python tools/ozaki_probe.py time > gpurun_out/ozaki_time.json 2> gpurun_out/ozaki.err; echo "ozaki time rc=$?"; cat gpurun_out/ozaki_time.json
timeout 300 python tools/ozaki_probe.py accuracy 2048 > gpurun_out/ozaki_accuracy.json 2>> gpurun_out/ozaki.err; echo "ozaki acc rc=$?"; cat gpurun_out/ozaki_accuracy.json
timeout 400 python tools/ozaki_probe.py solve 1024 > gpurun_out/ozaki_solve.json 2>> gpurun_out/ozaki.err; echo "ozaki solve rc=$?"; cat gpurun_out/ozaki_solve.json; tail -5 gpurun_out/ozaki.err
# ncu: launch list of one LP solve of the bench workload, then full captures of the two new persistent kernels
CMD="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --sections none"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/bench_launches_r02.csv $CMD > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/potrf_probe.py 8192 > gpurun_out/plain_potrf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:potrf_dag2 -c 1 -f -o gpurun_out/potrf_dag2_r02 python tools/potrf_probe.py 8192 > gpurun_out/ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
python tools/lasso_probe.py 4096 3 > gpurun_out/plain_lasso.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lasso_admm_multi -s 1 -c 1 -f -o gpurun_out/lasso_multi_r02 python tools/lasso_probe.py 4096 3 > gpurun_out/ncu_lasso.log 2>&1; echo "ncu lasso rc=$?"
tail -3 gpurun_out/ncu_bench.log gpurun_out/ncu_potrf.log gpurun_out/ncu_lasso.log
