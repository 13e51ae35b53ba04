#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -c "
import cProfile, pstats, sys, runpy
sys.argv=['tools/socp_probe.py','16384']
cProfile.run('runpy.run_path(\"tools/socp_probe.py\", run_name=\"__main__\")', 'gpurun_out/socp.prof')
p=pstats.Stats('gpurun_out/socp.prof'); p.sort_stats('tottime').print_stats(22)
" > gpurun_out/socp_cprofile.log 2>&1; echo "rc=$?"
grep -A40 "tottime" gpurun_out/socp_cprofile.log | cut -c1-160 | head -45
