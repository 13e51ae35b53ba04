#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke31.log 2>&1; echo "smoke rc=$?"
tail -6 gpurun_out/smoke31.log
