#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -12 gpurun_out/pytest_r2.log
cat gpurun_out/jit_warm.log
for i in 1 2 3; do timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
bash tools/refresh_profiles_r2.sh r2
