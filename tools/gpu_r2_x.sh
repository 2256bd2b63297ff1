#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -4 gpurun_out/pytest_r2.log; cat gpurun_out/jit_warm.log
run() { env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -1 | cut -c1-120; }
for i in 1 2 3 4 5 6 7 8 9 10; do run B2_JIT=off CUDA_LAUNCH_BLOCKING=1 ORDER=sf:host,sf:dev,agg:dev,agg:dev,agg:host; done | sort | uniq -c
for i in 1 2 3 4; do run X=1; done | sort | uniq -c
for i in 1 2 3; do timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
