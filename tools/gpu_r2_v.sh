#!/bin/bash
mkdir -p gpurun_out
R=r2v
run() { echo "=== $*"; env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -1 | cut -c1-160; }
{
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do run B2_JIT=off CUDA_LAUNCH_BLOCKING=1 ORDER=sf:host,sf:dev,agg:dev,agg:dev,agg:host; done
} > gpurun_out/smoke_debug_$R.log 2>&1
grep -c 'agg:host ok' gpurun_out/smoke_debug_$R.log; grep FAILED gpurun_out/smoke_debug_$R.log | head -5
