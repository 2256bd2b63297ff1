#!/bin/bash
mkdir -p gpurun_out
R=r2w
run() { echo "=== $*"; env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -1 | cut -c1-120; }
{
for i in 1 2 3 4 5 6; do run B2_JIT=off CUDA_LAUNCH_BLOCKING=1 SEED=1 KEYS=2000 ORDER=sf:host,sf:dev,agg:dev,agg:dev,agg:host; done
for i in 1 2 3 4 5 6; do run B2_JIT=off CUDA_LAUNCH_BLOCKING=1 SEED=3 KEYS=900 ORDER=sf:host,sf:dev,agg:dev,agg:dev,agg:host; done
for i in 1 2 3 4 5 6; do run B2_JIT=off CUDA_LAUNCH_BLOCKING=1 SEED=3 KEYS=2000 BLOCKS=1 ORDER=sf:host,sf:dev,agg:dev,agg:dev,agg:host; done
} > gpurun_out/smoke_debug_$R.log 2>&1
cat gpurun_out/smoke_debug_$R.log | grep -v '^===' | sort | uniq -c
grep -B1 FAILED gpurun_out/smoke_debug_$R.log | grep '===' | sort | uniq -c
