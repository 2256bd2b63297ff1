#!/bin/bash
mkdir -p gpurun_out
R=r2p
run() { echo "=== $*"; env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -5; }
{
run X=1
run CUDA_LAUNCH_BLOCKING=1
run B2_JIT=off
run B2_JIT=off CUDA_LAUNCH_BLOCKING=1
run B2_NO_FAST_KERNEL=1
run B2_NO_FAST_FRONT=1
run ORDER=agg:host
run ORDER=agg:dev
run ORDER=sf:host,agg:dev
run BLOCKS=1
run KEYS=900
run SEED=1
} > gpurun_out/smoke_debug_$R.log 2>&1
cat gpurun_out/smoke_debug_$R.log
