#!/bin/bash
mkdir -p gpurun_out
R=r2s
run() { echo "=== $*"; env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -2; }
{
for i in 1 2 3 4 5 6 7 8; do run CUDA_LAUNCH_BLOCKING=1; done
for i in 1 2 3 4; do run X=1; done
for i in 1 2 3 4 5 6; do run CUDA_LAUNCH_BLOCKING=1 ORDER=sf:host,sf:dev,agg:dev,agg:dev; done
for i in 1 2 3 4; do run CUDA_LAUNCH_BLOCKING=1 B2_FAST_KERNEL_ON_HOST_BLOCKS=1; done
} > gpurun_out/smoke_debug_$R.log 2>&1
grep -c ok gpurun_out/smoke_debug_$R.log; grep -B2 FAILED gpurun_out/smoke_debug_$R.log | cut -c1-200
for i in 1 2 3; do timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
