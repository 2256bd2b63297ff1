#!/bin/bash
mkdir -p gpurun_out
R=r2q
run() { echo "=== $*"; env "$@" timeout 120 python tools/smoke_debug.py 2>&1 | tail -2; }
{
for i in 1 2 3 4 5 6; do run CUDA_LAUNCH_BLOCKING=1 B2_NO_STAGE_PAD=1; done
for i in 1 2 3 4 5 6; do run CUDA_LAUNCH_BLOCKING=1; done
for i in 1 2 3; do run B2_JIT=off B2_NO_STAGE_PAD=1; done
for i in 1 2 3; do run B2_JIT=off; done
echo "=== initcheck"
B2_NO_STAGE_PAD=1 timeout 400 compute-sanitizer --tool initcheck --track-unused-memory no python tools/smoke_debug.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in \|^=========$' | head -60
} > gpurun_out/smoke_debug_$R.log 2>&1
cat gpurun_out/smoke_debug_$R.log | cut -c1-220
