#!/bin/bash
mkdir -p gpurun_out
R=r2r
{
for T in racecheck synccheck initcheck; do
echo "=== $T"
B2_JIT=off timeout 500 compute-sanitizer --tool $T python tools/smoke_debug.py 2>&1 | grep -v 'Host Frame\|^=========         in \|^=========$' | head -50
done
} > gpurun_out/smoke_san_$R.log 2>&1
cut -c1-260 gpurun_out/smoke_san_$R.log
