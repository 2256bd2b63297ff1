#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2i
C2="--only c2 --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity"
B2_JIT_DEFS="-DB2_NO_COLD_OUTLINE" timeout 300 python bench.py $C2 > gpurun_out/c2_inline_$R.json 2> gpurun_out/c2_inline_$R.err
timeout 300 python bench.py $C2 > gpurun_out/c2_cold_$R.json 2> gpurun_out/c2_cold_$R.err
B2_JIT_DEFS="-DB2_NO_COLD_OUTLINE -DB2_NO_IDX" timeout 300 python bench.py $C2 > gpurun_out/c2_inline_noidx_$R.json 2> gpurun_out/c2_inline_noidx_$R.err
B2_JIT_DEFS="-DB2_NO_IDX" timeout 300 python bench.py $C2 > gpurun_out/c2_cold_noidx_$R.json 2> gpurun_out/c2_cold_noidx_$R.err
ls -la gpurun_out/*$R*
