#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2h
C2="--only c2 --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity"
B2_JIT_DEFS="-DB2_NO_COLD_OUTLINE" timeout 300 python bench.py $C2 > gpurun_out/c2_inline_$R.json 2> gpurun_out/c2_inline_$R.err
timeout 300 python bench.py $C2 > gpurun_out/c2_cold_$R.json 2> gpurun_out/c2_cold_$R.err
B2_JIT_DEFS="-DB2_NO_COLD_OUTLINE" timeout 600 python bench.py --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_inline_$R.json 2> gpurun_out/bench_inline_$R.err
timeout 600 python bench.py --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cold_$R.json 2> gpurun_out/bench_cold_$R.err
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -5 gpurun_out/pytest_$R.log; cat gpurun_out/jit_warm.log
ls -la gpurun_out/*$R*
