#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2e
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -40 gpurun_out/pytest_$R.log
timeout 600 python bench.py --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_small_$R.json 2> gpurun_out/bench_small_$R.err
tail -c 300 gpurun_out/bench_small_$R.json; tail -5 gpurun_out/bench_small_$R.err
for CH in 4194304 8388608 33554432; do
  timeout 200 python bench.py --only c2 --chunk $CH --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/c2_chunk_${CH}_$R.json 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_c4_$R.csv \
    python bench.py --only c4 --rows 100000000 --blocks 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_l4_$R.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:fast_kernel|b2_fast_jit' -s 40 -c 1 -f -o gpurun_out/c4_fast_$R \
    python bench.py --only c4 --rows 100000000 --blocks 8 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_c4_$R.log 2>&1
ls -la gpurun_out/*$R*
