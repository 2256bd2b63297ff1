#!/bin/bash
# Runs ON THE GPU BOX: round-2 second pass — lean kernels: parity tests (incl. a memcheck subset), bench, ncu per mode.
set -x
mkdir -p gpurun_out
R=r2b
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -15 gpurun_out/pytest_$R.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "checksum or topn_device or exact_layout or real_sums" > gpurun_out/memcheck_$R.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_$R.log
tail -8 gpurun_out/memcheck_$R.log
timeout 600 python bench.py --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_small_$R.json 2> gpurun_out/bench_small_$R.err
tail -c 600 gpurun_out/bench_small_$R.json; tail -5 gpurun_out/bench_small_$R.err
for W in c3 c4 c5; do
  ONLY="--only $W"; [ $W = c3 ] && ONLY=""
  timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:fast_kernel|b2_fast_jit' -s 30 -c 1 -f -o gpurun_out/${W}_fast_$R \
    python bench.py $ONLY --rows 100000000 --blocks 8 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_${W}_$R.log 2>&1
done
ls -la gpurun_out/*$R*
