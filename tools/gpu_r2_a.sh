#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): round-2 first pass — box facts, GPU parity tests, a small and the full bench,
# one ncu --set full capture per kernel mode.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
R=r2a
{ nproc; free -g; lscpu | grep -i -E "numa|model name|socket|thread"; nvidia-smi topo -m; cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/memory.max 2>/dev/null; ulimit -l; } > gpurun_out/box_$R.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -5 gpurun_out/pytest_$R.log
timeout 600 python bench.py --rows 100000000 --blocks 8 --steps 5 --warmup 3 > gpurun_out/bench_small_$R.json 2> gpurun_out/bench_small_$R.err
tail -c 1500 gpurun_out/bench_small_$R.json; tail -5 gpurun_out/bench_small_$R.err
for W in c3 c2 c4 c5; do
  ONLY="--only $W"; [ $W = c3 ] && ONLY=""
  timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 30 -c 1 -f -o gpurun_out/${W}_kernel_$R \
    python bench.py $ONLY --rows 100000000 --blocks 8 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_${W}_$R.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 2500 gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
ls -la gpurun_out/*$R*
