#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): the round's bench lines, the ncu launch list of the same command and one
# `ncu --set full` capture per dominant kernel (C3 aggregation, C2 scan, C4 TopN, C5 checksum).  Outputs land in
# gpurun_out/; tools/summarize_profiles_r2.py turns them into the tracked profiles/ files.
set -x
R=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$R.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/launches_$R.log 2>&1
SMALL="--rows 100000000 --blocks 8 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity"   # one launch = one 12.5M-entry block
ncu --set full --import-source on --clock-control none -k 'regex:fast_kernel|b2_fast_jit' -s 30 -c 1 -f -o gpurun_out/agg_kernel_$R python bench.py $SMALL > gpurun_out/ncu_agg_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 30 -c 1 -f -o gpurun_out/scan_kernel_$R python bench.py --only c2 --chunk 12500000 $SMALL > gpurun_out/ncu_scan_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:fast_kernel|b2_fast_jit' -s 40 -c 1 -f -o gpurun_out/topn_kernel_$R python bench.py --only c4 $SMALL > gpurun_out/ncu_topn_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:fast_kernel|b2_fast_jit' -s 30 -c 1 -f -o gpurun_out/checksum_kernel_$R python bench.py --only c5 $SMALL > gpurun_out/ncu_checksum_$R.log 2>&1
python tools/sst_probe.py 20000000 > gpurun_out/sst_probe_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:sst_expand' -s 2 -c 1 -f -o gpurun_out/sst_expand_$R python tools/sst_probe.py 20000000 > gpurun_out/ncu_sst_$R.log 2>&1
tail -c 600 gpurun_out/bench_$R.json; tail -c 300 gpurun_out/bench_reference_$R.json
ls -la gpurun_out/*$R*
