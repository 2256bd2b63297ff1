#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): the round's bench lines, the ncu launch list and one full capture of the
# dominant kernel.  Outputs land in gpurun_out/; tools/summarize_profiles.py turns them into the tracked profiles/ files.
set -x
R=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 8 -c 1 -f -o gpurun_out/scan_kernel_$R \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --chunk 12500000 > gpurun_out/ncu_$R.log 2>&1  # one launch = one 12.5M-entry block
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-jit > gpurun_out/bench_generic_kernel_$R.json 2>/dev/null
python tools/bench_extra.py --rows 100000000 > gpurun_out/other_configs_$R.jsonl 2> gpurun_out/other_configs_$R.err
tail -c 600 gpurun_out/bench_$R.json; tail -c 300 gpurun_out/bench_reference_$R.json; tail -3 gpurun_out/other_configs_$R.jsonl | cut -c1-160
