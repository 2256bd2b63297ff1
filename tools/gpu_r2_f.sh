#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2f
C2="--only c2 --rows 100000000 --blocks 8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity --chunk 12500000"
(cd _old_tree && timeout 300 python bench.py $C2 > ../gpurun_out/c2_old_$R.json 2> ../gpurun_out/c2_old_$R.err)
timeout 300 python bench.py $C2 > gpurun_out/c2_new_$R.json 2> gpurun_out/c2_new_$R.err
(cd _old_tree && timeout 300 python bench.py $C2 > ../gpurun_out/c2_old2_$R.json 2>> ../gpurun_out/c2_old_$R.err)
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 12 -c 1 -f -o gpurun_out/c2_new_$R python bench.py $C2 --steps 1 > gpurun_out/ncu_c2_new_$R.log 2>&1
(cd _old_tree && timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 12 -c 1 -f -o ../gpurun_out/c2_old_$R python bench.py $C2 --steps 1 > ../gpurun_out/ncu_c2_old_$R.log 2>&1)
tail -c 400 gpurun_out/c2_old_$R.json; tail -c 400 gpurun_out/c2_new_$R.json; tail -c 400 gpurun_out/c2_old2_$R.json
ls -la gpurun_out/*$R*
