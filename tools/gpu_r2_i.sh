#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2i
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -5 gpurun_out/pytest_$R.log
cat gpurun_out/jit_warm.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 1000 gpurun_out/bench_$R.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2i.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'], 'kernel ms', d['roofline']['kernel_ms_per_step'], 'merge', d.get('merge_ms'))
for s in d.get('sub',[]): print(s['workload'][:20], s['value'], s['ms_per_step'], s['roofline']['frac'], s['roofline']['kernel_ms_per_step'], s.get('merge_ms'))
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 260 --csv --log-file gpurun_out/launches_c4_$R.csv \
    python bench.py --only c4 --rows 100000000 --blocks 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_l4_$R.log 2>&1
ls -la gpurun_out/*$R*
