#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2f
timeout 600 python -m pytest tests/test_gpu_sst.py -m gpu -x -q --durations=8 > gpurun_out/pytest_sst_$R.log 2>&1; echo "pytest sst rc=$?" >> gpurun_out/pytest_sst_$R.log
tail -30 gpurun_out/pytest_sst_$R.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 --deselect tests/test_gpu_sst.py > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -14 gpurun_out/pytest_$R.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 1500 gpurun_out/bench_$R.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2f.json'))
print(json.dumps(d['e2e'],indent=1)[:3000])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
for s in d.get('sub',[]): print(s['workload'][:20], s['value'], s['ms_per_step'], s['roofline']['frac'], s['roofline']['kernel_ms_per_step'], s.get('merge_ms'))
P
ls -la gpurun_out/*$R*
