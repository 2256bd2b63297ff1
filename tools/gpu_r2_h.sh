#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2n
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -25 gpurun_out/pytest_$R.log
cat gpurun_out/jit_warm.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; tail -3 gpurun_out/smoke_$R.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 1000 gpurun_out/bench_$R.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2n.json'))
e=d['e2e']
print({k:v for k,v in e.items() if k not in('flat','warm','source','note')})
print('flat',e.get('flat',{}).get('ms_per_step'),'warm',e.get('warm',{}).get('ms_per_step'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'], 'kernel ms', d['roofline']['kernel_ms_per_step'], 'merge', d.get('merge_ms'))
for s in d.get('sub',[]): print(s['workload'][:20], s['value'], s['ms_per_step'], s['roofline']['frac'], s['roofline']['kernel_ms_per_step'], s.get('merge_ms'))
P
ls -la gpurun_out/*$R*
