#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2g
timeout 900 python -m pytest tests/test_gpu_sst.py -m gpu -x -q --durations=8 > gpurun_out/pytest_sst_$R.log 2>&1; echo "pytest sst rc=$?" >> gpurun_out/pytest_sst_$R.log
tail -30 gpurun_out/pytest_sst_$R.log
timeout 300 python tools/sst_probe.py 20000000 > gpurun_out/sst_probe_$R.log 2>&1; cat gpurun_out/sst_probe_$R.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:sst_' --csv --log-file gpurun_out/sst_launches_$R.csv python tools/sst_probe.py 20000000 > /dev/null 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/sst_launches_r2g.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i+1; break
ki,vi=h.index('Kernel Name'),h.index('Metric Value')
for r in rows[st:][-12:]: print(r[ki][:60], r[vi])
P
timeout 900 python bench.py --steps 5 --warmup 3 --no-sub > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 1500 gpurun_out/bench_$R.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
e=d['e2e']
print({k:v for k,v in e.items() if k not in('flat','warm','source','note')})
print('flat',e.get('flat',{}).get('ms_per_step'),'warm',e.get('warm',{}).get('ms_per_step'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_c4_$R.csv \
    python bench.py --only c4 --rows 100000000 --blocks 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu_l4_$R.log 2>&1
ls -la gpurun_out/*$R*
