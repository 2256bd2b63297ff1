#!/bin/bash
# 2 GPUs: the driver's launch line for N=2 (one rank per GPU over NCCL), full default workload
set -x
mkdir -p gpurun_out
R=r2k
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_$R.json 2> gpurun_out/bench_n2_$R.err
tail -c 1500 gpurun_out/bench_n2_$R.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_n2_r2k.json').read().strip().splitlines()[-1])
e=d['e2e']
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'merge',d.get('merge_ms'))
print({k:v for k,v in e.items() if k not in('flat','warm','source','note')})
print('flat',e.get('flat',{}).get('ms_per_step'),'warm',e.get('warm',{}).get('ms_per_step'))
for s in d.get('sub',[]): print(s['workload'][:20], s['value'], s['ms_per_step'], s['roofline']['frac'], s.get('merge_ms'))
P
