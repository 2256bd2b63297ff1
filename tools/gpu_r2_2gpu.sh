#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --rows 300000000 --blocks 8 --sub-rows 50000000 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_$R.json 2> gpurun_out/bench_2gpu_$R.err
tail -c 1500 gpurun_out/bench_2gpu_$R.json; tail -5 gpurun_out/bench_2gpu_$R.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_reference_2gpu_$R.json 2> gpurun_out/bench_reference_2gpu_$R.err
tail -c 600 gpurun_out/bench_reference_2gpu_$R.json; tail -3 gpurun_out/bench_reference_2gpu_$R.err
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist_$R.log 2>&1; tail -5 gpurun_out/pytest_dist_$R.log
