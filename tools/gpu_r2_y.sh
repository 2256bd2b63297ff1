#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r2.json 2> gpurun_out/bench_reference_r2.err
tail -c 300 gpurun_out/bench_r2.err; python tools/r2_numbers.py gpurun_out/bench_r2.json
