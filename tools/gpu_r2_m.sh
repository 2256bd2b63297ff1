#!/bin/bash
mkdir -p gpurun_out
R=r2m
for H in on off on off; do
  if [ $H = off ]; then export B2_NO_SLOW_HINTS=1; else unset B2_NO_SLOW_HINTS; fi
  timeout 300 python bench.py --rows 500000000 --blocks 8 --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-sub > gpurun_out/ab_c3_$H.json 2>> gpurun_out/ab_$R.err
  timeout 300 python bench.py --only c5 --rows 100000000 --blocks 8 --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity --no-sub > gpurun_out/ab_c5_$H.json 2>> gpurun_out/ab_$R.err
  python - <<P
import json
for w in ("c3","c5"):
    d=json.load(open("gpurun_out/ab_%s_$H.json" % w))
    print("$H", w, "ms", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
P
done
tail -5 gpurun_out/ab_$R.err
