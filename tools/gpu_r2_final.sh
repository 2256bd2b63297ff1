#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -14 gpurun_out/pytest_r2.log; cat gpurun_out/jit_warm.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed or topn_matches or checksum_matches or exact_layout or desc_table or index_scan" > gpurun_out/memcheck_r2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_r2.log
tail -4 gpurun_out/memcheck_r2.log
bash tools/refresh_profiles_r2.sh r2
