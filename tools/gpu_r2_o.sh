#!/bin/bash
set -x
mkdir -p gpurun_out
R=r2o
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$R.log; tail -4 gpurun_out/smoke_$R.log
B2_JIT=off timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_nojit_$R.log 2>&1; echo "smoke(nojit) rc=$?" >> gpurun_out/smoke_nojit_$R.log; tail -3 gpurun_out/smoke_nojit_$R.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_memcheck_$R.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/smoke_memcheck_$R.log
grep -n 'Invalid\|at \|by thread\|Address\|========= ' gpurun_out/smoke_memcheck_$R.log | head -40
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$R.log
tail -12 gpurun_out/pytest_$R.log
cat gpurun_out/jit_warm.log
