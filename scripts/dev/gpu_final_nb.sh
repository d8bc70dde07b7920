#!/bin/bash
# end-of-round check without the ncu captures: GPU suite, smoke, the driver's two N = 1 bench invocations
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke_final.log
bash scripts/dev/gpu_final_bench.sh 2>&1 | tee gpurun_out/final_bench.log
