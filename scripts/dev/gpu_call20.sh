#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
echo "--- debug-bounds build (assert on every lane-computed pool index) ---"
GSKRIGE_LIB=$PWD/variants/bounds.so python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_bounds.log
python scripts/dev/plan_time.py 2>&1 | tail -5
python - <<'P'
import sys, time
sys.path.insert(0, '.')
import gskrige
c = gskrige.Context(0)
spec = gskrige.synth.config_spec("C4")
for _ in range(2):
    t0 = time.perf_counter(); c.plan(spec); c.synchronize(); print("C4 plan: %.1f ms (device %.1f ms)" % (1e3*(time.perf_counter()-t0), c.timing()["ms_plan"]))
P
