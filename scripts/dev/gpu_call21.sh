#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
echo "--- debug-bounds build (assert on every lane-computed pool index) ---"
GSKRIGE_LIB=$PWD/variants/bounds.so python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_bounds.log
scripts/dev/ab.sh "cb4" "C2 C5" 2097152
