#!/bin/bash
export GSKRIGE_LIB=$PWD/variants/dev.so
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_wpt2k.log
unset GSKRIGE_LIB
scripts/dev/ab.sh "dev" "C3a C3b C5" 2097152
GSK_WPT2_MIN_K=99 scripts/dev/ab.sh "dev" "C3a C3b" 2097152
