#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
scripts/dev/ab.sh "grp" "C5 C3a C3b C2" 2097152
