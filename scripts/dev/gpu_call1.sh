#!/bin/bash
# round 2, call 1: parity of the fully patched tree + A/B of each patch level
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_p05.log
scripts/dev/ab.sh "base p01 p02 p03 p04 p05" "C2" 1000000
scripts/dev/ab.sh "base p01 p04" "C3a C3b C5" 2097152
