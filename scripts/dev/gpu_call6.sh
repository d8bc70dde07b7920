#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_wpt2.log
scripts/dev/ab.sh "wpt wpt2" "C5" 2097152
scripts/dev/ab.sh "wpt2" "C2 C3a" 2097152
