#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_wpt.log
scripts/dev/ab.sh "pair wpt" "C5" 2097152
GSK_WPT_MIN_K=24 scripts/dev/ab.sh "wpt" "C3a C3b" 2097152
