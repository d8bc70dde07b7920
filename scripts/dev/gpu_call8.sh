#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_wpt3.log
scripts/dev/ab.sh "wpt2 wpt3" "C5" 2097152
