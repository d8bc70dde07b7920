#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
scripts/dev/ab.sh "t3s" "C2" 1000000
