#!/bin/bash
python -m pytest tests -m gpu -q --tb=no 2>&1 | grep -v "^\.\|^$" | cut -c1-200 | tail -20 | tee gpurun_out/pytest_full.log
scripts/dev/ab.sh "h4" "C5 C3a" 2097152
