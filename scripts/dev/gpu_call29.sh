#!/bin/bash
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^\.\|^$" | cut -c1-200 | tail -25 | tee gpurun_out/pytest_full.log
scripts/dev/ab.sh "spl" "C5 C3a C3b C2" 2097152
