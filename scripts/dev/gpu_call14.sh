#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_full.log
python scripts/dev/multi_time.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute"
scripts/dev/gpu_bench_n.sh 2
