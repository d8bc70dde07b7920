#!/bin/bash
python scripts/dev/extras_time.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | tee gpurun_out/extras_time.log
scripts/dev/gpu_final_bench.sh
