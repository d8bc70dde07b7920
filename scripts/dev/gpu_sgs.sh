#!/bin/bash
# SGS parity first, then the timings
timeout 600 python -m pytest tests/test_gpu_sgs.py -q -x 2>&1 | tail -40 | tee gpurun_out/pytest_sgs.log
timeout 600 python scripts/dev/sgs_timing.py 2>&1 | tail -20 | tee gpurun_out/sgs_time.txt
