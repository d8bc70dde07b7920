#!/bin/bash
timeout 600 python scripts/dev/sgs_timing.py 2>&1 | tail -20 | tee gpurun_out/sgs_time.txt
