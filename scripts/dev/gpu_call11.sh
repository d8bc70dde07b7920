#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_full.log
scripts/dev/gpu_ncu2.sh C3a r02_c3a
scripts/dev/gpu_ncu2.sh C5 r02_c5_v3
