#!/bin/bash
# SGS parity first (verbose tail), then the whole GPU suite
timeout 600 python -m pytest tests/test_gpu_sgs.py -q -x 2>&1 | tail -60 | tee gpurun_out/pytest_sgs.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_full.log
