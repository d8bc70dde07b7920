#!/bin/bash
python -m pytest tests -m gpu -q --tb=no 2>&1 | grep -v "^\.\|^$" | cut -c1-200 | tail -60 | tee gpurun_out/pytest_full.log
