#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
python scripts/dev/c1_accuracy.py 2>&1 | tail -8
cp geostatssolvers.jl_b200/csrc/libgskrige.so variants/cur.so
scripts/dev/ab.sh "cur" "C2 C5 C3a" 2097152
