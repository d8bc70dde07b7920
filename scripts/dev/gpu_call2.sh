#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_pair.log
scripts/dev/ab.sh "p05 pair" "C5" 2097152
scripts/dev/ab.sh "pair" "C3a" 2097152
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_v2.err
