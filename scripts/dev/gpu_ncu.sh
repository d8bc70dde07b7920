#!/bin/bash
# scripts/dev/gpu_ncu.sh <config> <kernel-regex> <out-name> [targets]
CFG=$1; RE=$2; OUT=$3; T=${4:-262144}
python bench.py --config $CFG --targets $T --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$RE -c 1 -o gpurun_out/$OUT -f \
  python bench.py --config $CFG --targets $T --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$OUT.log 2>&1
echo "ncu rc=$?"
