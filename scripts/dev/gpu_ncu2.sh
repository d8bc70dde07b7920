#!/bin/bash
# both kernels (search + solve) of a config: scripts/dev/gpu_ncu2.sh <config> <out-name> [targets]
CFG=$1; OUT=$2; T=${3:-262144}
python bench.py --config $CFG --targets $T --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|local_solve" -c 2 -o gpurun_out/$OUT -f \
  python bench.py --config $CFG --targets $T --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$OUT.log 2>&1
echo "ncu $CFG rc=$?"
