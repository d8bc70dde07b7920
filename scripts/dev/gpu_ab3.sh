#!/bin/bash
# A/B of library variants: scripts/dev/gpu_ab3.sh "<variants>" "<configs>"
mkdir -p gpurun_out
T=2097152
CFGS=${2:-C2 C3a C5}
for c in $CFGS; do for v in $1; do
  GSKRIGE_LIB=$PWD/variants/$v.so python bench.py --config $c --targets $T --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/$v $c: /" | cut -c1-230
done; done 2>&1 | tee -a gpurun_out/ab3.log
