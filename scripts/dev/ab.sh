#!/bin/bash
# A/B of kernel variants on the GPU box: scripts/dev/ab.sh "<variants>" "<configs>" [targets]
# each variant is a libgskrige.so built from another source state (variants/<name>.so, not tracked)
V=${1:-"base"}; C=${2:-"C2"}; T=${3:-2097152}
mkdir -p gpurun_out
for v in $V; do for c in $C; do
  GSKRIGE_LIB=$PWD/variants/$v.so python bench.py --config $c --targets $T --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/$v $c: /"
done; done | tee -a gpurun_out/ab.log
