#!/bin/bash
# first-bound (density threshold) in the search: GPU suite on the in-tree library, then A/B and a sweep of the bound
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/pytest_th.log
T=2097152
run() { # name lib config env...
  local name=$1 lib=$2 cfg=$3; shift 3
  env "$@" GSKRIGE_LIB=$PWD/variants/$lib.so python bench.py --config $cfg --targets $T --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/$name $cfg: /" | cut -c1-260
}
{
for c in C2 C3a C5; do
  run ck ck $c X=1
  run th_off th $c GSK_THRESH_A=0
  run th th $c X=1
  run th_1.3_8 th $c GSK_THRESH_A=1.3 GSK_THRESH_B=8
  run th_1.8_12 th $c GSK_THRESH_A=1.8 GSK_THRESH_B=12
  run th_1.25_16 th $c GSK_THRESH_A=1.25 GSK_THRESH_B=16
done
} 2>&1 | tee gpurun_out/ab_th.log
