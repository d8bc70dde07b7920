#!/bin/bash
# compact-key search: GPU suite on the in-tree library, then A/B against the previous library (variants/base.so)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/pytest_ck.log
T=2097152
run() { # name lib config env...
  local name=$1 lib=$2 cfg=$3; shift 3
  env "$@" GSKRIGE_LIB=$PWD/variants/$lib.so python bench.py --config $cfg --targets $T --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/$name $cfg: /"
}
{
for c in C2 C3a C5; do
  run base base $c X=1
  run ck ck $c X=1
done
run ck_scap128 ck C3a GSK_SCAP=128
run ck_scap512 ck C3a GSK_SCAP=512
run ck_scap128 ck C5 GSK_SCAP=128
run ck_scap512 ck C5 GSK_SCAP=512
run ck_scap256 ck C2 GSK_SCAP=256
run ck_heap16 ck C2 GSK_HEAP_MIN_K=16
run ck_ins40 ck C3a GSK_HEAP_MIN_K=40
} 2>&1 | tee gpurun_out/ab_ck.log
