#!/bin/bash
# full-size C5 (the default bench workload), library variants and chunk sizes
mkdir -p gpurun_out
run() { local name=$1 lib=$2; shift 2
  env "$@" GSKRIGE_LIB=$PWD/variants/$lib.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>&1 | python scripts/show_bench.py - | sed "s/^/$name: /" | cut -c1-260; }
{
run prod prod X=1
run c32 c32 X=1
run c32_chunk22 c32 GSK_CHUNK_LOG2=22
run c32_chunk21 c32 GSK_CHUNK_LOG2=21
for c in C2 C3a; do
GSKRIGE_LIB=$PWD/variants/prod.so python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/prod $c: /" | cut -c1-260
GSKRIGE_LIB=$PWD/variants/c32.so python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/c32 $c: /" | cut -c1-260
GSK_CHUNK_LOG2=22 GSKRIGE_LIB=$PWD/variants/c32.so python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python scripts/show_bench.py - | sed "s/^/c32_chunk22 $c: /" | cut -c1-260
done
} 2>&1 | tee gpurun_out/ab4.log
