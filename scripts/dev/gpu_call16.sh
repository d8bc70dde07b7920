#!/bin/bash
GSK_WPT2_MIN_K=1 GSK_NO_SMALL_KERNEL=1 scripts/dev/ab.sh "dev" "C2" 1000000
scripts/dev/ab.sh "dev" "C2" 1000000
python bench.py --config C5 --targets 2097152 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"search_kernel" -c 1 -o gpurun_out/r02_c5_search -f python bench.py --config C5 --targets 2097152 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c5_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|local_solve" -c 2 -o gpurun_out/r02_c2 -f python bench.py --config C2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"local_solve" -c 1 -o gpurun_out/r02_c3a_wpt2 -f python bench.py --config C3a --targets 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c3a_wpt2.log 2>&1
echo done
