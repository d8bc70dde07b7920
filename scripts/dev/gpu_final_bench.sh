#!/bin/bash
# the driver's own invocations at N = 1: reference arm first, then ours
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref_n1.json 2> gpurun_out/r02_ref_n1.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "ours rc=$?"
tail -c 400 gpurun_out/r02_bench_n1.err
python scripts/show_bench.py gpurun_out/r02_ref_n1.json | cut -c1-300
python scripts/show_bench.py gpurun_out/r02_bench_n1.json | cut -c1-500
