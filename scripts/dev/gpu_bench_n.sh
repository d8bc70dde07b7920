#!/bin/bash
# scripts/dev/gpu_bench_n.sh N [steps] [warmup]: the driver's own launch line for N GPUs
N=$1; K=${2:-2}; W=${3:-1}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps $K --warmup $W > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup $W > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.err; python scripts/show_bench.py gpurun_out/bench_n$N.json | cut -c1-600
