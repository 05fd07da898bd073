#!/bin/bash
# bench.py exactly as the driver launches it at N=2, final library (evidence only).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 \
    > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "bench N=2 exit $?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-200
echo session done
