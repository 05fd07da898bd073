#!/bin/bash
# 2-GPU call: latency-mode bench (one branch on 2 GPUs, CFG rows split).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --cfg-split > gpurun_out/bench_2gpu_cfgsplit.log 2> gpurun_out/bench_2gpu_cfgsplit.err
echo "bench cfg-split exit $?"; tail -1 gpurun_out/bench_2gpu_cfgsplit.log | cut -c1-600
