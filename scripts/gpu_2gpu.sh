#!/bin/bash
# 2-GPU call (gpurun --gpus 2): sharded worker through the manager's spawn protocol, then the bench launched as the driver launches it.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/gpus2.txt 2>&1
timeout 240 python -m pytest tests/test_sharded_worker_gpu.py -x -q > gpurun_out/t_sharded_gpu.log 2>&1
echo "sharded worker test exit $?"; tail -3 gpurun_out/t_sharded_gpu.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu exit $?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-400
echo session done
