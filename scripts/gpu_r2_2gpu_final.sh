#!/bin/bash
# Round 2, two-GPU session: the 2-GPU tests the driver's 1-GPU box skips, and bench.py exactly as the driver launches it at N=2
# (sharded-worker e2e + cfg_split_one_branch_ms keys).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
timeout 600 python -m pytest tests/test_cfg_split_gpu.py tests/test_sharded_worker_gpu.py tests/test_03_model_gpu.py -x -q -k "cfg_split or sharded or vae_encoder" > gpurun_out/t_2gpu.log 2>&1
echo "2-gpu tests exit $?"; tail -4 gpurun_out/t_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 \
    > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "bench N=2 exit $?"; tail -1 gpurun_out/bench_2gpu.log; tail -5 gpurun_out/bench_2gpu.err
echo session done
