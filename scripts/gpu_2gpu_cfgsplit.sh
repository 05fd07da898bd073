#!/bin/bash
# 2-GPU call: CFG split (one branch on 2 GPUs) parity test + latency-mode bench; quick re-check of the final kernel defaults.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_cfg_split_gpu.py -x -q > gpurun_out/t_cfg_split.log 2>&1
echo "cfg split test exit $?"; tail -3 gpurun_out/t_cfg_split.log
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm or attention_spatial" > gpurun_out/t_final_defaults.log 2>&1
echo "final defaults kernel tests exit $?"; tail -2 gpurun_out/t_final_defaults.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --cfg-split > gpurun_out/bench_2gpu_cfgsplit.log 2> gpurun_out/bench_2gpu_cfgsplit.err
echo "bench cfg-split exit $?"; tail -1 gpurun_out/bench_2gpu_cfgsplit.log | cut -c1-500
echo session done
