#!/bin/bash
# Round 2, session 3: MN-major B cost, pipeline timeline of the TS attention kernel, new GPU tests, launch list of the VAE decode.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 ./scripts/mma_bench2.bin > gpurun_out/mma_bench2.txt 2>&1; echo "mma_bench2 exit $?"; grep -E "MN-major|TS M=128 N= (64|80)" gpurun_out/mma_bench2.txt
timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "attn_trace exit $?"; cat gpurun_out/attn_trace.txt
timeout 600 python -m pytest tests/test_05_closed_loop_gpu.py tests/test_04_worker_gpu.py tests/test_kernels_gpu.py -x -q -k "closed_loop or worker or resize" > gpurun_out/t_new.log 2>&1
echo "new tests exit $?"; tail -5 gpurun_out/t_new.log; cat gpurun_out/closed_loop.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_vae.csv \
    python scripts/ncu_vae.py 2 > gpurun_out/launches_vae.log 2>&1
echo "vae launch list exit $?"; wc -l gpurun_out/launches_vae.csv
echo session done
