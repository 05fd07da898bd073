#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/stress_gemm_res3.py 800 2>&1 | tail -3 | tee gpurun_out/stress_res3_fixed.txt
timeout 600 python scripts/stress_unet.py 150 12 > gpurun_out/stress_new.txt 2>&1; echo "stress model exit $?"; grep "UNET\|VAE:" gpurun_out/stress_new.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_00_northstar_forward_gpu.py -x -q > gpurun_out/t16.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t16.log
timeout 300 python scripts/perf_unet.py --vae 2>&1 | grep "forward 3\|decode 2"
echo session done
