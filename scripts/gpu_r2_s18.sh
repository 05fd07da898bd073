#!/bin/bash
# Round 2, session 18: long determinism soak of the final kernels (bounds the rate of any remaining race).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/stress_gemm_res3.py 6000 2>&1 | tail -3 | tee gpurun_out/soak_res.txt
timeout 900 python scripts/stress_unet.py 700 60 > gpurun_out/soak_model.txt 2>&1; echo "soak model exit $?"; grep "UNET\|VAE:\|checksum" gpurun_out/soak_model.txt | tail -5
echo session done
