#!/bin/bash
# Round 2, session 22: LEAN = 2 (lean epilogue with the aux operand, separate instantiation) -- tests and A/B via WIW_GEMM_LEAN_AUX.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -1
{
for v in 1 0 1 0; do
  echo "== WIW_GEMM_LEAN_AUX=$v"; WIW_GEMM_LEAN_AUX=$v timeout 120 python scripts/perf_gemm_aux.py 2>&1 | tail -3
done
for v in 1 0; do
  echo "== WIW_GEMM_LEAN_AUX=$v"; WIW_GEMM_LEAN_AUX=$v timeout 200 python scripts/perf_unet.py 2>&1 | grep "forward 2\|forward 3"
done
} | tee gpurun_out/lean_aux_ab.txt
echo session done
