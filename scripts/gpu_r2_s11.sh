#!/bin/bash
# Round 2, session 11: what bounds the K = 320 GEGLU / qkv epilogue -- timing switches (math off, stores off, MMA off, loads off).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for dbg in 0 8 16 24 2 10 26 27; do
  WIW_GEMM_DBG=$dbg timeout 200 python scripts/perf_gemm_bound.py > gpurun_out/gb_$dbg.log 2>&1
  echo "== WIW_GEMM_DBG=$dbg"; grep -E "linear|conv3x3 28x72" gpurun_out/gb_$dbg.log
done
echo session done
