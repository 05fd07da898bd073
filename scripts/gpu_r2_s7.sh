#!/bin/bash
# Round 2, session 7: ss2 attention (smem operands + two threads per row): parity and same-box A/B.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
WIW_ATTN_TS=6 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn_ts6.log 2>&1
echo "attention tests TS=6 exit $?"; tail -3 gpurun_out/t_attn_ts6.log
for cfg in "0 1" "6 1" "5 1" "0 1" "6 1"; do
  set -- $cfg
  WIW_ATTN_TS=$1 WIW_ATTN_POLY=$2 timeout 200 python scripts/perf_attn.py attn nosdpa > gpurun_out/perf_attn_ts$1_poly$2.log 2>&1
  echo "perf_attn TS=$1 POLY=$2 exit $?"; grep attn_spatial gpurun_out/perf_attn_ts$1_poly$2.log
done
echo session done
