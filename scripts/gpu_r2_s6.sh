#!/bin/bash
# Round 2, session 6: ts4 attention (one Q tile, four threads per row, single-pass softmax): parity, A/B, timeline, bounds.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
WIW_ATTN_TS=4 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn_ts4.log 2>&1
echo "attention tests TS=4 exit $?"; tail -3 gpurun_out/t_attn_ts4.log
for cfg in "0 1 0" "4 1 0" "4 2 0" "4 1 1" "4 1 2" "4 1 3"; do
  set -- $cfg
  WIW_ATTN_TS=$1 WIW_ATTN_POLY=$2 WIW_ATTN_DBG=$3 timeout 200 python scripts/perf_attn.py attn nosdpa > gpurun_out/perf_attn_ts$1_poly$2_dbg$3.log 2>&1
  echo "perf_attn TS=$1 POLY=$2 DBG=$3 exit $?"; grep attn_spatial gpurun_out/perf_attn_ts$1_poly$2_dbg$3.log
done
for dbg in 0 3; do
WIW_ATTN_TS=4 WIW_ATTN_DBG=$dbg timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_ts4_dbg$dbg.txt 2>&1; echo "attn_trace dbg=$dbg exit $?"; cat gpurun_out/attn_trace_ts4_dbg$dbg.txt
done
timeout 400 python -m pytest tests/test_00_northstar_forward_gpu.py tests/test_03_model_gpu.py -x -q > gpurun_out/t_model_ts4.log 2>&1
echo "model tests exit $?"; tail -3 gpurun_out/t_model_ts4.log
echo session done
