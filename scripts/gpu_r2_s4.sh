#!/bin/bash
# Round 2, session 4: double-buffered-score TS attention (ts3): parity, same-box A/B, pipeline timeline.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
WIW_ATTN_TS=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn_ts3.log 2>&1
echo "attention tests TS=3 exit $?"; tail -3 gpurun_out/t_attn_ts3.log
for cfg in "0 1" "2 1" "3 1" "3 2"; do
  set -- $cfg
  WIW_ATTN_TS=$1 WIW_ATTN_POLY=$2 timeout 200 python scripts/perf_attn.py attn nosdpa > gpurun_out/perf_attn_ts$1_poly$2.log 2>&1
  echo "perf_attn TS=$1 POLY=$2 exit $?"; grep attn_spatial gpurun_out/perf_attn_ts$1_poly$2.log
done
WIW_ATTN_TS=3 timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_ts3.txt 2>&1; echo "attn_trace exit $?"; cat gpurun_out/attn_trace_ts3.txt
timeout 400 python -m pytest tests/test_00_northstar_forward_gpu.py tests/test_03_model_gpu.py -x -q > gpurun_out/t_model_ts3.log 2>&1
echo "model tests exit $?"; tail -3 gpurun_out/t_model_ts3.log
echo session done
