#!/bin/bash
# Round 2, session 5: which side bounds the ts3 attention kernel -- softmax-only / MMA-only timing switches + timelines.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for dbg in 0 1 2 3; do
  WIW_ATTN_TS=3 WIW_ATTN_DBG=$dbg timeout 200 python scripts/perf_attn.py attn nosdpa > gpurun_out/perf_attn_ts3_dbg$dbg.log 2>&1
  echo "perf_attn TS=3 DBG=$dbg exit $?"; grep "attn_spatial imgs=28 S=9216 C=320" gpurun_out/perf_attn_ts3_dbg$dbg.log
  WIW_ATTN_TS=3 WIW_ATTN_DBG=$dbg timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_ts3_dbg$dbg.txt 2>&1; sed -n 1,4p gpurun_out/attn_trace_ts3_dbg$dbg.txt; tail -1 gpurun_out/attn_trace_ts3_dbg$dbg.txt
done
echo session done
