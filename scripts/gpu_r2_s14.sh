#!/bin/bash
# Round 2, session 14: does the slow TMEM->register path of the epilogue depend on TMA loads (bit 0 off) or MMA issue (bit 1 off)?
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export WIW_B200_LIB=$PWD/world-in-world_b200/libwiw_b200_phases.so
for x in 0; do
  echo "== EXTRA_DBG=$x (1 = no TMA operand loads, 2 = no MMA issue)"
  EXTRA_DBG=$x timeout 300 python scripts/perf_epilogue_phases.py 2>&1 | head -8
done > gpurun_out/epilogue_phases_v5.txt
cat gpurun_out/epilogue_phases_v5.txt
