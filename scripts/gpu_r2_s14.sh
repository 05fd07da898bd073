#!/bin/bash
# Round 2, session 14: where do the ~400 cycles of the epilogue's accumulator load go?  Phases build, with/without prefetch and MMA.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export WIW_B200_LIB=$PWD/world-in-world_b200/libwiw_b200_phases.so
for x in 0 32 2 34; do
  echo "== EXTRA_DBG=$x (32 = no prefetch of the next chunk's accumulator, 2 = no MMA issue)"
  EXTRA_DBG=$x timeout 300 python scripts/perf_epilogue_phases.py 2>&1 | head -8
done > gpurun_out/epilogue_phases_v2.txt
cat gpurun_out/epilogue_phases_v2.txt
