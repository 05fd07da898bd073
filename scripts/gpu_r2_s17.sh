#!/bin/bash
# Round 2, session 17: attention kernel, mbarrier-timeout diagnostic as a call (new) vs inline trap vs the library before the change (lean1).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in new inline lean1 new; do
  if [ $lib = lean1 ]; then export WIW_B200_LIB=$PWD/ab/libwiw_lean1.so; elif [ $lib = inline ]; then export WIW_B200_LIB=$PWD/ab/libwiw_inline_trap.so; else unset WIW_B200_LIB; fi
  echo "== $lib"; timeout 300 python scripts/perf_attn.py attn nosdpa 2>&1 | grep "attn_spatial"
done | tee gpurun_out/attn_trap_ab.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -1
echo session done
