#!/bin/bash
# Round 2, session 17c: share of the exponentials on the FMA pipe (packed polynomial): 0, 1 or 2 of every 4 pairs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for poly in 1 2 0 1 2; do
  echo "== WIW_ATTN_POLY=$poly"; WIW_ATTN_POLY=$poly timeout 300 python scripts/perf_attn.py attn nosdpa 2>&1 | grep "attn_spatial" | head -3
done | tee gpurun_out/attn_poly_ab.txt
echo session done
