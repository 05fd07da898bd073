#!/bin/bash
# Round 2, session 2: TS-mode attention kernels (8 and 16 softmax warps) -- parity then same-box A/B against the SS kernel and the
# cuDNN SDPA bar; what bounds conv_gemm_kernel per shape (no-TMA / no-MMA timing switches); full suite; bench.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
for ts in 1 2; do
  WIW_ATTN_TS=$ts timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn_ts$ts.log 2>&1
  echo "attention tests TS=$ts exit $?"; tail -3 gpurun_out/t_attn_ts$ts.log
done
for cfg in "0 1" "1 1" "2 1" "2 2"; do
  set -- $cfg
  WIW_ATTN_TS=$1 WIW_ATTN_POLY=$2 timeout 200 python scripts/perf_attn.py attn nosdpa > gpurun_out/perf_attn_ts$1_poly$2.log 2>&1
  echo "perf_attn TS=$1 POLY=$2 exit $?"; grep attn_spatial gpurun_out/perf_attn_ts$1_poly$2.log
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
for dbg in 0 1 2; do
  WIW_GEMM_DBG=$dbg timeout 200 python scripts/perf_gemm_bound.py > gpurun_out/gemm_bound_dbg$dbg.log 2>&1; echo "gemm_bound dbg=$dbg exit $?"; cat gpurun_out/gemm_bound_dbg$dbg.log
done
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
echo session done
