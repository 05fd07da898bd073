#!/bin/bash
# A/B round 8: single-pass softmax in the spatial attention kernel (WIW_ATTN_ONEPASS=0|1).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
export OP=1
WIW_ATTN_ONEPASS=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention_spatial" > gpurun_out/t_attn_onepass.log 2>&1 || OP=0
echo "attention tests -> WIW_ATTN_ONEPASS=$OP"; tail -3 gpurun_out/t_attn_onepass.log
for o in 0 1; do WIW_ATTN_ONEPASS=$o timeout 200 python scripts/perf_attn.py attn > gpurun_out/perf_attn_onepass$o.log 2>&1; grep attn_spatial gpurun_out/perf_attn_onepass$o.log; done
export WIW_ATTN_ONEPASS=$OP
timeout 400 python -m pytest tests/test_model_full_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/t_model_onepass.log 2>&1; echo "model tests exit $?"; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
WIW_ATTN_ONEPASS=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-200
echo session done
