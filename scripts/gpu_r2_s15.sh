#!/bin/bash
# Round 2, session 15: attention kernel with setmaxnreg (softmax warps 232 registers) vs the previous library; then the whole GPU suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" > gpurun_out/t15.log 2>&1
echo "attention tests exit $?"; tail -2 gpurun_out/t15.log
for lib in new lean1; do
  if [ $lib = lean1 ]; then export WIW_B200_LIB=$PWD/ab/libwiw_lean1.so; else unset WIW_B200_LIB; fi
  timeout 300 python scripts/perf_attn.py nosdpa > gpurun_out/perf_attn_$lib.txt 2>&1
  echo "== $lib"; grep -i "tflop\|ms" gpurun_out/perf_attn_$lib.txt | tail -8
done
unset WIW_B200_LIB
rm -f gpurun_out/parity.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1600; tail -3 gpurun_out/bench.err
echo session done
