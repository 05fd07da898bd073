#!/bin/bash
# Round 2, session 1: GPU suite at HEAD (golden tests first), micro-benchmarks that settle the MMA cost model, SDPA bar, bench.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 120 ./scripts/mma_bench2.bin > gpurun_out/mma_bench2.txt 2>&1; echo "mma_bench2 exit $?"; cat gpurun_out/mma_bench2.txt
timeout 300 python scripts/perf_attn.py attn > gpurun_out/perf_attn.log 2>&1; echo "perf_attn exit $?"; cat gpurun_out/perf_attn.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
echo session done
