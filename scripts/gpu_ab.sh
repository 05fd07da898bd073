#!/bin/bash
# Same-box A/B of one development switch (boxes of the pod differ by ~5 %, so a change is only accepted on a pair measured
# back to back in ONE gpurun call):
#   gpurun --timeout 900 -- 'bash scripts/gpu_ab.sh WIW_GELU 1 2 "tests/test_gemm_gpu.py -k geglu" "scripts/perf_gemm.py geglu"'
# 1. parity tests with VAR=NEW (on failure everything after runs with OLD only), 2. the micro-benchmark with OLD and NEW side by side,
# 3. model parity tests (margins land in gpurun_out/parity.jsonl), 4. bench with NEW then OLD (device-timed value only).
set -u
VAR=$1; OLD=$2; NEW=$3; TESTS=${4:-}; PERF=${5:-}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
USE=$NEW
if [ -n "$TESTS" ]; then
    env $VAR=$NEW timeout 300 python -m pytest $TESTS -x -q > gpurun_out/ab_tests.log 2>&1 || USE=$OLD
    echo "parity tests with $VAR=$NEW -> using $VAR=$USE"; tail -3 gpurun_out/ab_tests.log
fi
if [ -n "$PERF" ]; then
    for v in $OLD $NEW; do env $VAR=$v timeout 200 python $PERF > gpurun_out/ab_perf_$v.log 2>&1; done
    paste -d'|' gpurun_out/ab_perf_$OLD.log gpurun_out/ab_perf_$NEW.log | cut -c1-230
fi
env $VAR=$USE timeout 400 python -m pytest tests/test_model_full_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/ab_model.log 2>&1
echo "model tests ($VAR=$USE) exit $?"; cat gpurun_out/parity.jsonl
env $VAR=$USE timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench $VAR=$USE exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
env $VAR=$OLD timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench $VAR=$OLD exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-200
echo session done
