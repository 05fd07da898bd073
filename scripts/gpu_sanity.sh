#!/bin/bash
# Short end-of-session sanity call on the final HEAD: the bench exactly as the driver launches it (device-timed part), then the GPU suite
# without its two slow model-scale tests.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 100 python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_sanity.log 2> gpurun_out/bench_sanity.err
echo "bench exit $?"; tail -1 gpurun_out/bench_sanity.log | cut -c1-260
timeout 100 python -m pytest tests -m gpu -x -q --deselect tests/test_model_full_gpu.py --deselect tests/test_eager_baseline_gpu.py > gpurun_out/tests_sanity.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/tests_sanity.log
