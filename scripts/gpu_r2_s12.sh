#!/bin/bash
# Round 2, session 12: CLIP engine on the GPU, then the whole suite and the bench on the new front-end.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_06_clip_gpu.py -x -q > gpurun_out/t_clip.log 2>&1
echo "clip tests exit $?"; tail -5 gpurun_out/t_clip.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_06_clip_gpu.py > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
echo session done
