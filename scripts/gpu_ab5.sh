#!/bin/bash
# A/B round 5: M-tiles spanning images / frames (WIW_BOX3=0|1) for the 9-row and 18-row feature maps; suite + bench new vs previous.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
export BOX3=1
WIW_BOX3=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm_box3.log 2>&1 || BOX3=0
echo "gemm tests -> WIW_BOX3=$BOX3"; tail -3 gpurun_out/t_gemm_box3.log
export WIW_BOX3=$BOX3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log
cp gpurun_out/gemm_profile.json gpurun_out/gemm_profile_box3.json
WIW_BOX3=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --dump-gemm-profile > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-300
echo session done
