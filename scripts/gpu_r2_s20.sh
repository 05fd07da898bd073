#!/bin/bash
# Round 2, session 20: sanity of the rebuilt library at the final HEAD (sources identical to the validated commit).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_00_northstar_forward_gpu.py tests/test_01_northstar_rollout_gpu.py -x -q > gpurun_out/t20.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t20.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 2 --warmup 3 2> gpurun_out/bench20.err | tail -1 | cut -c1-400
echo session done
