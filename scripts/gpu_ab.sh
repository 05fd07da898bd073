#!/bin/bash
# One gpurun call for an A/B round: validate the new kernel variants first, fall back per feature if a parity test
# fails, time old vs new in the same box, then run the full GPU suite and the bench with the chosen variants.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_ab.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export NORM_V=2 GEMM_EPI=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm" > gpurun_out/t_norm.log 2>&1 || NORM_V=1
echo "norm tests -> WIW_NORM_V=$NORM_V"; tail -3 gpurun_out/t_norm.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm.log 2>&1 || GEMM_EPI=0
echo "gemm tests -> WIW_GEMM_EPI=$GEMM_EPI"; tail -3 gpurun_out/t_gemm.log

WIW_GEMM_EPI=0 timeout 200 python scripts/perf_gemm.py > gpurun_out/perf_gemm_epi0.log 2>&1; echo "perf_gemm epi0 exit $?"
if [ "$GEMM_EPI" = 1 ]; then
    WIW_GEMM_EPI=1 timeout 200 python scripts/perf_gemm.py > gpurun_out/perf_gemm_epi1.log 2>&1; echo "perf_gemm epi1 exit $?"
fi
WIW_NORM_V=1 timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_v1.log 2>&1; echo "perf_norm v1 exit $?"
if [ "$NORM_V" = 2 ]; then
    WIW_NORM_V=2 timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_v2.log 2>&1; echo "perf_norm v2 exit $?"
fi
paste -d'|' gpurun_out/perf_gemm_epi0.log gpurun_out/perf_gemm_epi1.log 2>/dev/null | cut -c1-230
paste -d'|' gpurun_out/perf_norm_v1.log gpurun_out/perf_norm_v2.log 2>/dev/null | cut -c1-200

export WIW_NORM_V=$NORM_V WIW_GEMM_EPI=$GEMM_EPI
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log
echo session done
