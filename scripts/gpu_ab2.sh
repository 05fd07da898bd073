#!/bin/bash
# A/B round 2: streaming norm kernels (WIW_NORM_V), polynomial-exp share in attention (WIW_ATTN_POLY), then suite + bench.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export NORM_V=2
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm" > gpurun_out/t_norm.log 2>&1 || NORM_V=1
echo "norm tests -> WIW_NORM_V=$NORM_V"; tail -3 gpurun_out/t_norm.log
WIW_NORM_V=1 timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_v1.log 2>&1; echo "perf_norm v1 exit $?"
if [ "$NORM_V" = 2 ]; then
    WIW_NORM_V=2 timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_v2.log 2>&1; echo "perf_norm v2 exit $?"
fi
paste -d'|' gpurun_out/perf_norm_v1.log gpurun_out/perf_norm_v2.log 2>/dev/null | cut -c1-200
export WIW_NORM_V=$NORM_V
for p in 0 1 2; do
    WIW_ATTN_POLY=$p timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention_spatial" > gpurun_out/t_attn_poly$p.log 2>&1
    echo "attn poly $p tests exit $?"
    WIW_ATTN_POLY=$p timeout 200 python scripts/perf_attn.py attn > gpurun_out/perf_attn_poly$p.log 2>&1
    grep attn_spatial gpurun_out/perf_attn_poly$p.log
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log
echo session done
