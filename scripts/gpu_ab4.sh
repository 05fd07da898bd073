#!/bin/bash
# A/B round 4: epilogue staging depth (WIW_GEMM_OBUFS=2|4), lane-group LayerNorm (WIW_NORM_V=1|3); suite + bench new vs previous settings.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
export NORM_V=3 OBUFS=4
WIW_NORM_V=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" > gpurun_out/t_ln3.log 2>&1 || NORM_V=1
echo "layernorm v3 tests -> WIW_NORM_V=$NORM_V"; tail -3 gpurun_out/t_ln3.log
WIW_GEMM_OBUFS=4 timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm_ob4.log 2>&1 || OBUFS=2
echo "gemm tests -> WIW_GEMM_OBUFS=$OBUFS"; tail -3 gpurun_out/t_gemm_ob4.log
for v in 1 3; do WIW_NORM_V=$v timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_v$v.log 2>&1; done
paste -d'|' gpurun_out/perf_norm_v1.log gpurun_out/perf_norm_v3.log | grep -i layernorm | cut -c1-200
for o in 2 4; do WIW_GEMM_OBUFS=$o timeout 200 python scripts/perf_gemm.py > gpurun_out/perf_gemm_ob$o.log 2>&1; done
paste -d'|' gpurun_out/perf_gemm_ob2.log gpurun_out/perf_gemm_ob4.log | cut -c1-230
export WIW_NORM_V=$NORM_V WIW_GEMM_OBUFS=$OBUFS
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log
WIW_NORM_V=1 WIW_GEMM_OBUFS=2 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-300
echo session done
