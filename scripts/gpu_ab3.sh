#!/bin/bash
# A/B round 3: tcgen05.mma cost table, GELU formulation (WIW_GELU=1 A&S erf | 2 tanh-fit erf), norm-ring race fix, ncu of the norm kernels,
# full suite + bench with the new defaults and a same-box bench with the previous settings.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
timeout 120 scripts/mma_bench.bin > gpurun_out/mma_cost.txt 2>&1; echo "mma_bench exit $?"
for g in 1 2; do
    WIW_GELU=$g timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q -k geglu > gpurun_out/t_geglu$g.log 2>&1; echo "geglu mode $g tests exit $?"
    WIW_GELU=$g timeout 200 python scripts/perf_gemm.py geglu > gpurun_out/perf_gemm_gelu$g.log 2>&1
    grep geglu gpurun_out/perf_gemm_gelu$g.log
    WIW_GELU=$g timeout 400 python -m pytest tests/test_model_full_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/t_model_gelu$g.log 2>&1; echo "model tests gelu $g exit $?"
done
cat gpurun_out/parity.jsonl
for i in 1 2; do
    WIW_NORM_V=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm" > gpurun_out/t_norm_v2_$i.log 2>&1; echo "norm v2 run $i exit $?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'layernorm|gn_|concat' -o gpurun_out/full_norm -f python scripts/ncu_norm.py > gpurun_out/full_norm.log 2>&1
echo "ncu norm exit $?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log
WIW_GELU=1 WIW_ATTN_POLY=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-300
echo session done
