#!/bin/bash
# A/B round 6: 8 residual slots (WIW_GEMM_RSLOTS=4|8) + new multi-image tile tests; bench new vs previous on the same box.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
export RS=8
WIW_GEMM_RSLOTS=8 timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm_rs8.log 2>&1 || RS=4
echo "gemm tests -> WIW_GEMM_RSLOTS=$RS"; tail -3 gpurun_out/t_gemm_rs8.log
for r in 4 8; do WIW_GEMM_RSLOTS=$r timeout 200 python scripts/perf_gemm.py res > gpurun_out/perf_gemm_rs$r.log 2>&1; done
paste -d'|' gpurun_out/perf_gemm_rs4.log gpurun_out/perf_gemm_rs8.log | cut -c1-230
export WIW_GEMM_RSLOTS=$RS
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
cp gpurun_out/gemm_profile.json gpurun_out/gemm_profile_rs8.json
WIW_GEMM_RSLOTS=4 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --dump-gemm-profile > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-200
echo session done
