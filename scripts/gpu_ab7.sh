#!/bin/bash
# A/B round 7: wave-fitted grids for GroupNorm / LayerNorm (WIW_WAVEFIT=0|1); then same-box bench A/B.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export WF=1
WIW_WAVEFIT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm" > gpurun_out/t_norm_wf.log 2>&1 || WF=0
echo "norm tests -> WIW_WAVEFIT=$WF"; tail -3 gpurun_out/t_norm_wf.log
for w in 0 1; do WIW_WAVEFIT=$w timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_wf$w.log 2>&1; done
paste -d'|' gpurun_out/perf_norm_wf0.log gpurun_out/perf_norm_wf1.log | cut -c1-200
WIW_WAVEFIT=$WF timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
WIW_WAVEFIT=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-200
echo session done
