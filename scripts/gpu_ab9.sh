#!/bin/bash
# A/B round 9: GroupNorm finalize fused into the stats kernel's last CTA (WIW_GN_FUSEFIN=0|1).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
export FF=1
WIW_GN_FUSEFIN=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm" > gpurun_out/t_gn_ff.log 2>&1 || FF=0
echo "groupnorm tests -> WIW_GN_FUSEFIN=$FF"; tail -3 gpurun_out/t_gn_ff.log
for o in 0 1; do WIW_GN_FUSEFIN=$o timeout 200 python scripts/perf_attn.py norm > gpurun_out/perf_norm_ff$o.log 2>&1; done
paste -d'|' gpurun_out/perf_norm_ff0.log gpurun_out/perf_norm_ff1.log | grep -i groupnorm | cut -c1-200
export WIW_GN_FUSEFIN=$FF
timeout 400 python -m pytest tests/test_model_full_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/t_model_ff.log 2>&1; echo "model tests exit $?"; cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
WIW_GN_FUSEFIN=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_prev_settings.log 2>> gpurun_out/bench.err
echo "bench (previous settings) exit $?"; tail -1 gpurun_out/bench_prev_settings.log | cut -c1-200
echo session done
