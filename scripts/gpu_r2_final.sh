#!/bin/bash
# Round 2 evidence run on the final HEAD: GPU suite, smoke, bench (both arms), ncu launch list of one UNet forward and of the VAE decode,
# ncu --set full of the hot kernels (-> scripts/ncu_traffic.py on the CPU box).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/tests_gpu.log; cat gpurun_out/parity.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --dump-gemm-profile > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-600; tail -3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err
echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_gemm|attn_spatial|temporal_attn|gn_|layernorm|concat_kernel|upsample2x|emb_silu' --csv --log-file gpurun_out/launches_unet.csv \
    python scripts/ncu_unet.py 2 > gpurun_out/launches_unet.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_unet.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|attn_spatial' -o gpurun_out/full_kernels -f \
    python scripts/ncu_kernels.py all > gpurun_out/full_kernels.log 2>&1
echo "ncu full (gemm, attention) exit $?"; ls -la gpurun_out/*.ncu-rep
echo session done
