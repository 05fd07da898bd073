#!/bin/bash
# ncu launch list of one full-size UNet forward (second of two eager forwards), this library's kernels only.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'conv_gemm|attn_spatial|temporal_attn|gn_|layernorm|concat_kernel|upsample2x|emb_silu' --csv \
    --log-file gpurun_out/launches_unet.csv python scripts/ncu_unet.py 2 > gpurun_out/launches_unet.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_unet.csv
