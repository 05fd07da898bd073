#!/bin/bash
# Round 2, session 21: ncu launch list of one VAE decode with the final library (evidence only; no product change).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_vae.csv \
    python scripts/ncu_vae.py 2 > gpurun_out/launches_vae.log 2>&1
echo "vae launch list exit $?"; wc -l gpurun_out/launches_vae.csv; tail -2 gpurun_out/launches_vae.log
echo session done
