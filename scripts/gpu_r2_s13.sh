#!/bin/bash
# Round 2, session 13: bias staged in shared memory by the GEMM epilogue -- parity, phase table, per-shape and whole-model A/B
# against the previous library (ab/libwiw_prev_epi.so).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_00_northstar_forward_gpu.py -x -q > gpurun_out/t13.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/t13.log
for lib in new lean1; do
  if [ $lib = prev ]; then export WIW_B200_LIB=$PWD/ab/libwiw_prev_epi.so; elif [ $lib = lean1 ]; then export WIW_B200_LIB=$PWD/ab/libwiw_lean1.so; else unset WIW_B200_LIB; fi
  timeout 300 python scripts/perf_gemm.py > gpurun_out/perf_gemm_$lib.txt 2>&1
  timeout 300 python scripts/perf_unet.py > gpurun_out/perf_unet_$lib.txt 2>&1
  timeout 300 python scripts/perf_unet.py --vae > gpurun_out/perf_vae_$lib.txt 2>&1
  echo "== $lib"; tail -2 gpurun_out/perf_unet_$lib.txt; tail -2 gpurun_out/perf_vae_$lib.txt
done
paste -d"|" <(grep -o "'name': '[^']*', 'ms': [0-9.]*" gpurun_out/perf_gemm_new.txt) <(grep -o "'ms': [0-9.]*" gpurun_out/perf_gemm_lean1.txt)
WIW_B200_LIB=$PWD/world-in-world_b200/libwiw_b200_phases.so timeout 300 python scripts/perf_epilogue_phases.py > gpurun_out/epilogue_phases_v4.txt 2>&1; cat gpurun_out/epilogue_phases_v4.txt
echo session done
