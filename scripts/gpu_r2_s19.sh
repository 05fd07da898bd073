#!/bin/bash
# Round 2, session 19: lean epilogue with aux operand, GroupNorm finalize fused into apply, attention stagger sweep -- tests + same-box A/B.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_00_northstar_forward_gpu.py -x -q > gpurun_out/t19.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t19.log
{
for cfg in new nofuse prev new; do
  unset WIW_B200_LIB WIW_GN_FUSED_FINALIZE
  if [ $cfg = prev ]; then export WIW_B200_LIB=$PWD/ab/libwiw_attn_packed1.so; fi
  if [ $cfg = nofuse ]; then export WIW_GN_FUSED_FINALIZE=0; fi
  echo "== $cfg"; timeout 300 python scripts/perf_unet.py --vae 2>&1 | grep "forward 2\|forward 3\|decode 2"
done
unset WIW_B200_LIB WIW_GN_FUSED_FINALIZE
for st in 600 1000 1400; do
  echo "== WIW_ATTN_STAGGER=$st"; WIW_ATTN_STAGGER=$st timeout 300 python scripts/perf_attn.py attn nosdpa 2>&1 | grep "attn_spatial" | head -2
done
} | tee gpurun_out/s19_ab.txt
echo session done
