#!/bin/bash
# Round 2, session 10: same-box A/B of the parked GEMM / norm switches and of the wave-aware N-tile width.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python scripts/perf_gemm.py > gpurun_out/pg_$name.log 2>&1; echo "perf_gemm $name exit $?"
}
env WIW_GELU=3 WIW_GEMM_ACC3=1 WIW_GEMM_L2PF=1 WIW_SILU=2 timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/t_switches.log 2>&1
echo "tests with all switches on exit $?"; tail -3 gpurun_out/t_switches.log
run base WIW_WAVE_BN=0
run wavebn WIW_WAVE_BN=1
run bn160 WIW_FORCE_BN=160
run bn128 WIW_FORCE_BN=128
run gelu3 WIW_WAVE_BN=0 WIW_GELU=3
run acc3 WIW_WAVE_BN=0 WIW_GEMM_ACC3=1
run l2pf1 WIW_WAVE_BN=0 WIW_GEMM_L2PF=1
run l2pf2 WIW_WAVE_BN=0 WIW_GEMM_L2PF=2
python - <<'PY'
import re,glob,collections
rows=collections.OrderedDict()
names=["base","wavebn","bn160","bn128","gelu3","acc3","l2pf1","l2pf2"]
for n in names:
    for line in open(f"gpurun_out/pg_{n}.log"):
        m=re.search(r"'name': '([^']+)', 'ms': ([0-9.]+)", line)
        if m: rows.setdefault(re.sub(r" bn=\d+","",m.group(1)),{})[n]=float(m.group(2))
print("shape | "+" | ".join(names))
for k,v in rows.items(): print(k+" | "+" | ".join(f"{v.get(n,float('nan')):.4f}" for n in names))
PY
for v in 1 2; do WIW_SILU=$v timeout 200 python scripts/perf_attn.py norm > gpurun_out/pn_silu$v.log 2>&1; echo "perf norm SILU=$v"; grep groupnorm gpurun_out/pn_silu$v.log | head -6; done
echo session done
