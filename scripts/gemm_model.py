"""Per-shape floor model for the implicit-GEMM launches of one UNet forward (reads bench.py --dump-gemm-profile output).

Floors per launch (profiles/r01_mma_cost.txt):
  MMA   = (#tcgen05.mma instructions per CTA pair) x 128 cycles / SM clock: the MMA cost is per instruction (122-128 cycles for a
          cta_group::2 instruction at ANY M, N), so a launch issues ceil(M/256) x ceil(N/bn) x K/16 pair-instructions over 74 CTA pairs
  HBM   = algorithmic bytes (A once + weights once + output [+ residual / aux]) / measured copy bandwidth
  floor = max(MMA, HBM)
Usage: python scripts/gemm_model.py profiles/r01_gemm_profile_v9_per_shape.json [sm_clock_ghz] > profiles/r01_gemm_model_v9.md"""
import json
import math
import re
import sys

path = sys.argv[1]
clk = float(sys.argv[2]) if len(sys.argv) > 2 else 1.515          # median SM clock under load in the bench line (sw_power_cap)
HBM = 6.3e12                                                      # a plain copy kernel's DRAM rate (profiles/r01_ncu_full_v9.md)
rows = json.load(open(path))
print("# Implicit-GEMM launches of one UNet forward: measured vs floor (per shape, live CUDA events inside the forward)\n")
print(f"SM clock {clk} GHz (power-capped), MMA floor = pair-instructions x 128 cycles, HBM floor at {HBM/1e12:.1f} TB/s.  Source: `{path}`.\n")
print("| shape | launches | ms total | TFLOP/s | MMA floor us | HBM floor us | measured us | measured / floor | bound |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
tot_ms = tot_floor = 0.0
for r in rows:
    s = r["shape"]
    m, k, n = (int(re.search(p + r"=(\d+)", s).group(1)) for p in ("M", "K", "N"))
    bn = int(re.search(r"bn=(\d+)", s).group(1))
    geglu, res, aux = " geglu" in s, " res" in s, " aux" in s
    instr = math.ceil(m / 256) * math.ceil(n / bn) * (k // 16)
    mma_us = math.ceil(instr / 74) * 128 / (clk * 1e3)
    n_out = n // 2 if geglu else n
    taps = int(re.search(r"taps=(\d+)", s).group(1))
    a_bytes = m * (k // taps) * 2                                  # the activation tensor is read from HBM once (taps hit L2)
    byts = a_bytes + n * k * 2 + m * n_out * 2 * (1 + res + aux)
    hbm_us = byts / HBM * 1e6
    us = 1e3 * r["ms"] / r["launches"]
    floor = max(mma_us, hbm_us)
    tot_ms += r["ms"]; tot_floor += floor * r["launches"] / 1e3
    if r["ms"] >= 0.4:
        print(f"| {s} | {r['launches']} | {r['ms']:.2f} | {r['tflops']:.0f} | {mma_us:.0f} | {hbm_us:.0f} | {us:.0f} | {us / floor:.2f} | {'MMA' if mma_us >= hbm_us else 'HBM'} |")
print(f"\nAll {sum(r['launches'] for r in rows)} launches: measured {tot_ms:.1f} ms, sum of floors {tot_floor:.1f} ms -> {tot_ms / tot_floor:.2f}x the floor "
      f"(shapes under 0.4 ms total are not listed).")
