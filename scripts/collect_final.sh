#!/bin/bash
# Copy the artefacts of scripts/gpu_r2_final.sh from gpurun_out/ into profiles/ (tracked) and rebuild the summaries.
set -eu
cd "$(dirname "$0")/.."
cp gpurun_out/launches_unet.csv profiles/r02_launches_unet_final.csv
tail -1 gpurun_out/bench.log > profiles/r02_bench_1gpu_final.json
tail -1 gpurun_out/bench_ref.log > profiles/r02_bench_reference_arm_final.json
cp gpurun_out/parity.jsonl profiles/r02_parity_margins_final.jsonl
tail -1 gpurun_out/smoke.log > profiles/r02_smoke_final.log
tail -6 gpurun_out/tests_gpu.log > profiles/r02_tests_gpu_final.log
cp gpurun_out/gemm_profile.json profiles/r02_gemm_profile_unet_per_shape.json
cp gpurun_out/gemm_profile_vae.json profiles/r02_gemm_profile_vae_per_shape.json
cp gpurun_out/closed_loop.json profiles/r02_closed_loop_final.json
python scripts/ncu_traffic.py gpurun_out/full_kernels.ncu-rep profiles/r02_ncu_traffic.json profiles/r02_ncu_full.md > /dev/null
python - <<'PY'
import csv, re, json
rows = list(csv.reader(open('profiles/r02_launches_unet_final.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
sec = data[len(data) // 2:]
agg = {}
for r in sec:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('wiw::', '')
    v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    ms = v * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1, 'msecond': 1, 'nsecond': 1e-6}.get(u, 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
lines = [f"One UNet forward (second of two) under ncu --metrics gpu__time_duration.sum (cold-cache, serialised: compare SHARES): total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches"]
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:<42} {c:>4} {ms:>8.2f} ms {100 * ms / tot:>5.1f}%")
open('profiles/r02_launches_unet_final_summary.txt', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))
b = json.loads(open('profiles/r02_bench_1gpu_final.json').read())
print({k: b[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', round(b['e2e']['value'], 3), 'frac', b['roofline']['frac'], 'traffic', b['roofline']['traffic'], b['stages'], b['clocks'])
print('closed loop req/s', json.load(open('profiles/r02_closed_loop_final.json'))['requests_per_s'])
PY
