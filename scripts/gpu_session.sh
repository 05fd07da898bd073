#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench's timed region and
# `ncu --set full` captures of the hot kernels.  Every leg runs under its own timeout and logs to gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh [tests] [bench] [launches] [full] [perf]'
set -u
mkdir -p gpurun_out
LEGS="${*:-tests bench launches full perf}"
has() { [[ " $LEGS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1

if has tests; then
    timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
    echo "tests exit $?" | tee -a gpurun_out/tests_gpu.log
    tail -5 gpurun_out/tests_gpu.log
fi
if has bench; then
    timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
    echo "bench exit $?"; tail -1 gpurun_out/bench.log
    timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err
    echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
fi
if has perf; then
    timeout 300 python scripts/perf_gemm.py > gpurun_out/perf_gemm.log 2>&1; echo "perf_gemm exit $?"
    timeout 300 python scripts/perf_attn.py > gpurun_out/perf_attn.log 2>&1; echo "perf_attn exit $?"
    true
    true
fi
if has launches; then
    # launch list of the SAME command as the bench, timed region only.  ncu walks every kernel node of the replayed CUDA graphs
    # at ~5 launches/s: a whole clip (14.5 k launches) does not fit a sensible budget, so the list is capped at two Euler steps
    # (-c 1500); scripts/gpu_final.sh lists one eager UNet forward instead, which takes 30 s.
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv \
        --log-file gpurun_out/launches_bench.csv \
        python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/launches_bench.log 2>&1
    echo "launches exit $?"; wc -l gpurun_out/launches_bench.csv
fi
if has full; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|attn_spatial' \
        -o gpurun_out/full_kernels -f python scripts/ncu_kernels.py all > gpurun_out/full_kernels.log 2>&1
    echo "full exit $?"; ls -la gpurun_out/*.ncu-rep
fi
echo session done
