// Micro-benchmark: cost of one tcgen05.mma (kind::f16, K=16) on B200 as a function of M, N, of where the A operand
// lives (shared memory "SS" vs tensor memory "TS"), of the accumulator dependency (1 chain vs 2 interleaved
// accumulators), of cta_group (1 vs 2) and of the operand addresses (one smem stage vs 4 rotating stages).
// One thread issues `reps` MMAs back to back on resident operands (contents irrelevant), commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I world-in-world_b200/csrc scripts/mma_bench.cu -o scripts/mma_bench.bin
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "wiw_common.cuh"
using namespace wiw;

struct Cfg { int m, n, ts, chains, stages, cg, acc, single; };
constexpr int kStage = 16384 + 32768;

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_bench(Cfg c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 4 * kStage, slot = bar + 16;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 4 * kStage / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        if (CG == 2) { tmem_alloc_2sm(slot, 512); tmem_relinquish_2sm(); } else { tmem_alloc(slot, 512); tmem_relinquish(); }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
    if (threadIdx.x < 32 && rank == 0) {
        // converged warp, one elected lane issues (c.single = 1: the old `single thread in a divergent branch` pattern)
        const uint32_t idesc = make_idesc_f16(c.n, 0, 0, c.m);
        const uint32_t d_stride = (c.chains == 2) ? 256 : 0;
        const uint32_t a_tmem = tmem + ((c.chains == 2) ? 128 + 0 : 256);
        uint32_t ph = 0;
        long long best = 1LL << 60, best_ns = 0;
        for (int round = 0; round < 3; ++round) {
            unsigned long long g0, g1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            const long long t0 = clock64();
            for (int it = 0; it < reps / 4; ++it) {
                const int st = it % c.stages;
                const uint64_t da = make_desc_sw128(base + st * kStage), db = make_desc_sw128(base + st * kStage + 16384);
                const uint32_t d = tmem + ((c.chains == 2 && (it & 1)) ? ((c.ts) ? 64u : d_stride) : 0u);
                const uint32_t accum0 = c.acc ? (it >= c.chains ? 1u : 0u) : 0u;
                if (c.single ? (threadIdx.x == 0) : elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t accum = (kk > 0 && c.acc) ? 1u : accum0;
                        if (CG == 2) umma_f16_ss_2sm(d, da + 2 * kk, db + 2 * kk, idesc, accum);
                        else if (c.ts) umma_f16_ts(d, a_tmem + 8 * kk, db + 2 * kk, idesc, accum);
                        else umma_f16_ss(d, da + 2 * kk, db + 2 * kk, idesc, accum);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) { if (CG == 2) umma_commit_2sm(bar, 1); else umma_commit(bar); }
            __syncwarp();
            mbar_wait(bar, ph);
            ph ^= 1;
            const long long t1 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            if (t1 - t0 < best) { best = t1 - t0; best_ns = (long long)(g1 - g0); }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = best; out[1] = best_ns; }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512);
    }
}

static int run(Cfg c, int grid) {
    static long long* d = nullptr;
    if (!d) cudaMalloc(&d, 16);
    const int smem = 1024 + 4 * kStage + 64;
    const int reps = 2048;
    if (c.cg == 1) {
        cudaFuncSetAttribute(mma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        mma_bench<1><<<grid, 128, smem>>>(c, reps, d);
    } else {
        cudaFuncSetAttribute(mma_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, mma_bench<2>, c, reps, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double cyc = double(h[0]) / reps;
    printf("grid=%3d cg=%d %s %s M=%3d N=%3d chains=%d stages=%d : %6.1f cycles/MMA (%.2f GHz)  N-proportional ideal %5.1f  %s\n", grid, c.cg,
           c.single ? "single-thread" : "elected      ", c.ts ? "TS" : "SS", c.m, c.n, c.chains, c.stages, cyc, double(h[0]) / double(h[1]), (c.m / c.cg) * c.n / 256.0, e ? cudaGetErrorString(e) : "");
    return e ? 1 : 0;
}

int main(int argc, char** argv) {
    const int grid = 148;
    if (argc > 1 && !strcmp(argv[1], "swap")) {
        // cost of the M extent (the swap-AB question: weights as the M operand, 256 pixels as N): M = 64 / 128 per instruction
        for (int n : {64, 128, 256}) {
            if (run({128, n, 0, 1, 4, 2, 1, 0}, grid)) return 1;      // cta_group::2, M = 128 (64 rows per CTA)
            if (run({256, n, 0, 1, 4, 2, 1, 0}, grid)) return 1;      // cta_group::2, M = 256 (reference point)
            if (run({64, n, 0, 1, 4, 1, 1, 0}, grid)) return 1;       // cta_group::1, M = 64
            if (run({128, n, 0, 1, 4, 1, 1, 0}, grid)) return 1;      // cta_group::1, M = 128 (reference point)
        }
        return 0;
    }
    for (int single : {1, 0}) {
        for (int n : {32, 64, 80, 128, 160, 192, 224, 256})
            if (run({128, n, 0, 1, 4, 1, 1, single}, grid)) return 1;
        for (int n : {64, 80, 128, 256})
            if (run({128, n, 1, 1, 4, 1, 1, single}, grid)) return 1;
        for (int n : {64, 128, 160, 224, 256})
            if (run({256, n, 0, 1, 4, 2, 1, single}, grid)) return 1;
    }
    // two interleaved accumulators
    for (int n : {64, 128, 256})
        if (run({128, n, 0, 2, 4, 1, 1, 0}, grid)) return 1;
    return 0;
}
