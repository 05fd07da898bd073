// Micro-benchmark v2: throughput of tcgen05.mma (kind::f16, K = 16) on B200, separating the ISSUE rate of the single
// issuing thread from the COMPLETION rate of the tensor pipe, as a function of
//   cta_group (1 | 2), M, N, where A lives (shared memory "SS" | tensor memory "TS"), and the number of independent
//   TMEM accumulators rotated PER INSTRUCTION (1, 2, 4).
// v1 (scripts/mma_bench.cu) measured ~203 / ~122 cycles for every shape: its issue loop (runtime `%`, runtime
// branches around the asm under elect_one) was the bottleneck, and it alternated accumulators only every 4th instruction.
// Here every parameter is a template constant, the issue loop is straight-line code (64 instructions unrolled), smem
// descriptors rotate over 4 stages x 4 k-steps, and two clocks are taken: after the last issue (issue-bound rate) and after
// the commit's mbarrier flips (completion rate).  Also measures tcgen05.ld / tcgen05.st throughput per warp and per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I world-in-world_b200/csrc scripts/mma_bench2.cu -o scripts/mma_bench2.bin
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "wiw_common.cuh"
using namespace wiw;

constexpr int kStageA = 16384, kStageB = 32768, kStage = kStageA + kStageB;     // A [128][64] fp16, B [256][64] fp16 (SW128 K-major)
constexpr int kStages = 4;
constexpr int kUnroll = 64;

__device__ __forceinline__ void umma_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

template <int CG, int M, int N, int TS, int NACC, int BMN = 0>
__global__ void __launch_bounds__(128, 1) mma_bench(int outer, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kStages * kStage, slot = bar + 16;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < kStages * kStage / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3C003C00u;
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
        constexpr uint32_t idesc = make_idesc_f16(N, 0, BMN, M);
        // accumulators: NACC x N columns from column 0; the TMEM A operand (64 columns = K 128) sits above them
        constexpr uint32_t a_col = (NACC * N <= 448) ? 448 : 0;
        static_assert(NACC * N <= 512, "accumulators do not fit");
        uint32_t ph = 0;
        long long best_done = 1LL << 60, best_issue = 0, best_ns = 0;
        for (int round = 0; round < 3; ++round) {
            unsigned long long g0, g1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            const long long t0 = clock64();
            if (elect_one()) {
                for (int o = 0; o < outer; ++o) {
#pragma unroll
                    for (int i = 0; i < kUnroll; ++i) {
                        const uint32_t st = (i >> 2) & (kStages - 1), kk = i & 3;
                        const uint64_t da = make_desc_sw128(base + st * kStage) + 2 * kk;
                        // BMN: B consumed MN-major from a [k rows][64 elements] SW128 tile (the attention's V operand): K groups of 8 rows at
                        // SBO = 1024 B, the second 64-wide N atom (N = 80: the tile of ones) at LBO = 16 KB; a K = 16 step advances 2048 B
                        const uint32_t b_addr = base + st * kStage + kStageA;
                        const uint64_t db = BMN ? ((uint64_t((b_addr & 0x3FFFF) >> 4)) | (uint64_t(16384 >> 4) << 16) | (uint64_t(1024 >> 4) << 32) |
                                                   (uint64_t(1) << 46) | (uint64_t(2) << 61)) + (uint64_t)kk * (2048 >> 4)
                                                : make_desc_sw128(b_addr) + 2 * kk;
                        const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
                        if (CG == 2) { if (TS) umma_ts_2sm(d, tmem + a_col + 8 * kk, db, idesc, 1u); else umma_f16_ss_2sm(d, da, db, idesc, 1u); }
                        else { if (TS) umma_f16_ts(d, tmem + a_col + 8 * kk, db, idesc, 1u); else umma_f16_ss(d, da, db, idesc, 1u); }
                    }
                }
            }
            __syncwarp();
            const long long t1 = clock64();
            if (elect_one()) { if (CG == 2) umma_commit_2sm(bar, 1); else umma_commit(bar); }
            __syncwarp();
            mbar_wait(bar, ph);
            ph ^= 1;
            const long long t2 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            if (t2 - t0 < best_done) { best_done = t2 - t0; best_issue = t1 - t0; best_ns = (long long)(g1 - g0); }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = best_done; out[1] = best_issue; out[2] = best_ns; }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512);
    }
}

static long long* g_out = nullptr;

template <int CG, int M, int N, int TS, int NACC, int BMN = 0>
static int run() {
    if (!g_out) cudaMalloc(&g_out, 32);
    const int smem = 1024 + kStages * kStage + 64, outer = 32, grid = 148;
    auto k = mma_bench<CG, M, N, TS, NACC, BMN>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (CG == 1) {
        k<<<grid, 128, smem>>>(outer, g_out);
    } else {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k, outer, g_out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3] = {0, 0, 0}; cudaMemcpy(h, g_out, 24, cudaMemcpyDeviceToHost);
    const double n = double(outer) * kUnroll;
    // per-SM operand bytes one instruction reads from shared memory: A (M/CG rows x 16 x 2 B, unless TS) + B (N/CG rows x 32 B)
    const double smem_bytes = (TS ? 0.0 : (M / CG) * 32.0) + (N / CG) * 32.0;
    printf("cg=%d %s%s M=%3d N=%3d acc=%d : done %6.1f  issue %6.1f cycles/MMA (%.2f GHz) | tensor floor M*N/(256*cg) = %5.1f | smem operand bytes/128 = %5.1f  %s\n",
           CG, TS ? "TS" : "SS", BMN ? " B=MN-major" : "", M, N, NACC, double(h[0]) / n, double(h[1]) / n, double(h[0]) / double(h[2]),
           double(M) * N / (256.0 * CG), smem_bytes / 128.0, e ? cudaGetErrorString(e) : "");
    return e ? 1 : 0;
}

// ---------------------------------------------------------------- tcgen05.ld / tcgen05.st throughput
template <int STORE>
__global__ void __launch_bounds__(512, 1) tmem_rw_bench(int warps_active, int reps, long long* out) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps_active) {
        const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
        for (int it = 0; it < reps; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (STORE) tmem_st_32x32b_x16(t + c * 32, r), tmem_st_32x32b_x16(t + c * 32 + 16, r + 16);
                else tmem_ld_32x32b_x32(t + c * 32, r);
            }
            if (STORE) tmem_st_wait(); else tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (acc == 0x12345678u) out[3] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------- tcgen05.ld LATENCY: one x32 load, wait, consume -- the epilogue's pattern
__global__ void __launch_bounds__(512, 1) tmem_lat_bench(int warps_active, int reps, int loads, long long* out) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps_active) {
        const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
        for (int it = 0; it < reps; ++it) {
            uint32_t r[32], r2[32];
            tmem_ld_32x32b_x32(t + (acc & 1u), r);                 // address depends on the previous result: no overlap between iterations
            if (loads == 2) tmem_ld_32x32b_x32(t + 32 + (acc & 1u), r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
            if (loads == 2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += r2[i];
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (acc == 0x12345678u) out[3] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
static void run_tmem_lat(int warps, int loads) {
    const int reps = 256;
    tmem_lat_bench<<<148, 512>>>(warps, reps, loads, g_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, g_out, 8, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld latency, %d x32 load(s) + wait + consume, %2d warps/SM: %6.1f cycles per iteration  %s\n", loads, warps,
           double(h) / reps, e ? cudaGetErrorString(e) : "");
}

static void run_tmem(int store, int warps) {
    const int reps = 256;
    if (store) tmem_rw_bench<1><<<148, 512>>>(warps, reps, g_out); else tmem_rw_bench<0><<<148, 512>>>(warps, reps, g_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, g_out, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(reps) * 4 * 32 * 32 * 4 * warps;            // per SM
    printf("tcgen05.%s 32x32b.x32, %2d warps/SM: %7.1f B/clk/SM (%5.1f cycles per 4 KB warp access)  %s\n", store ? "st" : "ld", warps,
           bytes / double(h), double(h) / (double(reps) * 4), e ? cudaGetErrorString(e) : "");
}

int main() {
    int bad = 0;
    // cta_group::1, A from shared memory: N sweep at M = 128, 1 and 4 accumulators
    bad |= run<1, 128, 32, 0, 4>();  bad |= run<1, 128, 64, 0, 1>();  bad |= run<1, 128, 64, 0, 4>();
    bad |= run<1, 128, 80, 0, 4>();  bad |= run<1, 128, 128, 0, 1>(); bad |= run<1, 128, 128, 0, 2>(); bad |= run<1, 128, 128, 0, 4>();
    bad |= run<1, 128, 160, 0, 2>(); bad |= run<1, 128, 256, 0, 1>(); bad |= run<1, 128, 256, 0, 2>();
    bad |= run<1, 64, 64, 0, 4>();   bad |= run<1, 64, 128, 0, 4>();  bad |= run<1, 64, 256, 0, 2>();
    // cta_group::1, A from tensor memory
    bad |= run<1, 128, 32, 1, 4>();  bad |= run<1, 128, 64, 1, 1>();  bad |= run<1, 128, 64, 1, 4>();
    bad |= run<1, 128, 80, 1, 4>();  bad |= run<1, 128, 128, 1, 1>(); bad |= run<1, 128, 128, 1, 3>();
    bad |= run<1, 128, 256, 1, 1>();
    // cta_group::2
    bad |= run<2, 256, 64, 0, 4>();  bad |= run<2, 256, 128, 0, 2>(); bad |= run<2, 256, 160, 0, 2>(); bad |= run<2, 256, 224, 0, 2>();
    bad |= run<2, 256, 256, 0, 1>(); bad |= run<2, 256, 256, 0, 2>(); bad |= run<2, 128, 128, 0, 2>(); bad |= run<2, 128, 256, 0, 2>();
    bad |= run<2, 256, 128, 1, 2>(); bad |= run<2, 256, 256, 1, 1>();
    // B consumed MN-major (the attention's P V product: V tile as loaded by TMA, N = 64 / 80)
    bad |= run<1, 128, 64, 0, 4, 1>(); bad |= run<1, 128, 80, 0, 4, 1>(); bad |= run<1, 128, 64, 1, 4, 1>(); bad |= run<1, 128, 80, 1, 4, 1>();
    bad |= run<1, 128, 128, 1, 2, 1>();
    for (int w : {1, 4, 8, 16}) run_tmem(0, w);
    for (int w : {1, 4, 8}) { run_tmem_lat(w, 1); run_tmem_lat(w, 2); }
    for (int w : {1, 4, 8, 16}) run_tmem(1, w);
    return bad;
}
