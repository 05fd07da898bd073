// Throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, per SM, at 1..4 warps per scheduler.
// Question it answers: does the packed form halve the ISSUE slots of fp32 epilogue math (the GEGLU epilogue is issue-bound)?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int PACKED>
__global__ void bench(int reps, float seed, long long* out, float* sink) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    const float m = 0.999f + seed * 1e-6f, c = 1e-3f * seed;
    __syncthreads();
    const long long t0 = clock64();
    if (PACKED) {
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
        const u64 mm = pack(m, m), cc = pack(c, c);
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(mm), "l"(cc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack(p[i], a[2 * i], a[2 * i + 1]);
    } else {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const int reps = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int packed = 0; packed < 2; ++packed) {
            if (packed) bench<1><<<148, warps * 32>>>(reps, 1.0f, out, sink); else bench<0><<<148, warps * 32>>>(reps, 1.0f, out, sink);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            const double fma = double(reps) * 16 * warps * 32;
            printf("%s %2d warps/SM: %6.1f fp32 FMA/clk/SM (%lld cycles)\n", packed ? "FFMA2" : "FFMA ", warps, fma / double(h), h);
        }
    }
    return 0;
}
