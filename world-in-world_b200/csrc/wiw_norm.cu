// GroupNorm (+SiLU) and LayerNorm on fp16 channels-last activations.  HBM-bound kernels:
// 16-byte vector loads/stores with 4 independent rows in flight per thread, fp32 partial sums
// combined in fp64.
//
// Reference call sites: nn.GroupNorm(32) in ResnetBlock2D / TemporalResnetBlock
// (DP/models/resnet.py:267,283,569,583 -- the temporal one normalises over (C/32, F, h, w) of one
// batch element, which in the [b][f][s][c] layout is the same kernel with rows_per_sample = F*S),
// TransformerSpatioTemporalModel.norm (transformer_temporal.py:237), conv_norm_out (unet :252);
// nn.LayerNorm in BasicTransformerBlock / TemporalBasicTransformerBlock (attention.py:371-432,659-694).
#include "wiw_common.cuh"
#include "../../include/wiw.h"

#include <cstdlib>

namespace wiw {

constexpr int kGnMaxC = 4096;
constexpr int kGnMaxChunks = 1024;
constexpr int kUnroll = 4;

__device__ __forceinline__ void acc8(const uint4& u, float* s, float* q) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        s[2 * e] += f.x; q[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
    }
}

// ---------------------------------------------------------------- GroupNorm pass 1: partial sums
// grid (chunks, samples); block = vpr * RL threads, thread (v, rl) owns channels [8v, 8v+8) and rows rl, rl+RL, ...
__global__ void gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial, long long rows_per_sample,
                                int C, int G, int rows_per_chunk, int vpr, int RL) {
    // smem: per-row-lane partials [RL][C] for sum and sumsq, then [C] column totals; fixed summation order
    // (no atomics) so the statistics -- and the whole rollout -- are bit-reproducible run to run
    extern __shared__ float sm[];
    float* p_sum = sm;                       // [RL][C]
    float* p_sq = sm + (size_t)RL * C;       // [RL][C]
    float* s_sum = p_sq + (size_t)RL * C;    // [C]
    float* s_sq = s_sum + C;                 // [C]
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    float s[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
    const long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows_per_sample) r1 = rows_per_sample;
    const __half* base = x + ((long long)n * rows_per_sample) * C + v * 8;
    long long r = r0 + rl;
    const long long step = RL;
    for (; r + (kUnroll - 1) * step < r1; r += kUnroll * step) {
        uint4 u[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (r + k * step) * C));
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) acc8(u[k], s, q);
    }
    for (; r < r1; r += step) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + r * C));
        acc8(u, s, q);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        p_sum[rl * C + v * 8 + e] = s[e];
        p_sq[rl * C + v * 8 + e] = q[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < RL; ++k) { a += p_sum[k * C + c]; b += p_sq[k * C + c]; }
        s_sum[c] = a; s_sq[c] = b;
    }
    __syncthreads();
    const int cg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = 0; c < cg; ++c) { a += s_sum[g * cg + c]; b += s_sq[g * cg + c]; }
        float* p = partial + (((long long)n * gridDim.x + chunk) * G + g) * 2;
        p[0] = a; p[1] = b;
    }
}

// ---------------------------------------------------------------- GroupNorm pass 1b: combine partials -> mean, rstd
// one warp per (sample, group); fp64 accumulation
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int samples, int G,
                                   int chunks, double count, float eps) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= samples * G) return;
    const int n = wid / G, g = wid % G;
    double a = 0.0, b = 0.0;
    for (int k = lane; k < chunks; k += 32) {
        const float* p = partial + (((long long)n * chunks + k) * G + g) * 2;
        a += (double)p[0]; b += (double)p[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        const double mean = a / count;
        double var = b / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[wid * 2] = (float)mean;
        stats[wid * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// ---------------------------------------------------------------- GroupNorm pass 2: normalise (+SiLU)
__device__ __forceinline__ uint4 gn_apply8(const uint4& u, const float* sc, const float* sh, int silu) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        float a = fmaf(f.x, sc[2 * e], sh[2 * e]);
        float b = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
        if (silu) { a = silu_f(a); b = silu_f(b); }
        oh[e] = __floats2half2_rn(a, b);
    }
    return o;
}

__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                long long rows_per_sample, int C, int G, int rows_per_chunk,
                                int vpr, int RL, int silu) {
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int cg = C / G;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = v * 8 + e;
        const int g = c / cg;
        const float mean = stats[(n * G + g) * 2], rstd = stats[(n * G + g) * 2 + 1];
        sc[e] = rstd * gamma[c];
        sh[e] = beta[c] - mean * sc[e];
    }
    const long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows_per_sample) r1 = rows_per_sample;
    const long long off = ((long long)n * rows_per_sample) * C + v * 8;
    long long r = r0 + rl;
    const long long step = RL;
    for (; r + (kUnroll - 1) * step < r1; r += kUnroll * step) {
        uint4 u[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(x + off + (r + k * step) * C));
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) *reinterpret_cast<uint4*>(y + off + (r + k * step) * C) = gn_apply8(u[k], sc, sh, silu);
    }
    for (; r < r1; r += step) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off + r * C));
        *reinterpret_cast<uint4*>(y + off + r * C) = gn_apply8(u, sc, sh, silu);
    }
}

// ---------------------------------------------------------------- LayerNorm: one warp per R rows (all loads issued first)
// y = LN(fp16(x + tab[(row / tab_div) % tab_mod])) * gamma + beta
template <int VPL, int R>     // VPL 16-byte vectors per lane (C <= 256*VPL), R rows per warp
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const __half* __restrict__ tab, int tab_div, int tab_mod,
                                 long long M, int C, float eps) {
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
    if (row0 >= M) return;
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 3;
    uint4 raw[R][VPL];
#pragma unroll
    for (int k = 0; k < R; ++k) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            raw[k][i] = (vi < nvec && row0 + k < M) ? __ldg(reinterpret_cast<const uint4*>(x + (row0 + k) * C + vi * 8)) : make_uint4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const long long row = row0 + k;
        if (row >= M) break;
        float v[VPL][8];
        const __half* trow = tab ? tab + (size_t)((row / tab_div) % tab_mod) * C : nullptr;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            const __half2* h = reinterpret_cast<const __half2*>(&raw[k][i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
            }
            if (trow && vi < nvec) {
                const uint4 tu = __ldg(reinterpret_cast<const uint4*>(trow + vi * 8));
                const __half2* th = reinterpret_cast<const __half2*>(&tu);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // the reference adds in fp16 (hidden_states + emb, transformer_temporal.py:360)
                    const float2 tf = __half22float2(th[e]);
                    v[i][2 * e] = __half2float(__float2half_rn(v[i][2 * e] + tf.x));
                    v[i][2 * e + 1] = __half2float(__float2half_rn(v[i][2 * e + 1] + tf.y));
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[i][e];       // lanes beyond nvec hold zeros
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)C;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nvec) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; sq += d * d; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nvec) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    oh[e] = __floats2half2_rn((v[i][2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e],
                                              (v[i][2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1]);
                *reinterpret_cast<uint4*>(y + row * C + vi * 8) = o;
            }
        }
    }
}


// ---------------------------------------------------------------- LayerNorm, lane-group variant (default)
// ncu of layernorm_kernel<2,4> at C = 320 (profiles/r01_ncu_norm_v7.md): 255 warp-instructions per row, 64 % issue
// utilisation, 9x L1 load amplification (gamma / beta re-read per row, second vector 25 % full) at 37 % of DRAM peak.
// Here LPR lanes share a row so every lane owns VPL full vectors (C = 320: 8 lanes x 5 vectors, 4 rows per warp),
// gamma / beta sit in registers as packed halves for the warp's whole life, statistics are single-pass fp32 sums
// (sum, sum of squares) reduced over the LPR lanes, and the next row group's loads are issued before the current
// group's arithmetic.  ~80 warp-instructions per row.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256, 2)
layernorm_lpr_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const __half* __restrict__ tab, int tab_div, int tab_mod,
                     long long M, int C, float eps, long long groups, long long total_warps) {
    constexpr int RPW = 32 / LPR;                       // rows per warp per group
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, j = lane % LPR;
    const int nvec = C >> 3;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp_global >= total_warps) return;
    // contiguous, balanced share of the row groups: [w * G / W, (w + 1) * G / W)
    const long long g0 = warp_global * groups / total_warps;
    const int groups_per_warp = (int)((warp_global + 1) * groups / total_warps - g0);
    if (groups_per_warp <= 0) return;
    uint32_t gpk[VPL][4], bpk[VPL][4];                  // gamma / beta of this lane's channels as half2
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = j + LPR * i;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 gv = (vi < nvec) ? *reinterpret_cast<const float2*>(gamma + vi * 8 + 2 * e) : make_float2(0.f, 0.f);
            const float2 bv = (vi < nvec) ? *reinterpret_cast<const float2*>(beta + vi * 8 + 2 * e) : make_float2(0.f, 0.f);
            const __half2 gh = __floats2half2_rn(gv.x, gv.y), bh = __floats2half2_rn(bv.x, bv.y);
            gpk[i][e] = *reinterpret_cast<const uint32_t*>(&gh);
            bpk[i][e] = *reinterpret_cast<const uint32_t*>(&bh);
        }
    }
    const float inv_c = 1.0f / (float)C;
    auto load_group = [&](long long g, uint4* raw) {
        const long long row = g * RPW + sub;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int vi = j + LPR * i;
            raw[i] = (vi < nvec && row < M) ? __ldg(reinterpret_cast<const uint4*>(x + row * C + vi * 8)) : make_uint4(0, 0, 0, 0);
        }
    };
    uint4 cur[VPL], nxt[VPL];
    load_group(g0, cur);
    for (int gi = 0; gi < groups_per_warp; ++gi) {
        const long long g = g0 + gi;
        if (g * RPW >= M) break;                        // warp-uniform
        if (gi + 1 < groups_per_warp) load_group(g + 1, nxt);
        const long long row = g * RPW + sub;
        if (tab != nullptr && row < M) {                // the reference adds the embedding in fp16 (transformer_temporal.py:360)
            const __half* trow = tab + (size_t)((row / tab_div) % tab_mod) * C;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int vi = j + LPR * i;
                if (vi < nvec) {
                    const uint4 tu = __ldg(reinterpret_cast<const uint4*>(trow + vi * 8));
                    __half2* h = reinterpret_cast<__half2*>(&cur[i]);
                    const __half2* th = reinterpret_cast<const __half2*>(&tu);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = __hadd2(h[e], th[e]);
                }
            }
        }
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const __half2* h = reinterpret_cast<const __half2*>(&cur[i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                sum += f.x; sq = fmaf(f.x, f.x, sq);
                sum += f.y; sq = fmaf(f.y, f.y, sq);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        const float mean = sum * inv_c;
        float var = fmaf(-mean, mean, sq * inv_c);
        var = var < 0.f ? 0.f : var;
        const float rstd = rsqrtf(var + eps);
        const float nmr = -mean * rstd;
        if (row < M) {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int vi = j + LPR * i;
                if (vi < nvec) {
                    const __half2* h = reinterpret_cast<const __half2*>(&cur[i]);
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // normalise in fp32, affine as one packed fp16 FMA (gamma / beta stay packed: no loop-invariant
                        // unpacking for the compiler to hoist into 80 more registers)
                        const float2 f = __half22float2(h[e]);
                        const __half2 t = __floats2half2_rn(fmaf(f.x, rstd, nmr), fmaf(f.y, rstd, nmr));
                        oh[e] = __hfma2(t, *reinterpret_cast<const __half2*>(&gpk[i][e]), *reinterpret_cast<const __half2*>(&bpk[i][e]));
                    }
                    *reinterpret_cast<uint4*>(y + row * C + vi * 8) = o;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
    }
}

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_groupnorm_workspace_floats(int samples, int groups) { return samples * (kGnMaxChunks + 1) * groups * 2; }

extern "C" int wiw_groupnorm(const void* x, void* y, int samples, long long rows_per_sample, int C, int G,
                             const float* gamma, const float* beta, float eps, int silu, float* workspace,
                             wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(x && y && gamma && beta && workspace, "null pointer");
    WIW_REQUIRE(C % 8 == 0 && C % G == 0 && C <= kGnMaxC && G <= 64, "unsupported channel/group count");
    WIW_REQUIRE(samples > 0 && rows_per_sample > 0, "empty input");
    const int vpr = C / 8;
    int RL = 256 / vpr;
    if (RL < 1) RL = 1;
    const int threads = vpr * RL;
    WIW_REQUIRE(threads <= 1024, "C too large");
    float* partial = workspace;
    float* stats = workspace + (size_t)samples * kGnMaxChunks * G * 2;
    const double count = (double)rows_per_sample * (C / G);
    // Whole waves of resident CTAs: 43 chunks x 28 frames = 1204 CTAs on 592 resident slots is 2.03 waves -- the 0.03 costs a
    // full CTA time.  Round the chunk count DOWN so samples * chunks <= waves * occupancy * SMs.
    static const int wavefit = getenv("WIW_WAVEFIT") ? atoi(getenv("WIW_WAVEFIT")) : 1;       // development aid for A/B runs
    int occ_s = 4, occ_a = 4;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, gn_stats_kernel, threads, (size_t)(2 * RL + 2) * C * sizeof(float));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, gn_apply_kernel, threads, 0);
    if (occ_s < 1) occ_s = 1;
    if (occ_a < 1) occ_a = 1;
    auto pick_chunks = [&](int target_ctas) {
        long long chunks = wavefit ? target_ctas / samples : (target_ctas + samples - 1) / samples;
        if (chunks > kGnMaxChunks) chunks = kGnMaxChunks;
        if (chunks < 1) chunks = 1;
        long long rpc = (rows_per_sample + chunks - 1) / chunks;
        const long long min_rows = (long long)RL * kUnroll;
        if (rpc < min_rows) rpc = min_rows;
        return rpc;
    };
    const long long rpc = pick_chunks(wavefit ? 2 * occ_s * num_sms() : 8 * num_sms());
    const int chunks = (int)((rows_per_sample + rpc - 1) / rpc);
    gn_stats_kernel<<<dim3(chunks, samples), threads, (size_t)(2 * RL + 2) * C * sizeof(float), stream>>>(
        static_cast<const __half*>(x), partial, rows_per_sample, C, G, (int)rpc, vpr, RL);
    WIW_CHECK_CUDA(cudaGetLastError());
    const int warps = samples * G;
    gn_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, stream>>>(partial, stats, samples, G, chunks, count, eps);
    WIW_CHECK_CUDA(cudaGetLastError());
    const long long arpc = pick_chunks(wavefit ? 4 * occ_a * num_sms() : 16 * num_sms());
    const int achunks = (int)((rows_per_sample + arpc - 1) / arpc);
    gn_apply_kernel<<<dim3(achunks, samples), threads, 0, stream>>>(
        static_cast<const __half*>(x), static_cast<__half*>(y), stats, gamma, beta, rows_per_sample, C, G, (int)arpc, vpr, RL,
        silu);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_layernorm(const void* x, void* y, long long M, int C, const float* gamma, const float* beta,
                             float eps, const void* tab_, int tab_div, int tab_mod, wiw_stream_t stream_) {
    const __half* tab = static_cast<const __half*>(tab_);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(x && y && gamma && beta, "null pointer");
    WIW_REQUIRE(C % 8 == 0 && C <= 2048, "LayerNorm supports C % 8 == 0, C <= 2048");
    WIW_REQUIRE(!tab || (tab_div > 0 && tab_mod > 0), "tab needs div/mod");
    if (M == 0) return 0;
    const __half* xi = static_cast<const __half*>(x);
    __half* yo = static_cast<__half*>(y);
    const int nvec = C / 8;
    static const int norm_v = getenv("WIW_NORM_V") ? atoi(getenv("WIW_NORM_V")) : 3;          // 3 = lane-group kernel (default), 1 = one warp per row (development aid for A/B runs)
    static const int wavefit = getenv("WIW_WAVEFIT") ? atoi(getenv("WIW_WAVEFIT")) : 1;       // development aid for A/B runs
    if (norm_v == 3) {
        auto launch = [&](auto kern, int lpr) -> int {
            const int rpw = 32 / lpr;
            const long long groups = (M + rpw - 1) / rpw;
            // whole waves of resident CTAs (no ragged last wave): W = waves * occupancy * SMs CTAs of 8 warps, the row groups
            // split evenly over the warps (~8-16 groups each so gamma / beta and the prefetch pipeline amortise)
            int occ = 2;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
            if (occ < 1) occ = 1;
            const long long slots = (long long)occ * num_sms();
            long long waves = (groups + slots * 8 * 12 - 1) / (slots * 8 * 12);
            if (waves < 1) waves = 1;
            long long blocks = wavefit ? waves * slots : (groups + 8 * 6 - 1) / (8 * 6);
            if (blocks * 8 > groups) blocks = (groups + 7) / 8;
            const long long warps = blocks * 8 < groups ? blocks * 8 : groups;
            kern<<<(unsigned)blocks, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps, groups, warps);
            WIW_CHECK_CUDA(cudaGetLastError());
            return 0;
        };
        if (nvec <= 8) return launch(layernorm_lpr_kernel<8, 1>, 8);
        if (nvec <= 16) return launch(layernorm_lpr_kernel<8, 2>, 8);
        if (nvec <= 24) return launch(layernorm_lpr_kernel<8, 3>, 8);
        if (nvec <= 32) return launch(layernorm_lpr_kernel<8, 4>, 8);
        if (nvec <= 40) return launch(layernorm_lpr_kernel<8, 5>, 8);
        if (nvec <= 48) return launch(layernorm_lpr_kernel<16, 3>, 16);
        if (nvec <= 64) return launch(layernorm_lpr_kernel<16, 4>, 16);
        if (nvec <= 80) return launch(layernorm_lpr_kernel<16, 5>, 16);
        if (nvec <= 96) return launch(layernorm_lpr_kernel<32, 3>, 32);
        if (nvec <= 128) return launch(layernorm_lpr_kernel<32, 4>, 32);
        // C > 1024: fall through to the one-warp-per-row kernel (A/B on B200: C = 320 +52 %, 640 +11 %, 960 +34 %, 1280 -10 %)
    }
    const int wpb = 8;                                    // warps per block
    auto grid = [&](int R) { return (unsigned)((M + (long long)wpb * R - 1) / ((long long)wpb * R)); };
    if (nvec <= 32)
        layernorm_kernel<1, 4><<<grid(4), 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 64)
        layernorm_kernel<2, 4><<<grid(4), 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 96)
        layernorm_kernel<3, 2><<<grid(2), 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 160)
        layernorm_kernel<5, 2><<<grid(2), 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else
        layernorm_kernel<8, 1><<<grid(1), 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
