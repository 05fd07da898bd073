// GroupNorm (+SiLU) and LayerNorm on fp16 channels-last activations.  HBM-bound kernels:
// 16-byte vector loads/stores, fp32 statistics combined in fp64.
//
// Reference call sites: nn.GroupNorm(32) in ResnetBlock2D / TemporalResnetBlock
// (DP/models/resnet.py:267,283,569,583 -- the temporal one normalises over (C/32, F, h, w) of one
// batch element, which in the [b][f][s][c] layout is the same kernel with rows_per_sample = F*S),
// TransformerSpatioTemporalModel.norm (transformer_temporal.py:237), conv_norm_out (unet :252);
// nn.LayerNorm in BasicTransformerBlock / TemporalBasicTransformerBlock (attention.py:371-432,659-694).
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

constexpr int kGnMaxC = 4096;

// ---------------------------------------------------------------- GroupNorm pass 1: partial sums
// grid (chunks, samples); block = vpr * RL threads, thread (v, rl) owns channels [8v, 8v+8) and rows rl, rl+RL, ...
__global__ void gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial, long long rows_per_sample,
                                int C, int G, int rows_per_chunk, int vpr, int RL) {
    extern __shared__ float sm[];            // [2*C]
    float* s_sum = sm;
    float* s_sq = sm + C;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float s[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
    const long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows_per_sample) r1 = rows_per_sample;
    const __half* base = x + ((long long)n * rows_per_sample) * C + v * 8;
    for (long long r = r0 + rl; r < r1; r += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + r * C));
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            s[2 * e] += f.x; q[2 * e] += f.x * f.x;
            s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_sum[v * 8 + e], s[e]);
        atomicAdd(&s_sq[v * 8 + e], q[e]);
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

// ---------------------------------------------------------------- GroupNorm pass 2: normalise (+SiLU)
__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                long long rows_per_sample, int C, int G, int stat_chunks, int rows_per_chunk,
                                int vpr, int RL, float eps, int silu) {
    extern __shared__ float sm[];            // scale[C], shift[C], mean[G], rstd[G]
    float* s_scale = sm;
    float* s_shift = sm + C;
    float* s_mean = sm + 2 * C;
    float* s_rstd = s_mean + G;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int cg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < stat_chunks; ++k) {
            const float* p = partial + (((long long)n * stat_chunks + k) * G + g) * 2;
            a += (double)p[0]; b += (double)p[1];
        }
        const double cnt = (double)rows_per_sample * cg;
        const double mean = a / cnt;
        double var = b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = (float)mean;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cg;
        const float sc = s_rstd[g] * gamma[c];
        s_scale[c] = sc;
        s_shift[c] = beta[c] - s_mean[g] * sc;
    }
    __syncthreads();
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = s_scale[v * 8 + e]; sh[e] = s_shift[v * 8 + e]; }
    const long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows_per_sample) r1 = rows_per_sample;
    const long long off = ((long long)n * rows_per_sample) * C + v * 8;
    for (long long r = r0 + rl; r < r1; r += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off + r * C));
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            float a = f.x * sc[2 * e] + sh[2 * e];
            float b = f.y * sc[2 * e + 1] + sh[2 * e + 1];
            if (silu) { a = silu_f(a); b = silu_f(b); }
            oh[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(y + off + r * C) = o;
    }
}

// ---------------------------------------------------------------- LayerNorm: one warp per row
// y = LN(x + tab[(row / tab_div) % tab_mod]) * gamma + beta; optionally also writes nothing else.
template <int VPL>     // 16-byte vectors per lane (C <= 256*VPL)
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const __half* __restrict__ tab, int tab_div, int tab_mod,
                                 long long M, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 3;
    float v[VPL][8];
    const __half* trow = tab ? tab + (size_t)((row / tab_div) % tab_mod) * C : nullptr;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * C + vi * 8));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
            }
            if (trow) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    // the reference adds in fp16 (hidden_states + emb, transformer_temporal.py:360)
                    v[i][e] = __half2float(__float2half_rn(v[i][e] + __half2float(trow[vi * 8 + e])));
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[i][e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
        }
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

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_groupnorm_workspace_floats(int samples, int groups) { return samples * 64 * groups * 2; }

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
    int chunks = (2 * num_sms() + samples - 1) / samples;
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    long long rpc = (rows_per_sample + chunks - 1) / chunks;
    if (rpc < RL) rpc = RL;
    chunks = (int)((rows_per_sample + rpc - 1) / rpc);
    dim3 grid(chunks, samples);
    gn_stats_kernel<<<grid, threads, 2 * C * sizeof(float), stream>>>(
        static_cast<const __half*>(x), workspace, rows_per_sample, C, G, (int)rpc, vpr, RL);
    WIW_CHECK_CUDA(cudaGetLastError());
    // apply: more, smaller chunks for parallelism
    int achunks = (8 * num_sms() + samples - 1) / samples;
    long long arpc = (rows_per_sample + achunks - 1) / achunks;
    if (arpc < RL) arpc = RL;
    achunks = (int)((rows_per_sample + arpc - 1) / arpc);
    dim3 agrid(achunks, samples);
    gn_apply_kernel<<<agrid, threads, (2 * C + 2 * G) * sizeof(float), stream>>>(
        static_cast<const __half*>(x), static_cast<__half*>(y), workspace, gamma, beta, rows_per_sample, C, G, chunks,
        (int)arpc, vpr, RL, eps, silu);
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
    const int rows_per_block = 8;
    const unsigned grid = (unsigned)((M + rows_per_block - 1) / rows_per_block);
    const __half* xi = static_cast<const __half*>(x);
    __half* yo = static_cast<__half*>(y);
    const int nvec = C / 8;
    if (nvec <= 32)
        layernorm_kernel<1><<<grid, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 64)
        layernorm_kernel<2><<<grid, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 96)
        layernorm_kernel<3><<<grid, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else if (nvec <= 160)
        layernorm_kernel<5><<<grid, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    else
        layernorm_kernel<8><<<grid, 256, 0, stream>>>(xi, yo, gamma, beta, tab, tab_div, tab_mod, M, C, eps);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
