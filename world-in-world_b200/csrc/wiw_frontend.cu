// Conditioning front-end kernels (once per request): the CLIP ViT image encoder of
// StableVideoDiffusionPipeline._encode_image (DP/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:183-229)
// runs on the implicit-GEMM / LayerNorm kernels of this library plus the three small kernels below.
//   wiw_resize_separable_f32   the pipeline's _resize_with_antialiasing (Gaussian blur + bicubic, align_corners=True; :643-747)
//                              as ONE separable linear operator per axis (weights composed on the host, wiw_b200/clip.py), with
//                              the CLIP mean / std normalisation applied to the result: uint8 planes in, fp16 planes out
//   wiw_patchify               [B][3][224][224] -> im2col rows of the non-overlapping patch conv (transformers CLIPVisionEmbeddings)
//   wiw_attention_small        multi-head self-attention for short sequences and any even head_dim <= 128 (ViT-H/14: 257 tokens,
//                              16 heads of 80) -- 0.08 GFLOP per layer: a shared-memory fp32 kernel, not a tensor-core one
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

// ---- separable resampler: src uint8 [P][H][W] (values / 255) -> tmp fp32 [P][H][ow] -> dst fp16 [P][oh][ow]
__global__ void rs_h_kernel(const uint8_t* __restrict__ src, float* __restrict__ tmp, long long rows, int W, int ow,
                            const int* __restrict__ bounds, const float* __restrict__ wt, int ksize) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ow) return;
    const int xx = (int)(i % ow);
    const long long row = i / ow;
    const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
    const uint8_t* p = src + row * W + x0;
    const float* k = wt + (long long)xx * ksize;
    float acc = 0.f;
    for (int x = 0; x < n; ++x) acc = fmaf((float)p[x], __ldg(k + x), acc);
    tmp[i] = acc * (1.0f / 255.0f);
}

__global__ void rs_v_kernel(const float* __restrict__ tmp, __half* __restrict__ dst, long long planes, int H, int ow, int oh,
                            const int* __restrict__ bounds, const float* __restrict__ wt, int ksize, int channels,
                            const float* __restrict__ mean, const float* __restrict__ inv_std) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= planes * oh * ow) return;
    const int x = (int)(i % ow);
    const long long r = i / ow;
    const int yy = (int)(r % oh);
    const long long pl = r / oh;
    const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
    const float* p = tmp + (pl * H + y0) * ow + x;
    const float* k = wt + (long long)yy * ksize;
    float acc = 0.f;
    for (int y = 0; y < n; ++y) acc = fmaf(p[(long long)y * ow], __ldg(k + y), acc);
    const int c = (int)(pl % channels);
    dst[i] = __float2half_rn((acc - mean[c]) * inv_std[c]);
}

// ---- im2col of a stride == kernel patch conv: x fp16 [B][C][G*ps][G*ps] -> rows [B*G*G][kpad], k = (c*ps + ky)*ps + kx
__global__ void patchify_kernel(const __half* __restrict__ x, __half* __restrict__ out, int B, int C, int G, int ps, int kpad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * G * G * kpad;
    if (i >= total) return;
    const int k = (int)(i % kpad);
    const long long row = i / kpad;
    const int px = (int)(row % G), py = (int)((row / G) % G), b = (int)(row / ((long long)G * G));
    __half v = __float2half_rn(0.f);
    if (k < C * ps * ps) {
        const int kx = k % ps, ky = (k / ps) % ps, c = k / (ps * ps);
        const int side = G * ps;
        v = x[(((long long)b * C + c) * side + (py * ps + ky)) * side + (px * ps + kx)];
    }
    out[i] = v;
}

// ---- short-sequence attention: one CTA = 32 queries of one (batch, head); K and V of the head live in shared memory
constexpr int kSmallWarps = 4, kSmallQ = 32;

__global__ void __launch_bounds__(kSmallWarps * 32)
attn_small_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int heads, int d, float scale) {
    extern __shared__ __align__(16) uint8_t smem_small[];
    const int C = heads * d, ld = 3 * C;
    const int kw = (d / 2) | 1;                                  // K / V row pitch in 32-bit words (odd: conflict-free row-parallel reads)
    uint32_t* sK = reinterpret_cast<uint32_t*>(smem_small);
    uint32_t* sV = sK + (size_t)S * kw;
    float* sP = reinterpret_cast<float*>(sV + (size_t)S * kw);  // [warps][S] probabilities
    float* sQ = sP + (size_t)kSmallWarps * S;                    // [warps][d] the warp's current query, pre-scaled
    const int head = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* base = qkv + (long long)b * S * ld + head * d;
    for (int i = threadIdx.x; i < S * (d / 2); i += blockDim.x) {
        const int j = i / (d / 2), e = i % (d / 2);
        sK[(size_t)j * kw + e] = *reinterpret_cast<const uint32_t*>(base + (long long)j * ld + C + 2 * e);
        sV[(size_t)j * kw + e] = *reinterpret_cast<const uint32_t*>(base + (long long)j * ld + 2 * C + 2 * e);
    }
    __syncthreads();
    float* p = sP + (size_t)warp * S;
    float* q = sQ + (size_t)warp * d;
    for (int qi = blockIdx.x * kSmallQ + warp; qi < min(S, (int)(blockIdx.x + 1) * kSmallQ); qi += kSmallWarps) {
        for (int e = lane; e < d; e += 32) q[e] = __half2float(base[(long long)qi * ld + e]) * scale;
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) {
            const uint32_t* kr = sK + (size_t)j * kw;
            float s = 0.f;
            for (int e = 0; e < d / 2; ++e) {
                const float2 kv = __half22float2(*reinterpret_cast<const __half2*>(kr + e));
                s = fmaf(q[2 * e], kv.x, s);
                s = fmaf(q[2 * e + 1], kv.y, s);
            }
            p[j] = s;
            mx = fmaxf(mx, s);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float e = __expf(p[j] - mx);
            p[j] = e;
            sum += e;
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.f / sum;
        for (int e2 = lane; e2 < d / 2; e2 += 32) {              // two output dims per lane per pass
            float ox = 0.f, oy = 0.f;
            for (int j = 0; j < S; ++j) {
                const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(sV + (size_t)j * kw + e2));
                ox = fmaf(p[j], vv.x, ox);
                oy = fmaf(p[j], vv.y, oy);
            }
            *reinterpret_cast<__half2*>(out + ((long long)b * S + qi) * C + head * d + 2 * e2) = __floats2half2_rn(ox * inv, oy * inv);
        }
        __syncwarp();
    }
}

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_resize_separable_f32(const void* src, void* tmp, void* dst, long long planes, int channels, int H, int W, int oh, int ow,
                                        const int* bounds_h, const float* w_h, int ksize_h, const int* bounds_v, const float* w_v,
                                        int ksize_v, const float* mean, const float* inv_std, wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(src && tmp && dst && bounds_h && w_h && bounds_v && w_v && mean && inv_std, "null pointer");
    WIW_REQUIRE(planes > 0 && channels > 0 && planes % channels == 0 && H > 0 && W > 0 && oh > 0 && ow > 0 && ksize_h > 0 && ksize_v > 0, "bad shape");
    const long long n1 = planes * H * ow, n2 = planes * oh * ow;
    rs_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, stream>>>((const uint8_t*)src, (float*)tmp, planes * H, W, ow, bounds_h, w_h, ksize_h);
    WIW_CHECK_CUDA(cudaGetLastError());
    rs_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>((const float*)tmp, (__half*)dst, planes, H, ow, oh, bounds_v, w_v, ksize_v,
                                                                   channels, mean, inv_std);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_patchify(const void* x, void* out, int B, int C, int G, int ps, int kpad, wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(x && out, "null pointer");
    WIW_REQUIRE(B > 0 && C > 0 && G > 0 && ps > 0 && kpad >= C * ps * ps && kpad % 64 == 0, "bad shape");
    const long long total = (long long)B * G * G * kpad;
    patchify_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const __half*)x, (__half*)out, B, C, G, ps, kpad);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_attention_small(const void* qkv, void* out, int batch, int S, int heads, int d, float scale, wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(qkv && out, "null pointer");
    WIW_REQUIRE(batch > 0 && S > 0 && heads > 0 && d > 0 && d % 2 == 0 && d <= 128, "bad shape (even head_dim <= 128)");
    const int kw = (d / 2) | 1;
    const size_t smem = (size_t)2 * S * kw * 4 + (size_t)kSmallWarps * S * 4 + (size_t)kSmallWarps * d * 4;
    WIW_REQUIRE(smem <= 200 * 1024, "sequence too long for the shared-memory attention kernel");
    static unsigned long long attr_set = 0;
    if (first_call_on_device(&attr_set))
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((S + kSmallQ - 1) / kSmallQ, heads, batch);
    attn_small_kernel<<<grid, kSmallWarps * 32, smem, stream>>>((const __half*)qkv, (__half*)out, S, heads, d, scale);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
