// Small HBM-bound kernels around the GEMMs: EDM/Euler sampler glue, embedding glue, channel
// concat, nearest 2x upsample, temporal (seq = F) attention, row softmax, transposes, frame
// quantisation.  All fp16 channels-last, 16-byte vector accesses where the layout allows.
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

// ---------------------------------------------------------------- sampler: build the UNet input
// Pipeline :589-593: cat([latents]*2) -> scale_model_input (x / sqrt(sigma^2+1), scheduler :313-314,
// fp16 tensor / fp32 scalar -> fp32 divide, round to fp16) -> channel-concat with image_latents
// (CFG row 0 = zeros).  Writes channels 0..7 of the 64-channel-padded conv_in input
// [2][F][h][w][64]; channels 8..63 stay zero (the caller zero-fills the buffer once).
__global__ void sampler_prepare_kernel(const __half* __restrict__ lat /*[F][hw][4]*/, const __half* __restrict__ img /*[hw][4]*/,
                                       __half* __restrict__ xin, int F, int hw, float inv_div /*unused*/, float div) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over F*hw
    if (i >= (long long)F * hw) return;
    const int p = (int)(i % hw);
    const uint2 u = *reinterpret_cast<const uint2*>(lat + i * 4);
    const __half* l = reinterpret_cast<const __half*>(&u);
    __half o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = f2h(__fdiv_rn(h2f(l[c]), div));
#pragma unroll
    for (int c = 4; c < 8; ++c) o[c] = __float2half(0.f);
    *reinterpret_cast<uint4*>(xin + i * 64) = *reinterpret_cast<const uint4*>(o);              // uncond row
    const uint2 ui = *reinterpret_cast<const uint2*>(img + (long long)p * 4);
    const __half* im = reinterpret_cast<const __half*>(&ui);
#pragma unroll
    for (int c = 0; c < 4; ++c) o[4 + c] = im[c];
    *reinterpret_cast<uint4*>(xin + ((long long)F * hw + i) * 64) = *reinterpret_cast<const uint4*>(o);   // cond row
}

// ---------------------------------------------------------------- sampler: CFG combine + Euler step
// Pipeline :606-611 and scheduler.step :635-673, reproducing the reference's rounding points:
//   d = fp16(vc - vu); m = fp16(gs_f * d); v = fp16(vu + m)                       (fp16 tensor ops)
//   pred = fp32( fp16(v * c_out) ) + x / (sigma^2+1)                              (fp16*0-dim fp32 -> fp16)
//   deriv = (x - pred) / sigma; x' = fp16(x + deriv * dt)                         (fp32)
// No FMA contraction (explicit _rn intrinsics) so the update is bit-exact vs the CPU path given equal v.
__global__ void sampler_step_kernel(const __half* __restrict__ v /*[2][F][hw][vstride]*/, __half* __restrict__ lat /*[F][hw][4]*/,
                                    const __half* __restrict__ gs /*[F]*/, int F, int hw, int vstride,
                                    float c_out, float sig2p1, float sigma, float dt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * hw) return;
    const int f = (int)(i / hw);
    const __half g = gs[f];
    const uint2 uu = *reinterpret_cast<const uint2*>(v + i * vstride);
    const uint2 uc = *reinterpret_cast<const uint2*>(v + ((long long)F * hw + i) * vstride);
    uint2 ux = *reinterpret_cast<const uint2*>(lat + i * 4);
    const __half* vu = reinterpret_cast<const __half*>(&uu);
    const __half* vc = reinterpret_cast<const __half*>(&uc);
    __half* x = reinterpret_cast<__half*>(&ux);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const __half d = f2h(__fsub_rn(h2f(vc[c]), h2f(vu[c])));
        const __half m = f2h(__fmul_rn(h2f(g), h2f(d)));
        const __half vv = f2h(__fadd_rn(h2f(vu[c]), h2f(m)));
        const float xs = h2f(x[c]);
        const float pred = __fadd_rn(h2f(f2h(__fmul_rn(h2f(vv), c_out))), __fdiv_rn(xs, sig2p1));
        const float deriv = __fdiv_rn(__fsub_rn(xs, pred), sigma);
        x[c] = f2h(__fadd_rn(xs, __fmul_rn(deriv, dt)));
    }
    *reinterpret_cast<uint2*>(lat + i * 4) = ux;
}

// ---------------------------------------------------------------- emb = silu(time + const[r] ), fp16 rounding as the reference
// unet :482-487: emb = emb_only_time + action (fp16 add) + noise (fp16 add); ResnetBlock2D :345 silu(temb).
// te [C] (per step), act [R][C], noise [R][C] -> out [R][C] = fp16(silu(fp16(fp16(te + act) + noise)))
__global__ void emb_silu_kernel(const __half* __restrict__ te, const __half* __restrict__ act, const __half* __restrict__ noise,
                                __half* __restrict__ out, int R, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    const int c = i % C;
    const __half e1 = f2h(h2f(te[c]) + h2f(act[i]));
    const __half e2 = f2h(h2f(e1) + h2f(noise[i]));
    out[i] = f2h(silu_f(h2f(e2)));
}

// ---------------------------------------------------------------- channel concat [M,C1] ++ [M,C2] -> [M,C1+C2]
__global__ void concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o,
                              long long M, int v1, int v2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int vt = v1 + v2;
    if (i >= M * vt) return;
    const long long r = i / vt;
    const int c = (int)(i - r * vt);
    o[i] = (c < v1) ? __ldg(a + r * v1 + c) : __ldg(b + r * v2 + (c - v1));
}

// ---------------------------------------------------------------- nearest 2x upsample [N,H,W,C] -> [N,2H,2W,C]
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int vc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * 2 * H * 2 * W * vc;
    if (i >= total) return;
    const int c = (int)(i % vc);
    long long r = i / vc;
    const int xo = (int)(r % (2 * W)); r /= (2 * W);
    const int yo = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    y[i] = __ldg(x + (((long long)n * H + (yo >> 1)) * W + (xo >> 1)) * vc + c);
}

// ---------------------------------------------------------------- temporal self-attention (sequence = F frames)
// attention.py:736 via AttnProcessor2_0: softmax(q k^T / 8) v per (batch element, pixel, head), d = 64.
// qkv [B][F][S][3C] (q | k | v), out [B][F][S][C].  One CTA = (b, head, tile of TS pixels);
// thread = (pixel, query frame); K and V of the tile staged in smem (padded rows).
template <int F>
__global__ void temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int C, int TS, float scale) {
    extern __shared__ __align__(16) uint8_t smem_t[];
    constexpr int ROWH = 72;                                 // 64 + 8 halves padding -> conflict-free 16-B reads
    __half* sk = reinterpret_cast<__half*>(smem_t);          // [F][TS][ROWH]
    __half* sv = sk + (size_t)F * TS * ROWH;
    const int heads = C / 64;
    const int tile = blockIdx.x, h = blockIdx.y % heads, b = blockIdx.y / heads;
    const int s0 = tile * TS;
    const long long row_stride = 3LL * C;
    const __half* base = qkv + ((long long)b * F * S) * row_stride + h * 64;
    // cooperative load of K, V: (f, s, 8 vec) items
    for (int i = threadIdx.x; i < F * TS * 8; i += blockDim.x) {
        const int vq = i & 7, s = (i >> 3) % TS, f = (i >> 3) / TS;
        if (s0 + s < S) {
            const __half* p = base + ((long long)f * S + s0 + s) * row_stride;
            *reinterpret_cast<uint4*>(sk + ((size_t)f * TS + s) * ROWH + vq * 8) = __ldg(reinterpret_cast<const uint4*>(p + C + vq * 8));
            *reinterpret_cast<uint4*>(sv + ((size_t)f * TS + s) * ROWH + vq * 8) = __ldg(reinterpret_cast<const uint4*>(p + 2 * C + vq * 8));
        }
    }
    __syncthreads();
    const int s = threadIdx.x % TS, fi = threadIdx.x / TS;
    if (fi >= F || s0 + s >= S) return;
    const __half* qp = base + ((long long)fi * S + s0 + s) * row_stride;
    __half2 q[32];
#pragma unroll
    for (int vq = 0; vq < 8; ++vq) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(qp + vq * 8));
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) q[vq * 4 + e] = hh[e];
    }
    float sc[F];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < F; ++j) {
        const __half* kp = sk + ((size_t)j * TS + s) * ROWH;
        float acc = 0.f;
#pragma unroll
        for (int vq = 0; vq < 8; ++vq) {
            const uint4 u = *reinterpret_cast<const uint4*>(kp + vq * 8);
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(q[vq * 4 + e]), k2 = __half22float2(hh[e]);
                acc += a.x * k2.x + a.y * k2.y;
            }
        }
        sc[j] = acc * scale;
        mx = fmaxf(mx, sc[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.f / sum;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
        const __half* vp = sv + ((size_t)j * TS + s) * ROWH;
        const float p = sc[j] * inv;
#pragma unroll
        for (int vq = 0; vq < 8; ++vq) {
            const uint4 u = *reinterpret_cast<const uint4*>(vp + vq * 8);
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 v2 = __half22float2(hh[e]);
                o[vq * 8 + 2 * e] += p * v2.x;
                o[vq * 8 + 2 * e + 1] += p * v2.y;
            }
        }
    }
    __half* op = out + (((long long)b * F + fi) * S + s0 + s) * C + h * 64;
#pragma unroll
    for (int vq = 0; vq < 8; ++vq) {
        uint4 u;
        __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) hh[e] = __floats2half2_rn(o[vq * 8 + 2 * e], o[vq * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(op + vq * 8) = u;
    }
}

// ---------------------------------------------------------------- in-place row softmax (VAE mid-block attention scores)
// p = softmax(scale * s) over each row of [R][S] fp16; one CTA per row, fp32 maths.
__global__ void softmax_rows_kernel(__half* __restrict__ s, int S, float scale) {
    __shared__ float red[32];
    __half* row = s + (long long)blockIdx.x * S;
    const int nv = S >> 3;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hh[e]); mx = fmaxf(mx, fmaxf(f.x, f.y)); }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hh[e]);
            sum += __expf((f.x - mx) * scale) + __expf((f.y - mx) * scale);
        }
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
    const float inv = 1.f / sum;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
        __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hh[e]);
            hh[e] = __floats2half2_rn(__expf((f.x - mx) * scale) * inv, __expf((f.y - mx) * scale) * inv);
        }
        *reinterpret_cast<uint4*>(row + i * 8) = u;
    }
}

// ---------------------------------------------------------------- batched 2-D transpose [B][R][C] -> [B][C][R] (fp16)
__global__ void transpose_kernel(const __half* __restrict__ x, __half* __restrict__ y, int R, int C) {
    __shared__ __half tile[32][34];
    const int b = blockIdx.z;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    x += (long long)b * R * C; y += (long long)b * R * C;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = x[(long long)r * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) y[(long long)c * R + r] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------- decoded frames -> uint8 RGB, NCHW
// postprocess_video (video_processor.py:89-113 -> image_processor.py:239,147): (x/2+0.5).clamp(0,1) -> round(x*255).
// x [F][H][W][cs] fp16 (first 3 channels) -> out [F][3][H][W] uint8
__global__ void frames_to_u8_kernel(const __half* __restrict__ x, uint8_t* __restrict__ out, long long F, long long HW, int cs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * HW) return;
    const long long f = i / HW, p = i - f * HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = h2f(x[i * cs + c]);                     // decode_latents casts frames to fp32 (pipeline :307)
        v = fminf(fmaxf(__fadd_rn(__fdiv_rn(v, 2.f), 0.5f), 0.f), 1.f);
        out[(f * 3 + c) * HW + p] = (uint8_t)rintf(__fmul_rn(v, 255.f));
    }
}


// ---------------------------------------------------------------- VAE tail: time_conv_out + quantise
// TemporalDecoder.forward tail (autoencoder_kl_temporal_decoder.py:154-160): Conv3d(3,3,(3,1,1)) over
// frames (zero padded at the clip ends), fp16 result; then decode_latents .float() (pipeline :307) and
// postprocess (x/2+0.5).clamp(0,1) -> round(x*255) (image_processor.py:239,147).
// x [F][HW][cs] fp16 (conv_out result in channels 0..2); w [co][ci][kt] fp32; outputs NCHW per frame.
__global__ void vae_finish_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  uint8_t* __restrict__ out_u8, float* __restrict__ out_f32, int F, long long HW, int cs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * HW) return;
    const int f = (int)(i / HW);
    const long long p = i - (long long)f * HW;
    float acc[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        const int ff = f + kt - 1;
        if (ff < 0 || ff >= F) continue;
        const __half* px = x + ((long long)ff * HW + p) * cs;
        const uint2 u = *reinterpret_cast<const uint2*>(px);
        const __half* hx = reinterpret_cast<const __half*>(&u);
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float xv = h2f(hx[ci]);
#pragma unroll
            for (int co = 0; co < 3; ++co) acc[co] += w[(co * 3 + ci) * 3 + kt] * xv;
        }
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) {
        const float v = h2f(f2h(acc[co]));                 // the reference's conv output is fp16
        if (out_f32) out_f32[((long long)f * 3 + co) * HW + p] = v;
        if (out_u8) {
            const float q = fminf(fmaxf(__fadd_rn(__fdiv_rn(v, 2.f), 0.5f), 0.f), 1.f);
            out_u8[((long long)f * 3 + co) * HW + p] = (uint8_t)rintf(__fmul_rn(q, 255.f));
        }
    }
}

// ---------------------------------------------------------------- layout converters at the model boundary
// src NCHW-like [N][C][P] (fp16) -> dst [N][P][cs] channels-last (channels >= C zero)
__global__ void nchw_to_nhwc_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long long N, int C, long long P, int cs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * P * cs) return;
    const int c = (int)(i % cs);
    const long long np = i / cs, p = np % P, n = np / P;
    dst[i] = (c < C) ? src[(n * C + c) * P + p] : __float2half(0.f);
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long long N, int C, long long P, int cs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C * P) return;
    const long long p = i % P, nc = i / P;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    dst[i] = src[(n * P + p) * cs + c];
}

}  // namespace wiw

using namespace wiw;
#define STREAM(s) static_cast<cudaStream_t>(s)
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int wiw_sampler_prepare(const void* latents, const void* image_latents, void* unet_in, int F, int hw,
                                   float div, wiw_stream_t st) {
    WIW_REQUIRE(latents && image_latents && unet_in && F > 0 && hw > 0, "bad args");
    sampler_prepare_kernel<<<nblk((long long)F * hw, 256), 256, 0, STREAM(st)>>>(
        (const __half*)latents, (const __half*)image_latents, (__half*)unet_in, F, hw, 0.f, div);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_sampler_step(const void* v, void* latents, const void* gs, int F, int hw, int vstride, float c_out,
                                float sig2p1, float sigma, float dt, wiw_stream_t st) {
    WIW_REQUIRE(v && latents && gs && F > 0 && hw > 0 && vstride >= 4 && vstride % 4 == 0, "bad args");
    sampler_step_kernel<<<nblk((long long)F * hw, 256), 256, 0, STREAM(st)>>>(
        (const __half*)v, (__half*)latents, (const __half*)gs, F, hw, vstride, c_out, sig2p1, sigma, dt);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_emb_silu(const void* te, const void* act, const void* noise, void* out, int R, int C, wiw_stream_t st) {
    WIW_REQUIRE(te && act && noise && out, "null pointer");
    emb_silu_kernel<<<nblk((long long)R * C, 256), 256, 0, STREAM(st)>>>((const __half*)te, (const __half*)act, (const __half*)noise, (__half*)out, R, C);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_concat_channels(const void* a, const void* b, void* out, long long M, int C1, int C2, wiw_stream_t st) {
    WIW_REQUIRE(a && b && out && C1 % 8 == 0 && C2 % 8 == 0, "bad args");
    concat_kernel<<<nblk(M * ((C1 + C2) / 8), 256), 256, 0, STREAM(st)>>>((const uint4*)a, (const uint4*)b, (uint4*)out, M, C1 / 8, C2 / 8);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_upsample_nearest2x(const void* x, void* y, int N, int H, int W, int C, wiw_stream_t st) {
    WIW_REQUIRE(x && y && C % 8 == 0, "bad args");
    upsample2x_kernel<<<nblk((long long)N * 4 * H * W * (C / 8), 256), 256, 0, STREAM(st)>>>((const uint4*)x, (uint4*)y, N, H, W, C / 8);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int F>
static int launch_tattn(const void* qkv, void* out, int B, int S, int C, cudaStream_t st) {
    int TS = 256 / F;                       // pixels per CTA (threads = TS * F <= 256)
    if (TS > S) TS = S;
    if (TS < 1) TS = 1;
    const int threads = ((TS * F + 31) / 32) * 32;
    const size_t smem = (size_t)2 * F * TS * 72 * sizeof(__half);
    auto kern = temporal_attn_kernel<F>;
    WIW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((S + TS - 1) / TS, B * (C / 64));
    kern<<<grid, threads, smem, st>>>((const __half*)qkv, (__half*)out, S, C, TS, 0.125f);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_attention_temporal_mma(const void* qkv, void* out, int B, int F, int S, int C, void* stream);   // wiw_tattn.cu

extern "C" int wiw_attention_temporal(const void* qkv, void* out, int B, int F, int S, int C, wiw_stream_t st) {
    WIW_REQUIRE(qkv && out && C % 64 == 0, "bad args");
    if (F <= 16) return wiw_attention_temporal_mma(qkv, out, B, F, S, C, st);     // warp-level tensor-core kernel
    if (F == 25) return launch_tattn<25>(qkv, out, B, S, C, STREAM(st));             // SVD-XT clip length: scalar kernel
    wiw::set_last_error("temporal attention is built for F <= 16 (tensor-core kernel) and F == 25");
    return 2;
}

extern "C" int wiw_softmax_rows(void* s, long long rows, int S, float scale, wiw_stream_t st) {
    WIW_REQUIRE(s && S % 8 == 0 && rows < (1LL << 31), "bad args");
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, STREAM(st)>>>((__half*)s, S, scale);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_transpose(const void* x, void* y, int batch, int R, int C, wiw_stream_t st) {
    WIW_REQUIRE(x && y, "null pointer");
    dim3 grid((C + 31) / 32, (R + 31) / 32, batch), block(32, 8);
    transpose_kernel<<<grid, block, 0, STREAM(st)>>>((const __half*)x, (__half*)y, R, C);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_frames_to_u8(const void* x, void* out, long long F, long long HW, int cs, wiw_stream_t st) {
    WIW_REQUIRE(x && out && cs >= 3, "bad args");
    frames_to_u8_kernel<<<nblk(F * HW, 256), 256, 0, STREAM(st)>>>((const __half*)x, (uint8_t*)out, F, HW, cs);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_nchw_to_nhwc(const void* src, void* dst, long long N, int C, long long P, int cs, wiw_stream_t st) {
    WIW_REQUIRE(src && dst && cs >= C, "bad args");
    nchw_to_nhwc_kernel<<<nblk(N * P * cs, 256), 256, 0, STREAM(st)>>>((const __half*)src, (__half*)dst, N, C, P, cs);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_nhwc_to_nchw(const void* src, void* dst, long long N, int C, long long P, int cs, wiw_stream_t st) {
    WIW_REQUIRE(src && dst && cs >= C, "bad args");
    nhwc_to_nchw_kernel<<<nblk(N * C * P, 256), 256, 0, STREAM(st)>>>((const __half*)src, (__half*)dst, N, C, P, cs);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wiw_vae_finish(const void* x, const float* w, const float* bias, void* out_u8, float* out_f32, int F,
                              long long HW, int cs, wiw_stream_t st) {
    WIW_REQUIRE(x && w && bias && (out_u8 || out_f32) && cs >= 4 && cs % 4 == 0, "bad args");
    vae_finish_kernel<<<nblk((long long)F * HW, 256), 256, 0, STREAM(st)>>>((const __half*)x, w, bias, (uint8_t*)out_u8, out_f32, F, HW, cs);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
