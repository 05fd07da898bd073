// Pillow-exact bicubic resize of uint8 planes (horizontal then vertical pass, 22-bit fixed-point weights,
// 8-bit intermediate, clip8) + the reference's requantisation LUT.  Replaces the per-frame CPU
// `img.resize(save_size, Image.BICUBIC)` + to_tensor + (clip*255).astype(uint8) of
// downstream/api_models/__init__.py:159-166,229.  Integer arithmetic only: bit-exact vs PIL given the host-built
// coefficient tables (wiw_b200/resize.py).
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8_fx(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src [P][H][W] -> tmp [P][H][ow]
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, long long rows, int W, int ow,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ow) return;
    const int xx = (int)(i % ow);
    const long long row = i / ow;
    const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
    const uint8_t* p = src + row * W + x0;
    const int* k = kk + (long long)xx * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < n; ++x) acc += (int)p[x] * __ldg(k + x);
    tmp[i] = clip8_fx(acc);
}

// tmp [P][H][ow] -> dst [P][oh][ow], then LUT
__global__ void resize_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ dst, long long planes, int H, int ow, int oh,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, const uint8_t* __restrict__ lut) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= planes * oh * ow) return;
    const int x = (int)(i % ow);
    const long long r = i / ow;
    const int yy = (int)(r % oh);
    const long long pl = r / oh;
    const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
    const uint8_t* p = tmp + (pl * H + y0) * ow + x;
    const int* k = kk + (long long)yy * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < n; ++y) acc += (int)p[(long long)y * ow] * __ldg(k + y);
    dst[i] = __ldg(lut + clip8_fx(acc));
}

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_resize_bicubic_u8(const void* src, void* tmp, void* dst, long long planes, int H, int W, int oh, int ow,
                                     const int* bounds_h, const int* kk_h, int ksize_h, const int* bounds_v, const int* kk_v,
                                     int ksize_v, const void* lut, wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(src && tmp && dst && bounds_h && kk_h && bounds_v && kk_v && lut, "null pointer");
    WIW_REQUIRE(planes > 0 && H > 0 && W > 0 && oh > 0 && ow > 0 && ksize_h > 0 && ksize_v > 0, "bad shape");
    const long long n1 = planes * H * ow, n2 = planes * oh * ow;
    resize_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, stream>>>((const uint8_t*)src, (uint8_t*)tmp, planes * H, W, ow, bounds_h, kk_h, ksize_h);
    WIW_CHECK_CUDA(cudaGetLastError());
    resize_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>((const uint8_t*)tmp, (uint8_t*)dst, planes, H, ow, oh, bounds_v, kk_v, ksize_v, (const uint8_t*)lut);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
