// Spatial self-attention for the SVD transformer blocks on sm_100a:
//   out = softmax(Q K^T / sqrt(64)) V   per (image, head), head_dim 64, no mask, S up to 9216
// (reference: F.scaled_dot_product_attention in AttnProcessor2_0, DP/models/attention_processor.py:2383,
//  called from BasicTransformerBlock.attn1, DP/models/attention.py:507).
//
// One CTA = 128 query rows of one (image, head); two CTAs per SM overlap each other's MMA and
// softmax phases.  Per KV tile of 128 keys:
//   warp 0   TMA: Q once, then K_j / V_j tiles ([128][64] fp16, SWIZZLE_128B) through a 3-slot ring
//   warp 1   tcgen05.mma: S = Q K_j^T  (M128 N128 K64, both K-major)  -> TMEM cols [0,128)
//                         O += P_j V_j (M128 N64 K128, A = P from smem K-major, B = V MN-major) -> TMEM cols [128,192)
//   warps 2-5 (one thread per query row): tcgen05.ld S, running max / sum, exp2, P -> fp16 smem
//            (swizzled, A operand), lazy rescale of O in TMEM when the row max grew by > 2^8,
//            final O / l -> fp16 -> smem -> TMA store (clipped to the image's S rows).
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

constexpr int kQ = 128, kKV = 128, kD = 64;
constexpr int kTileBytes = 128 * 64 * 2;        // 16 KB: Q, K, V tiles and each half of P
constexpr int kRing = 3;
constexpr int kAttnThreads = 6 * 32;
constexpr int kAttnSmem = 1024 + kTileBytes * (1 + kRing + 2) + 128;
constexpr uint32_t kTmemCols = 256;              // S: cols [0,128), O: cols [128,192)

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnArgs {
    int S, nkv;
    float scale_log2;                            // (1/sqrt(d)) * log2(e)
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_spatial_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                    const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sRing = sQ + kTileBytes;
    const uint32_t sP = sRing + kRing * kTileBytes;              // 2 x 16 KB ([128][64] k-halves); reused for O staging
    const uint32_t bars = sP + 2 * kTileBytes;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing]
    const uint32_t bar_empty = bar_full + 8 * kRing;             // [kRing]
    const uint32_t bar_s_full = bar_empty + 8 * kRing;
    const uint32_t bar_s_free = bar_s_full + 8;
    const uint32_t bar_p_full = bar_s_free + 8;
    const uint32_t bar_pv_done = bar_p_full + 8;
    const uint32_t tmem_slot = bar_pv_done + 8;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));       // generic pointer to `base`

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_s_full, 1);
        mbar_init(bar_s_free, 128);
        mbar_init(bar_p_full, 128);
        mbar_init(bar_pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
    const uint32_t tS = tmem_base, tO = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_q, kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            for (int i = 0; i < 2 * nkv; ++i) {
                const int st = i % kRing;
                const uint32_t ph = (uint32_t)(i / kRing) & 1u;
                mbar_wait(bar_empty + 8 * st, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int j = i >> 1;
                const int col = ((i & 1) ? 2 * C : C) + head * kD;              // even: K_j, odd: V_j
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, j * kKV, img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
            const uint32_t idesc_pv = make_idesc_f16(kD, 0, 1);               // B (= V) is MN-major
            mbar_wait(bar_q, 0);
            for (int j = 0; j < nkv; ++j) {
                // ---- S = Q K_j^T
                int i = 2 * j, st = i % kRing;
                mbar_wait(bar_full + 8 * st, (uint32_t)(i / kRing) & 1u);
                if (j > 0) mbar_wait(bar_s_free, (uint32_t)(j - 1) & 1u);     // softmax finished reading S_{j-1}
                tc_fence_after();
                {
                    const uint64_t dq = make_desc_sw128(sQ), dk = make_desc_sw128(sRing + st * kTileBytes);
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk) umma_f16_ss(tS, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * st);
                umma_commit(bar_s_full);
                // ---- O += P_j V_j
                i = 2 * j + 1; st = i % kRing;
                mbar_wait(bar_full + 8 * st, (uint32_t)(i / kRing) & 1u);
                mbar_wait(bar_p_full, (uint32_t)j & 1u);
                tc_fence_after();
                {
                    const uint64_t dv = make_desc_sw128(sRing + st * kTileBytes);
#pragma unroll
                    for (int kk = 0; kk < kKV / 16; ++kk) {
                        const uint64_t dp = make_desc_sw128(sP + (kk >> 2) * kTileBytes) + 2 * (kk & 3);
                        umma_f16_ss(tO, dp, dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                umma_commit(bar_empty + 8 * st);
                umma_commit(bar_pv_done);
            }
        }
    } else {
        // ================= softmax / correction / epilogue: one thread per query row =================
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pP = gen + (sP - base) + row * 128;
        float m_ref = -INFINITY, l = 0.f;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int valid = min(kKV, a.S - j * kKV);              // keys beyond S are zero-filled: mask them
            mbar_wait(bar_s_full, (uint32_t)j & 1u);
            tc_fence_after();
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tS + lane_sel + ch * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float s = (ch * 32 + e < valid) ? __uint_as_float(r[e]) : -INFINITY;
                    mx = fmaxf(mx, s);
                }
            }
            // lazy rescale: keep exponent reference m_ref unless some row of the warp grew by > 8 (log2 units)
            float alpha = 1.f;
            const bool grow = (mx - m_ref) * c > 8.0f;
            const bool any_grow = __any_sync(0xffffffffu, grow || (j == 0));
            if (any_grow) {
                const float m_new = fmaxf(m_ref, mx);
                alpha = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_new) * c);
                m_ref = m_new;
                l *= alpha;
            }
            if (j > 0) {
                // P buffer and O are owned by the tensor core until PV_{j-1} has retired
                mbar_wait(bar_pv_done, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (any_grow) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + lane_sel + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + lane_sel + ch * 16, r);
                    }
                    tmem_st_wait();
                }
            }
            // pass 2: p = exp2((s - m_ref) * c) -> fp16 -> P smem (K-major, SWIZZLE_128B, two 64-key halves)
            const float mc = m_ref * c;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tS + lane_sel + ch * 32, r);
                tmem_ld_wait();
                if (ch == 3) {                                   // last read of S_j: release it for QK_{j+1}
                    tc_fence_before();
                    mbar_arrive(bar_s_free);
                }
                uint8_t* dst = pP + (ch >> 1) * kTileBytes;
#pragma unroll
                for (int v8 = 0; v8 < 4; ++v8) {
                    uint4 u;
                    __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k0 = ch * 32 + v8 * 8 + 2 * e;
                        float p0 = (k0 < valid) ? fast_exp2(__uint_as_float(r[v8 * 8 + 2 * e]) * c - mc) : 0.f;
                        float p1 = (k0 + 1 < valid) ? fast_exp2(__uint_as_float(r[v8 * 8 + 2 * e + 1]) * c - mc) : 0.f;
                        hh[e] = __floats2half2_rn(p0, p1);
                        const float2 back = __half22float2(hh[e]);   // sum what the tensor core will see
                        l += back.x + back.y;
                    }
                    const uint32_t chunk16 = (uint32_t)((ch & 1) * 4 + v8);
                    *reinterpret_cast<uint4*>(dst + ((chunk16 ^ sw) << 4)) = u;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full);
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128) -> TMA store
        mbar_wait(bar_pv_done, (uint32_t)(nkv - 1) & 1u);
        tc_fence_after();
        const float inv = 1.f / l;
        uint8_t* dst = gen + (sP - base) + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + lane_sel + ch * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)(ch * 2 + v8);
                *reinterpret_cast<uint4*>(dst + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sP), "r"(head * kD), "r"(q0), "r"(img) : "memory");
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_attention_spatial(const void* qkv, void* out, int images, int S, int C, wiw_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(qkv && out, "null pointer");
    WIW_REQUIRE(C % 64 == 0 && S > 0 && images > 0, "bad shape");
    CUtensorMap tm_qkv, tm_out;
    {
        uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)S, (uint64_t)images};
        uint64_t strides[2] = {(uint64_t)3 * C * 2, (uint64_t)S * 3 * C * 2};
        uint32_t box[3] = {64, 128, 1};
        if (encode_tmap(&tm_qkv, qkv, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)S, (uint64_t)images};
        uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)S * C * 2};
        uint32_t box[3] = {64, 128, 1};
        if (encode_tmap(&tm_out, out, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    AttnArgs a;
    a.S = S;
    a.nkv = (S + kKV - 1) / kKV;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    static bool attr_set = false;
    if (!attr_set) {
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        attr_set = true;
    }
    dim3 grid((S + kQ - 1) / kQ, C / 64, images);
    attn_spatial_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tm_qkv, tm_out, a, C);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
