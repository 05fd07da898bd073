// Spatial self-attention for the SVD transformer blocks on sm_100a:
//   out = softmax(Q K^T / sqrt(64)) V   per (image, head), head_dim 64, no mask, S up to 9216
// (reference: F.scaled_dot_product_attention in AttnProcessor2_0, DP/models/attention_processor.py:2383,
//  called from BasicTransformerBlock.attn1, DP/models/attention.py:507).
//
// One CTA (one per SM) = 256 query rows (two 128-row tiles) of one (image, head); the two tiles
// ping-pong so the tensor core works on one while the other is in softmax.  Per KV tile of 128 keys:
//   warp 0     TMA: Q0,Q1 once, then K_j / V_j tiles ([128][64] fp16, SWIZZLE_128B) through a 4-slot ring
//   warp 1     tcgen05.mma issuer: S_t = Q_t K_j^T (M128 N128 K64, K-major x K-major) -> TMEM,
//              O_t += P_t V_j (M128 N64 K128, A = P_t from smem, B = V consumed MN-major from its TMA tile)
//   warps 2-5  softmax of tile 0, warps 6-9 softmax of tile 1 (one thread per query row):
//              tcgen05.ld S (2 passes: max, then exp2), P -> fp16 -> swizzled smem, lazy rescale of O in TMEM
//              when some row max of the warp grew by more than 2^8.  The row sum l is NOT accumulated by the
//              softmax threads: the PV MMA runs with N = 80 where B columns 64..79 come from a tile of ones, so
//              O[:, 64] = sum_k P[:, k] in fp32 for free.  Final O / l -> fp16 -> TMA store.
#include "wiw_common.cuh"
#include "../../include/wiw.h"

#include <cstdlib>

namespace wiw {

constexpr int kQ = 128, kKV = 128, kD = 64;
constexpr int kTileBytes = 128 * 64 * 2;        // 16 KB: a Q / K / V tile and each 64-key half of P
constexpr int kRing = 4;
constexpr int kAttnThreads = 10 * 32;
constexpr int kAttnSmem = 1024 + kTileBytes * (2 + kRing + 1 + 4) + 256;   // Q0,Q1 | K/V ring | ones | P0,P1
constexpr uint32_t kTmemCols = 512;              // S0 [0,128) S1 [128,256) O0 [256,336) O1 [384,464)
constexpr int kOCols = 80;                       // 64 output columns + 16 copies of the row sum l (P @ ones)
constexpr float kRescaleThreshold = 8.0f;        // log2 units (P <= 2^8 fits fp16 comfortably)

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for x <= ~2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax, max rel. error 7.5e-5 -- below the
// fp16 rounding of P).  Offloads part of the exponentials from the MUFU pipe, which bounds the softmax warps.
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -26.0f);                              // 2^-26 rounds to +0 in the fp16 pack (half of the smallest subnormal is 2^-25):
                                                       // masked (-inf) and strongly negative scores contribute exactly 0, like MUFU.EX2
    const float xr = x + 12582912.0f;                  // 1.5 * 2^23: round(x) lands in the low mantissa bits
    const float f = x - (xr - 12582912.0f);            // in [-0.5, 0.5]
    const float p = fmaf(fmaf(fmaf(0.05517167f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
// POLY template parameter: of every 4 (pairs of) exponentials, this many use exp2_poly (0 = all on the MUFU pipe)

struct AttnArgs {
    int S, nkv;
    float scale_log2;                            // (1/sqrt(d)) * log2(e)
    long long* trace;                            // development aid (wiw_debug_attn_trace): clock64 stamps of CTA 0, else null
    int dbg;                                     // development aid (WIW_ATTN_DBG, ts3 only; results are garbage): bit 0 = softmax threads skip
                                                 // their TMEM loads / exponentials / stores, bit 1 = the MMA warp skips the MMAs (commits only)
};
constexpr int kTraceIters = 8, kTraceFirst = 16, kTraceSlots = 16;      // stamps of KV tiles [16, 24): [iter][slot]
__device__ __forceinline__ void trace_stamp(const AttnArgs& a, int j, int slot) {
    if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j >= kTraceFirst && j < kTraceFirst + kTraceIters &&
        (threadIdx.x & 31) == 0)
        a.trace[(j - kTraceFirst) * kTraceSlots + slot] = clock64();
}

// keys at or beyond `valid` (zero-filled by TMA past the image's S rows) -> -inf before max / exp
__device__ __forceinline__ void mask_chunk(uint32_t* r, int base, int valid) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
        if (base + e >= valid) r[e] = 0xff800000u;
}

__device__ __forceinline__ float row_max_chunk(const uint32_t* r, float mx) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[e]), __uint_as_float(r[e + 1])));
    return mx;
}

// 32 scores -> 32 fp16 probabilities p = 2^(s*c - mc) written to the swizzled P tile
// (fp32 MUFU.EX2 + one F2FP pack per pair: cheaper than ex2.f16x2, which splits into two MUFU + a PRMT)
template <int kPolyPairs>
__device__ __forceinline__ void exp_chunk(const uint32_t* r, int ch, float c, float mc, uint8_t* dst, uint32_t sw) {
#pragma unroll
    for (int v8 = 0; v8 < 4; ++v8) {
        uint32_t p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a0 = fmaf(__uint_as_float(r[v8 * 8 + 2 * e]), c, -mc);
            const float a1 = fmaf(__uint_as_float(r[v8 * 8 + 2 * e + 1]), c, -mc);
            if (e < kPolyPairs) p[e] = cvt_f16x2(exp2_poly(a0), exp2_poly(a1));
            else p[e] = cvt_f16x2(fast_exp2(a0), fast_exp2(a1));
        }
        const uint32_t chunk16 = (uint32_t)((ch & 1) * 4 + v8);
        *reinterpret_cast<uint4*>(dst + ((chunk16 ^ sw) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_spatial_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                    const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRing * kTileBytes;           // [128][64] fp16 ones: the extra N-atom of the PV B operand
    const uint32_t sP = sOnes + kTileBytes;                      // tile t: 2 x 16 KB at sP + t*32 KB; reused for O staging
    const uint32_t bars = sP + 4 * kTileBytes;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing]
    const uint32_t bar_empty = bar_full + 8 * kRing;             // [kRing]
    const uint32_t bar_s_full = bar_empty + 8 * kRing;           // [2]
    const uint32_t bar_s_free = bar_s_full + 16;                 // [2]
    const uint32_t bar_p_full = bar_s_free + 16;                 // [2]
    const uint32_t bar_pv_done = bar_p_full + 16;                // [2]
    const uint32_t tmem_slot = bar_pv_done + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));       // generic pointer to `base`

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s_full + 8 * t, 1);
            mbar_init(bar_s_free + 8 * t, 128);
            mbar_init(bar_p_full + 8 * t, 128);
            mbar_init(bar_pv_done + 8 * t, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer ================================
        // whole warp in the loop, one elected lane issues (a TMA / tcgen05 instruction under a divergent `lane == 0`
        // branch gets an ELECT + BRA.U.ANY serialisation loop from ptxas: ~100 cycles each, profiles/r01_mma_cost.md)
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_qkv, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        for (int i = 0; i < 2 * nkv; ++i) {                     // item 2j = K_j, 2j+1 = V_j
            const int st = i % kRing;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(i / kRing) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int col = ((i & 1) ? 2 * C : C) + head * kD;
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, (i >> 1) * kKV, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        {
            const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
            const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);           // B (= [V | ones]) is MN-major, N = 80
            auto issue_qk = [&](int t, int stage) {
                const uint64_t dq = make_desc_sw128(sQ + t * kTileBytes), dk = make_desc_sw128(sRing + stage * kTileBytes);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_f16_ss(tmem_base + t * 128, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                    umma_commit(bar_s_full + 8 * t);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int t, int stage, bool first) {
                // MN-major B: N-atom 0 (columns 0..63) = the V tile, N-atom 1 (columns 64..79) = the ones tile at +LBO;
                // K groups of 8 rows at SBO = 1024 B.  O[:, 64..79] therefore accumulates the row sums of P in fp32.
                const uint32_t v_addr = sRing + stage * kTileBytes;
                const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                    (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
                const uint64_t dp0 = make_desc_sw128(sP + t * 2 * kTileBytes);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < kKV / 16; ++kk) {
                        const uint64_t dp = dp0 + (uint64_t)((kk >> 2) * (kTileBytes >> 4) + 2 * (kk & 3));
                        umma_f16_ss(tmem_base + 256 + t * 128, dp, dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_pv_done + 8 * t);
                }
                __syncwarp();
            };
            mbar_wait(bar_q, 0);
            mbar_wait(bar_full + 0, 0);                                       // K_0 in ring slot 0
            tc_fence_after();
            issue_qk(0, 0);
            issue_qk(1, 0);
            if (elect_one()) umma_commit(bar_empty + 0);
            __syncwarp();
            for (int j = 0; j < nkv; ++j) {
                const bool more = (j + 1 < nkv);
                const int ik = 2 * j + 2, iv = 2 * j + 1;
                const int sk = ik % kRing, sv = iv % kRing;
                if (more) mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRing) & 1u);      // K_{j+1}
                mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRing) & 1u);                // V_j
                for (int t = 0; t < 2; ++t) {
                    if (more) {
                        mbar_wait(bar_s_free + 8 * t, (uint32_t)j & 1u);      // softmax t finished reading S_t of tile j
                        tc_fence_after();
                        issue_qk(t, sk);
                    }
                    mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);          // P_t of tile j is in smem
                    tc_fence_after();
                    issue_pv(t, sv, j == 0);
                }
                if (elect_one()) {
                    if (more) umma_commit(bar_empty + 8 * sk);
                    umma_commit(bar_empty + 8 * sv);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= softmax / correction / epilogue: one thread per query row of tile t =================
        const int t = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_sel, tO = tmem_base + 256 + t * 128 + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pP = gen + (sP - base) + t * 2 * kTileBytes + row * 128;
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int valid = a.S - j * kKV;                       // keys beyond S are zero-filled by TMA: mask them
            const bool masked = valid < kKV;
            mbar_wait(bar_s_full + 8 * t, (uint32_t)j & 1u);
            tc_fence_after();
            // pass 1: row max.  The TMEM load of chunk ch+1 is in flight while chunk ch is reduced.
            float mx = -INFINITY;
            if (!masked) {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(tS, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 32, rb);
                mx = row_max_chunk(ra, mx);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 64, ra);
                mx = row_max_chunk(rb, mx);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 96, rb);
                mx = row_max_chunk(ra, mx);
                tmem_ld_wait();
                mx = row_max_chunk(rb, mx);
            } else {                                             // ragged last KV tile (S % 128 != 0): simple rolled loop
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tS + ch * 32, r);
                    tmem_ld_wait();
                    mask_chunk(r, ch * 32, valid);
                    mx = row_max_chunk(r, mx);
                }
            }
            float alpha = 1.f;
            const bool grow = (mx - m_ref) * c > kRescaleThreshold;
            const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
            if (any_grow) {
                const float m_new = fmaxf(m_ref, mx);
                alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                m_ref = m_new;
            }
            if (j > 0) {
                // the P_t buffer and O_t belong to the tensor core until PV of tile j-1 has retired
                mbar_wait(bar_pv_done + 8 * t, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (any_grow) {
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {      // 64 output columns + the row-sum columns
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                    tmem_st_wait();
                }
            }
            // pass 2: p = 2^((s - m_ref) * c) in packed fp16 -> P smem (K-major, SWIZZLE_128B, two 64-key halves)
            const float mc = m_ref * c;
            if (!masked) {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(tS, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 32, rb);
                exp_chunk<POLY>(ra, 0, c, mc, pP, sw);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 64, ra);
                exp_chunk<POLY>(rb, 1, c, mc, pP, sw);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 96, rb);
                exp_chunk<POLY>(ra, 2, c, mc, pP + kTileBytes, sw);
                tmem_ld_wait();
                tc_fence_before();                               // last read of S_t: release it for the next QK
                mbar_arrive(bar_s_free + 8 * t);
                exp_chunk<POLY>(rb, 3, c, mc, pP + kTileBytes, sw);
            } else {
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tS + ch * 32, r);
                    tmem_ld_wait();
                    if (ch == 3) {
                        tc_fence_before();
                        mbar_arrive(bar_s_free + 8 * t);
                    }
                    mask_chunk(r, ch * 32, valid);
                    exp_chunk<POLY>(r, ch, c, mc, pP + (ch >> 1) * kTileBytes, sw);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * t);
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128, first half of this tile's P buffer) -> TMA store
        mbar_wait(bar_pv_done + 8 * t, (uint32_t)(nkv - 1) & 1u);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);                      // row sum l = (P @ ones), fp32 from the tensor core
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + ch * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)(ch * 2 + v8);
                *reinterpret_cast<uint4*>(pP + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (qd == 2 && lane == 0) {                              // first warp of each softmax group (warps 2 and 6)
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sP + t * 2 * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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


// =====================================================================================================================
// TS-mode variant: both A operands of the two GEMMs live in TENSOR MEMORY, so the only shared-memory operand traffic
// of the tensor pipe is the K / V tile itself.
//
// Why: a tcgen05.mma reads its shared-memory operands at ~64 B per cycle per SM (profiles/r02_mma_cost_v2.txt).  With A and B
// both in shared memory an M128 N128 K16 `Q K^T` step moves 4 KB + 4 KB = 128 cycles for 64 cycles of tensor work, a
// `P V` step (N = 80) 4 KB + 2.5 KB = 102 cycles for 40: the SS kernel above spends 2 x (4 x 128 + 8 x 102) = 2 656 cycles per KV
// tile on operand fetch (measured: 2 960 per iteration) while the tensor pipe is busy for 1 152.  Here
//   * Q_t is copied once per CTA from its TMA tile into TMEM (32 columns per tile) by the softmax threads;
//   * P_t is written by the softmax threads straight into TMEM with tcgen05.st, packed fp16, ALIASED onto the first 64 columns
//     of its own S_t (a thread only overwrites score columns it has already read: P chunk c lands in columns [16c, 16c+16),
//     inside score chunks <= c/2);
// so `Q K^T` fetches 4 KB (64 cycles = its tensor time) and `P V` 2.5 KB (40 cycles = its tensor time) per instruction.
// The tensor pipe executes one thread's MMAs in issue order, so the MMA warp issues PV(t, j) and then QK(t, j+1) -- which
// overwrites S_t / P_t -- back to back without waiting in between, and the single commit after QK(t, j+1) tells the softmax
// threads both that S_t(j+1) is ready and that O_t is idle again (PV(t, j) retired): no separate "S free" / "PV done" barriers.
//
// TMEM columns (every operand starts on a multiple of 32): S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,336)  Q0 [352,384)  O1 [384,464)  Q1 [480,512)
constexpr int kRingTS = 6;
constexpr int kAttnSmemTS = 1024 + kTileBytes * (2 + kRingTS + 1) + 256;    // Q0,Q1 (later: O staging) | K/V ring | ones
constexpr uint32_t kColO = 256, kColQ = 352, kColStride = 128;     // tile t: O at kColO + t*128, Q at kColQ + t*128

// 32 scores -> 16 packed fp16x2 probabilities p = 2^(s*c - mc); word i = (key 2i | key 2i+1 << 16)
template <int kPolyPairs>
__device__ __forceinline__ void exp_pack(const uint32_t* r, float c, float mc, uint32_t* p) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a0 = fmaf(__uint_as_float(r[2 * i]), c, -mc);
        const float a1 = fmaf(__uint_as_float(r[2 * i + 1]), c, -mc);
        if ((i & 3) < kPolyPairs) p[i] = cvt_f16x2(exp2_poly(a0), exp2_poly(a1));
        else p[i] = cvt_f16x2(fast_exp2(a0), fast_exp2(a1));
    }
}

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_spatial_ts_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                       const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles; reused as the O staging tiles of the epilogue
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRingTS * kTileBytes;         // [128][64] fp16 ones: the extra N-atom of the PV B operand
    const uint32_t bars = sOnes + kTileBytes;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRingTS]
    const uint32_t bar_empty = bar_full + 8 * kRingTS;           // [kRingTS]
    const uint32_t bar_s_full = bar_empty + 8 * kRingTS;         // [2]  S_t(j) written, O_t idle
    const uint32_t bar_p_full = bar_s_full + 16;                 // [2]  P_t(j) in TMEM, S_t(j) fully read
    const uint32_t bar_q_ready = bar_p_full + 16;                // [2]  Q_t copied into TMEM
    const uint32_t bar_o_done = bar_q_ready + 16;                // [2]  last PV of tile t retired
    const uint32_t tmem_slot = bar_o_done + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));       // generic pointer to `base`

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRingTS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s_full + 8 * t, 1);
            mbar_init(bar_p_full + 8 * t, 128);
            mbar_init(bar_q_ready + 8 * t, 128);
            mbar_init(bar_o_done + 8 * t, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_qkv, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        for (int i = 0; i < 2 * nkv; ++i) {                     // item 2j = K_j, 2j+1 = V_j
            const int st = i % kRingTS;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(i / kRingTS) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int col = ((i & 1) ? 2 * C : C) + head * kD;
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, (i >> 1) * kKV, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);               // B (= [V | ones]) is MN-major, N = 80
        auto issue_qk = [&](int t, int stage) {                              // S_t = Q_t[tmem] K^T[smem]
            const uint64_t dk = make_desc_sw128(sRing + stage * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ts(tmem_base + t * 128, tmem_base + kColQ + t * kColStride + 8 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                umma_commit(bar_s_full + 8 * t);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int stage, bool first, bool last) {      // O_t += P_t[tmem] [V | ones][smem]
            const uint32_t v_addr = sRing + stage * kTileBytes;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk)
                    umma_f16_ts(tmem_base + kColO + t * kColStride, tmem_base + t * 128 + 8 * kk, dv + (uint64_t)kk * (2048 >> 4), idesc_pv,
                                (!first || kk > 0) ? 1u : 0u);
                if (last) umma_commit(bar_o_done + 8 * t);
            }
            __syncwarp();
        };
        mbar_wait(bar_q_ready + 0, 0);
        mbar_wait(bar_q_ready + 8, 0);
        mbar_wait(bar_full + 0, 0);                                           // K_0 in ring slot 0
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        if (elect_one()) umma_commit(bar_empty + 0);
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
            const bool more = (j + 1 < nkv);
            const int ik = 2 * j + 2, iv = 2 * j + 1;
            const int sk = ik % kRingTS, sv = iv % kRingTS;
            mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRingTS) & 1u);                  // V_j
            if (more) mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRingTS) & 1u);        // K_{j+1}
            for (int t = 0; t < 2; ++t) {
                mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);              // P_t(j) in TMEM, every read of S_t(j) done
                tc_fence_after();
                issue_pv(t, sv, j == 0, !more);
                if (more) issue_qk(t, sk);                                    // in issue order behind PV(t, j): may overwrite S_t | P_t
            }
            if (elect_one()) {
                if (more) umma_commit(bar_empty + 8 * sk);
                umma_commit(bar_empty + 8 * sv);
            }
            __syncwarp();
        }
    } else {
        // ================= softmax / correction / epilogue: one thread per query row of tile t =================
        const int t = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_sel, tO = tmem_base + kColO + t * kColStride + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pQ = gen + (sQ - base) + t * kTileBytes + row * 128;      // this row of the Q tile (SW128), later of the O staging tile
        // ---- Q row: TMA tile -> registers -> TMEM (A operand of every QK^T of this tile)
        mbar_wait(bar_q, 0);
        {
            uint32_t r[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const uint4 u = *reinterpret_cast<const uint4*>(pQ + (((uint32_t)ch ^ sw) << 4));
                r[4 * ch] = u.x; r[4 * ch + 1] = u.y; r[4 * ch + 2] = u.z; r[4 * ch + 3] = u.w;
            }
            const uint32_t tQ = tmem_base + kColQ + t * kColStride + lane_sel;
            tmem_st_32x32b_x16(tQ, r);
            tmem_st_32x32b_x16(tQ + 16, r + 16);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_ready + 8 * t);
        }
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int valid = a.S - j * kKV;                       // keys beyond S are zero-filled by TMA: mask them
            const bool masked = valid < kKV;
            mbar_wait(bar_s_full + 8 * t, (uint32_t)j & 1u);       // S_t(j) ready; PV(t, j-1) retired (issue order): O_t is ours
            tc_fence_after();
            // pass 1: row max.  The TMEM load of chunk ch+1 is in flight while chunk ch is reduced.
            float mx = -INFINITY;
            if (!masked) {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(tS, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 32, rb);
                mx = row_max_chunk(ra, mx);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 64, ra);
                mx = row_max_chunk(rb, mx);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 96, rb);
                mx = row_max_chunk(ra, mx);
                tmem_ld_wait();
                mx = row_max_chunk(rb, mx);
            } else {                                             // ragged last KV tile (S % 128 != 0): simple rolled loop
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tS + ch * 32, r);
                    tmem_ld_wait();
                    mask_chunk(r, ch * 32, valid);
                    mx = row_max_chunk(r, mx);
                }
            }
            float alpha = 1.f;
            const bool grow = (mx - m_ref) * c > kRescaleThreshold;
            const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
            if (any_grow) {
                const float m_new = fmaxf(m_ref, mx);
                alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                m_ref = m_new;
                if (j > 0) {
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {      // 64 output columns + the row-sum columns
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                }
            }
            // pass 2: p = 2^((s - m_ref) * c), packed fp16, written over the first half of this row's own scores
            const float mc = m_ref * c;
            if (!masked) {
                uint32_t ra[32], rb[32], p[16];
                tmem_ld_32x32b_x32(tS, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + 32, rb);
                exp_pack<POLY>(ra, c, mc, p);
                tmem_ld_wait();
                tmem_st_32x32b_x16(tS, p);                       // P keys 0..31 -> columns [0,16): inside score chunk 0 (read)
                tmem_ld_32x32b_x32(tS + 64, ra);
                exp_pack<POLY>(rb, c, mc, p);
                tmem_ld_wait();
                tmem_st_32x32b_x16(tS + 16, p);                  // keys 32..63 -> [16,32)
                tmem_ld_32x32b_x32(tS + 96, rb);
                exp_pack<POLY>(ra, c, mc, p);
                tmem_ld_wait();
                tmem_st_32x32b_x16(tS + 32, p);                  // keys 64..95 -> [32,48): score chunk 1 was consumed two steps ago
                exp_pack<POLY>(rb, c, mc, p);
                tmem_st_32x32b_x16(tS + 48, p);                  // keys 96..127 -> [48,64)
            } else {
                uint32_t pk[64];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {                 // every score chunk is read before any P word is stored
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tS + ch * 32, r);
                    tmem_ld_wait();
                    mask_chunk(r, ch * 32, valid);
                    exp_pack<0>(r, c, mc, pk + 16 * ch);
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) tmem_st_32x32b_x16(tS + 16 * ch, pk + 16 * ch);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * t);
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128, this tile's Q buffer) -> TMA store
        mbar_wait(bar_o_done + 8 * t, 0);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);                      // row sum l = (P @ ones), fp32 from the tensor core
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + ch * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)(ch * 2 + v8);
                *reinterpret_cast<uint4*>(pQ + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (qd == 2 && lane == 0) {                              // first warp of each softmax group (warps 2 and 6)
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sQ + t * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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


// =====================================================================================================================
// TS-mode, TWO threads per query row (18 warps): the softmax of the 8-warp kernels is latency-bound -- ncu: every softmax
// warp is busy 87 % of the time at 0.28 IPC (~790 dependent-ish instructions per KV tile), MUFU 55 %, issue slots 55 %,
// with only two softmax warps per SM sub-partition to hide each other's latencies.  Here every row of tile t is owned by a
// PRIMARY thread (warps 2-5 / 10-13) and a HELPER thread (warps 6-9 / 14-17) on the same TMEM lane:
//   primary: pass 1 over all 128 scores (row max), reference-max update + lazy O rescale, publishes m_ref * c through
//            shared memory, `bar.arrive`; then exponentials of keys 0..63 -> P columns [0,32)
//   helper:  `bar.sync` on the same named barrier (so every primary has finished READING the scores of this KV tile),
//            reads m_ref * c, exponentials of keys 64..127 -> P columns [64,96) (aliased inside its own half of S_t)
// Four softmax warps per sub-partition, ~370 / ~300 instructions per KV tile each.  P_t is therefore two 32-column pieces
// (k-steps 0..3 at column 8 kk, k-steps 4..7 at column 64 + 8 (kk - 4)); everything else is as in attn_spatial_ts_kernel.
constexpr int kAttnThreads2 = 18 * 32;
constexpr int kAttnSmemTS2 = kAttnSmemTS + 1024;                            // + m_ref * c hand-off [2 tiles][128 rows] fp32

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads2, 1)
attn_spatial_ts2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                        const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles; reused as the O staging tiles of the epilogue
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRingTS * kTileBytes;
    const uint32_t sMref = sOnes + kTileBytes;                   // float [2][128]
    const uint32_t bars = sMref + 1024;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRingTS]
    const uint32_t bar_empty = bar_full + 8 * kRingTS;           // [kRingTS]
    const uint32_t bar_s_full = bar_empty + 8 * kRingTS;         // [2]
    const uint32_t bar_p_full = bar_s_full + 16;                 // [2]  256 arrivals: primary + helper threads
    const uint32_t bar_q_ready = bar_p_full + 16;                // [2]  256 arrivals
    const uint32_t bar_o_done = bar_q_ready + 16;                // [2]
    const uint32_t tmem_slot = bar_o_done + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRingTS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s_full + 8 * t, 1);
            mbar_init(bar_p_full + 8 * t, 256);
            mbar_init(bar_q_ready + 8 * t, 256);
            mbar_init(bar_o_done + 8 * t, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads2)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_qkv, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        for (int i = 0; i < 2 * nkv; ++i) {                     // item 2j = K_j, 2j+1 = V_j
            const int st = i % kRingTS;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(i / kRingTS) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int col = ((i & 1) ? 2 * C : C) + head * kD;
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, (i >> 1) * kKV, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);               // B (= [V | ones]) is MN-major, N = 80
        auto issue_qk = [&](int t, int stage) {                              // S_t = Q_t[tmem] K^T[smem]
            const uint64_t dk = make_desc_sw128(sRing + stage * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ts(tmem_base + t * 128, tmem_base + kColQ + t * kColStride + 8 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                umma_commit(bar_s_full + 8 * t);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int stage, bool first, bool last) {      // O_t += P_t[tmem] [V | ones][smem]
            const uint32_t v_addr = sRing + stage * kTileBytes;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk)           // keys 0..63 in P columns [0,32), keys 64..127 in [64,96)
                    umma_f16_ts(tmem_base + kColO + t * kColStride, tmem_base + t * 128 + (kk >> 2) * 64 + 8 * (kk & 3),
                                dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
                if (last) umma_commit(bar_o_done + 8 * t);
            }
            __syncwarp();
        };
        mbar_wait(bar_q_ready + 0, 0);
        mbar_wait(bar_q_ready + 8, 0);
        mbar_wait(bar_full + 0, 0);                                           // K_0 in ring slot 0
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        if (elect_one()) umma_commit(bar_empty + 0);
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
            const bool more = (j + 1 < nkv);
            const int ik = 2 * j + 2, iv = 2 * j + 1;
            const int sk = ik % kRingTS, sv = iv % kRingTS;
            mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRingTS) & 1u);                  // V_j
            if (more) mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRingTS) & 1u);        // K_{j+1}
            for (int t = 0; t < 2; ++t) {
                mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);
                trace_stamp(a, j, 0 + 2 * t);                     // slots 0 / 2: P_t(j) seen by the MMA warp
                tc_fence_after();
                issue_pv(t, sv, j == 0, !more);
                if (more) issue_qk(t, sk);
                trace_stamp(a, j, 1 + 2 * t);                     // slots 1 / 3: PV(t, j) + QK(t, j+1) issued
            }
            if (elect_one()) {
                if (more) umma_commit(bar_empty + 8 * sk);
                umma_commit(bar_empty + 8 * sv);
            }
            __syncwarp();
        }
    } else {
        // ================= softmax: tile t, primary (hf = 0) / helper (hf = 1) thread of one query row =================
        const int sidx = warp - 2;
        const int t = sidx >> 3;
        const int hf = (sidx >> 2) & 1;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_sel, tO = tmem_base + kColO + t * kColStride + lane_sel;
        const uint32_t tSh = tS + hf * 64;                                   // this thread's 64 scores; its P words go to [tSh, tSh + 32)
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pQ = gen + (sQ - base) + t * kTileBytes + row * 128;
        volatile float* mref = reinterpret_cast<volatile float*>(gen + (sMref - base)) + t * 128 + row;
        // named barrier 1 + t: the 128 primaries arrive, the 128 helpers sync
        // ---- Q row: TMA tile -> registers -> TMEM; each of the two threads copies half of the row (32 halves = 16 columns)
        mbar_wait(bar_q, 0);
        {
            uint32_t r[16];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const uint4 u = *reinterpret_cast<const uint4*>(pQ + (((uint32_t)(hf * 4 + ch) ^ sw) << 4));
                r[4 * ch] = u.x; r[4 * ch + 1] = u.y; r[4 * ch + 2] = u.z; r[4 * ch + 3] = u.w;
            }
            tmem_st_32x32b_x16(tmem_base + kColQ + t * kColStride + hf * 16 + lane_sel, r);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_ready + 8 * t);
        }
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int valid = a.S - j * kKV;
            const bool masked = valid < kKV;
            float mc;
            if (hf == 0) {
                mbar_wait(bar_s_full + 8 * t, (uint32_t)j & 1u);   // S_t(j) ready; PV(t, j-1) retired: O_t is ours
                if (qd == 2) trace_stamp(a, j, 4 + 4 * t);        // slots 4 / 8: S_t(j) seen by the primary warp of quadrant 2
                tc_fence_after();
                float mx = -INFINITY;
                if (!masked) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(tS, ra);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 32, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 64, ra);
                    mx = row_max_chunk(rb, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 96, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    mx = row_max_chunk(rb, mx);
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tS + ch * 32, r);
                        tmem_ld_wait();
                        mask_chunk(r, ch * 32, valid);
                        mx = row_max_chunk(r, mx);
                    }
                }
                const bool grow = (mx - m_ref) * c > kRescaleThreshold;
                const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
                float alpha = 1.f;
                if (any_grow) {
                    const float m_new = fmaxf(m_ref, mx);
                    alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                    m_ref = m_new;
                }
                mc = m_ref * c;
                *mref = mc;                                      // hand-off to the helper thread of this row
                tc_fence_before();                               // the helper's tcgen05 accesses are ordered after what this thread observed
                if (t == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
                else        asm volatile("bar.arrive 2, 256;" ::: "memory");
                if (qd == 2) trace_stamp(a, j, 5 + 4 * t);        // slots 5 / 9: pass 1 done, reference max published
                if (any_grow && j > 0) {
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                }
            } else {
                // every primary of this tile has read the scores of KV tile j (and published its reference max)
                if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                else        asm volatile("bar.sync 2, 256;" ::: "memory");
                tc_fence_after();
                mc = *mref;
                if (qd == 2) trace_stamp(a, j, 6 + 4 * t);        // slots 6 / 10: helper released
            }
            // exponentials of this thread's 64 keys -> 32 packed words over the first half of its own score columns
            {
                uint32_t ra[32], rb[32], p[16];
                tmem_ld_32x32b_x32(tSh, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tSh + 32, rb);
                if (masked) mask_chunk(ra, hf * 64, valid);
                exp_pack<POLY>(ra, c, mc, p);
                tmem_ld_wait();
                tmem_st_32x32b_x16(tSh, p);
                if (masked) mask_chunk(rb, hf * 64 + 32, valid);
                exp_pack<POLY>(rb, c, mc, p);
                tmem_st_32x32b_x16(tSh + 16, p);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * t);
            if (qd == 2) trace_stamp(a, j, hf ? 12 + t : 7 + 4 * t);   // slots 7 / 11: primary's P written; 12 / 13: helper's
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128, this tile's Q buffer) -> TMA store; each thread converts 32 columns
        mbar_wait(bar_o_done + 8 * t, 0);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + (hf * 2 + ch) * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)((hf * 2 + ch) * 2 + v8);
                *reinterpret_cast<uint4*>(pQ + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
        else        asm volatile("bar.sync 4, 256;" ::: "memory");
        if (hf == 0 && qd == 2 && lane == 0) {                    // warps 2 and 10
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sQ + t * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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


// =====================================================================================================================
// TS-mode, two threads per row, 64-key KV steps with DOUBLE-BUFFERED scores.
//
// The pipeline timeline of attn_spatial_ts2_kernel (profiles/r02_attn_trace_ts2.txt, cycles per 128-key tile of one Q tile):
//   S seen -> pass 1 done 743 -> P stored 2001 -> MMA warp sees P 2077 -> PV + next QK issued 2709 (the issuing thread is throttled
//   to the tensor pipe's execution rate) -> next S seen 3067
// i.e. softmax (1965) and the MMA phase (1066: PV then the next QK^T, plus two barrier hops) of ONE Q tile are strictly serial, the
// two Q tiles only fill each other's MMA phase, and MUFU sits at 55 %.  Here the scores of a Q tile live in two 64-column TMEM
// buffers: while the softmax threads work on S_t(j) in buffer j & 1, QK^T(t, j+1) has already been issued into the other buffer
// (right behind PV(t, j-1), which vacated it -- in-order tensor pipe), so a finished softmax step finds its next scores ready and
// both Q tiles keep the exp / issue pipes busy at the same time.
//   MMA warp, step j, tile t:  wait P_t(j) -> PV(t, j) -> QK^T(t, j+2) into buffer j & 1 -> commit "S index j+2" (real or virtual)
//   "S index i ready" also means PV(t, i-2) has retired, which is what the lazy O rescale of step i-1 and the epilogue wait for.
// TMEM columns: S_t buffers at t*128 + {0, 64} (P aliased: primary keys 0..31 -> [0,16), helper keys 32..63 -> [32,48));
//               O_t at 256 + t*128 (80 columns);  Q_t at 352 + t*128 (32 columns).
constexpr int kKV3 = 64;
constexpr int kTileKV3 = kKV3 * kD * 2;                                     // 8 KB K / V tile
constexpr int kRing3 = 12;
constexpr int kAttnSmemTS3 = 1024 + 2 * kTileBytes + kRing3 * kTileKV3 + kTileKV3 + 1024 + 512;   // Q0,Q1 | ring | ones | m_ref | barriers

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads2, 1)
attn_spatial_ts3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles; reused as the O staging tiles of the epilogue
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRing3 * kTileKV3;            // [64][64] fp16 ones: the extra N-atom of the PV B operand
    const uint32_t sMref = sOnes + kTileKV3;                     // float [2][128]
    const uint32_t bars = sMref + 1024;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing3]
    const uint32_t bar_empty = bar_full + 8 * kRing3;            // [kRing3]
    const uint32_t bar_s_full = bar_empty + 8 * kRing3;          // [tile][buffer]
    const uint32_t bar_p_full = bar_s_full + 32;                 // [tile][buffer], 256 arrivals
    const uint32_t bar_q_ready = bar_p_full + 32;                // [2], 256 arrivals
    const uint32_t tmem_slot = bar_q_ready + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;                                       // 64-key steps

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing3; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_p_full + 8 * i, 256); }
        for (int t = 0; t < 2; ++t) mbar_init(bar_q_ready + 8 * t, 256);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileKV3 / 16; i += kAttnThreads2)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer: K_0, K_1, then per step V_j, K_{j+2} ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_q, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_q, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        int cnt = 0;
        auto load = [&](int col, int key0) {
            const int st = cnt % kRing3;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(cnt / kRing3) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileKV3);
                tma_load_3d(sRing + st * kTileKV3, &tm_kv, bar_full + 8 * st, col + head * kD, key0, img);
            }
            __syncwarp();
            ++cnt;
        };
        load(C, 0);
        if (nkv > 1) load(C, kKV3);
        for (int j = 0; j < nkv; ++j) {
            load(2 * C, j * kKV3);
            if (j + 2 < nkv) load(C, (j + 2) * kKV3);
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t idesc_qk = make_idesc_f16(kKV3, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);               // B (= [V | ones]) is MN-major, N = 80
        auto issue_qk = [&](int t, int b, int stage) {                       // S_t[b] = Q_t[tmem] K^T[smem]
            const uint64_t dk = make_desc_sw128(sRing + stage * kTileKV3);
            if (!(a.dbg & 2) && elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ts(tmem_base + t * 128 + b * 64, tmem_base + kColQ + t * kColStride + 8 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int b, int stage, bool first) {           // O_t += P_t[b][tmem] [V | ones][smem]
            const uint32_t v_addr = sRing + stage * kTileKV3;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            if (!(a.dbg & 2) && elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV3 / 16; ++kk)          // keys 0..31 in P columns [0,16), keys 32..63 in [32,48)
                    umma_f16_ts(tmem_base + kColO + t * kColStride, tmem_base + t * 128 + b * 64 + (kk >> 1) * 32 + 8 * (kk & 1),
                                dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
            }
            __syncwarp();
        };
        auto commit = [&](uint32_t bar) {
            if (elect_one()) umma_commit(bar);
            __syncwarp();
        };
        int cnt = 0;
        auto next_slot = [&]() {                                             // waits for the next ring item, returns its slot
            const int st = cnt % kRing3;
            mbar_wait(bar_full + 8 * st, (uint32_t)(cnt / kRing3) & 1u);
            ++cnt;
            return st;
        };
        mbar_wait(bar_q_ready + 0, 0);
        mbar_wait(bar_q_ready + 8, 0);
        for (int b = 0; b < 2 && b < nkv; ++b) {                             // prologue: S_t(0) -> buffer 0, S_t(1) -> buffer 1
            const int sk = next_slot();
            tc_fence_after();
            for (int t = 0; t < 2; ++t) {
                issue_qk(t, b, sk);
                commit(bar_s_full + 8 * (2 * t + b));
            }
            commit(bar_empty + 8 * sk);
        }
        for (int j = 0; j < nkv; ++j) {
            const int b = j & 1;
            const bool more2 = (j + 2 < nkv);
            const int sv = next_slot();                                      // V_j
            const int sk = more2 ? next_slot() : 0;                          // K_{j+2}
            for (int t = 0; t < 2; ++t) {
                mbar_wait(bar_p_full + 8 * (2 * t + b), (uint32_t)(j >> 1) & 1u);   // P_t(j) in TMEM, every read of S_t(j) done
                trace_stamp(a, j >> 1, (j & 1) * 8 + 0 + 2 * t);
                tc_fence_after();
                issue_pv(t, b, sv, j == 0);
                if (more2) issue_qk(t, b, sk);                               // behind PV(t, j) in issue order: overwrites S_t[b] | P_t[b]
                commit(bar_s_full + 8 * (2 * t + b));                        // "S index j+2" (virtual past the end): PV(t, j) retired
                trace_stamp(a, j >> 1, (j & 1) * 8 + 1 + 2 * t);
            }
            commit(bar_empty + 8 * sv);
            if (more2) commit(bar_empty + 8 * sk);
        }
    } else {
        // ================= softmax: tile t, primary (hf = 0) / helper (hf = 1) thread of one query row =================
        const int sidx = warp - 2;
        const int t = sidx >> 3;
        const int hf = (sidx >> 2) & 1;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS0 = tmem_base + t * 128 + lane_sel, tO = tmem_base + kColO + t * kColStride + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pQ = gen + (sQ - base) + t * kTileBytes + row * 128;
        volatile float* mref = reinterpret_cast<volatile float*>(gen + (sMref - base)) + t * 128 + row;
        // ---- Q row: TMA tile -> registers -> TMEM; each of the two threads copies half of the row
        mbar_wait(bar_q, 0);
        {
            uint32_t r[16];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const uint4 u = *reinterpret_cast<const uint4*>(pQ + (((uint32_t)(hf * 4 + ch) ^ sw) << 4));
                r[4 * ch] = u.x; r[4 * ch + 1] = u.y; r[4 * ch + 2] = u.z; r[4 * ch + 3] = u.w;
            }
            tmem_st_32x32b_x16(tmem_base + kColQ + t * kColStride + hf * 16 + lane_sel, r);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_ready + 8 * t);
        }
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int b = j & 1;
            const int valid = a.S - j * kKV3;
            const bool masked = valid < kKV3;
            const uint32_t tS = tS0 + b * 64;
            float mc;
            if (hf == 0) {
                mbar_wait(bar_s_full + 8 * (2 * t + b), (uint32_t)(j >> 1) & 1u);     // S index j
                if (qd == 2) trace_stamp(a, j >> 1, (j & 1) * 8 + 4 + 2 * t);
                tc_fence_after();
                float mx = -INFINITY;
                if (a.dbg & 1) mx = 1.f;
                else {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(tS, ra);
                    tmem_ld_32x32b_x32(tS + 32, rb);
                    tmem_ld_wait();
                    if (masked) { mask_chunk(ra, 0, valid); mask_chunk(rb, 32, valid); }
                    mx = row_max_chunk(ra, mx);
                    mx = row_max_chunk(rb, mx);
                }
                const bool grow = (mx - m_ref) * c > kRescaleThreshold;
                const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
                float alpha = 1.f;
                if (any_grow) {
                    const float m_new = fmaxf(m_ref, mx);
                    alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                    m_ref = m_new;
                }
                mc = m_ref * c;
                *mref = mc;                                      // hand-off to the helper thread of this row
                tc_fence_before();
                if (t == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
                else        asm volatile("bar.arrive 2, 256;" ::: "memory");
                if (any_grow && j > 0) {
                    // O_t must be idle: "S index j+1 ready" was committed behind PV(t, j-1)
                    mbar_wait(bar_s_full + 8 * (2 * t + (b ^ 1)), (uint32_t)((j + 1) >> 1) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                }
            } else {
                if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                else        asm volatile("bar.sync 2, 256;" ::: "memory");
                tc_fence_after();
                mc = *mref;
            }
            // exponentials of this thread's 32 keys -> 16 packed words at the start of its own 32 score columns
            if (!(a.dbg & 1)) {
                uint32_t r[32], p[16];
                tmem_ld_32x32b_x32(tS + hf * 32, r);
                tmem_ld_wait();
                if (masked) mask_chunk(r, hf * 32, valid);
                exp_pack<POLY>(r, c, mc, p);
                tmem_st_32x32b_x16(tS + hf * 32, p);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * (2 * t + b));
            if (qd == 2 && hf == 0) trace_stamp(a, j >> 1, (j & 1) * 8 + 5 + 2 * t);
        }
        // ---- epilogue: "S index nkv+1" == PV(t, nkv-1) retired
        mbar_wait(bar_s_full + 8 * (2 * t + ((nkv + 1) & 1)), (uint32_t)((nkv + 1) >> 1) & 1u);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + (hf * 2 + ch) * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)((hf * 2 + ch) * 2 + v8);
                *reinterpret_cast<uint4*>(pQ + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
        else        asm volatile("bar.sync 4, 256;" ::: "memory");
        if (hf == 0 && qd == 2 && lane == 0) {                    // warps 2 and 10
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sQ + t * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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


// =====================================================================================================================
// ts4: ONE 128-row Q tile per CTA, FOUR threads per query row, 128-key steps with double-buffered scores, single-pass softmax.
//
// What the timing switches of ts3 showed (profiles/r02_attn_ts3_bounds.txt, cycles per 128 keys of two Q tiles): full kernel 3 900,
// MMAs only 2 990, softmax only 3 090, NEITHER (barriers, commits and TMA only) 2 170 -- the synchronisation skeleton itself is the
// bound: every tcgen05.commit costs the issuing warp ~200 cycles and every (already complete) mbarrier try_wait ~90, and the one MMA
// warp served two Q tiles x two 64-key steps with 8 commits + 8 waits per 128 keys.  This variant keeps the ideas that are sound
// (A operands in TMEM, P aliased on S, next QK^T issued behind the current PV into the other score buffer) and minimises the
// hand-offs:
//   * one Q tile, so the MMA warp does ONE commit and three waits per 128-key step: wait V_j -> wait P(j) -> PV(j) -> wait K_{j+2}
//     -> QK^T(j+2) -> commit "S index j+2".  K/V ring slots are released by a softmax thread when it observes "S index j+2"
//     (that commit sits behind PV(j) and QK^T(j+2) in issue order), not by extra commits;
//   * sixteen softmax warps on the same 128 rows: thread (role q, row r) owns keys 32q..32q+31 of row r.  It loads its 32 scores ONCE,
//     publishes its partial row max through shared memory, one 512-thread named barrier, reads the four partials -> the row max
//     and (identically in all four threads of the row) the reference-max / lazy-rescale decision, exponentiates its 32 scores from
//     registers and stores 16 packed words over the start of its own score columns.  No second TMEM sweep, no primary / helper
//     asymmetry, 4 warps per SM sub-partition.
// TMEM columns: S buffers [0,128) and [128,256) (P of role q at buffer + 32q .. +16), O [256,336), Q [352,384).
constexpr int kRing4 = 8;
constexpr int kAttnThreads4 = 18 * 32;
constexpr int kAttnSmemTS4 = 1024 + kTileBytes * (1 + kRing4 + 1) + 4096 + 512;      // Q | K/V ring | ones | partial maxima [2][4][128] | barriers

__device__ __forceinline__ int ring_pos_v(int j, int nkv) { return min(nkv, 2) + j + min(j, max(nkv - 2, 0)); }
__device__ __forceinline__ int ring_pos_k(int i, int nkv) { return i < 2 ? i : ring_pos_v(i - 2, nkv) + 1; }

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads4, 1)
attn_spatial_ts4_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                        const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // Q tile; reused as the O staging tile of the epilogue
    const uint32_t sRing = sQ + kTileBytes;
    const uint32_t sOnes = sRing + kRing4 * kTileBytes;
    const uint32_t sMax = sOnes + kTileBytes;                    // float [2 (step parity)][4 roles][128 rows]
    const uint32_t bars = sMax + 4096;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing4]
    const uint32_t bar_empty = bar_full + 8 * kRing4;            // [kRing4]  one arrival (a softmax thread)
    const uint32_t bar_s_full = bar_empty + 8 * kRing4;          // [2 buffers]
    const uint32_t bar_p_full = bar_s_full + 16;                 // [2 buffers], 512 arrivals
    const uint32_t bar_q_ready = bar_p_full + 16;                // 256 arrivals (roles 0 and 1 copy Q)
    const uint32_t tmem_slot = bar_q_ready + 8;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;
    constexpr uint32_t kColO4 = 256, kColQ4 = 352;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing4; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_s_full + 8 * b, 1); mbar_init(bar_p_full + 8 * b, 512); }
        mbar_init(bar_q_ready, 256);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads4)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer: Q, K_0, K_1, then per step V_j, K_{j+2} ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
        }
        __syncwarp();
        int cnt = 0;
        auto load = [&](int col, int key0) {
            const int st = cnt % kRing4;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(cnt / kRing4) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col + head * kD, key0, img);
            }
            __syncwarp();
            ++cnt;
        };
        load(C, 0);
        if (nkv > 1) load(C, kKV);
        for (int j = 0; j < nkv; ++j) {
            load(2 * C, j * kKV);
            if (j + 2 < nkv) load(C, (j + 2) * kKV);
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);               // B (= [V | ones]) is MN-major, N = 80
        auto issue_qk = [&](int b, int stage) {                              // S[b] = Q[tmem] K^T[smem]
            const uint64_t dk = make_desc_sw128(sRing + stage * kTileBytes);
            if (!(a.dbg & 2) && elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ts(tmem_base + b * 128, tmem_base + kColQ4 + 8 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int b, int stage, bool first) {                  // O += P[b][tmem] [V | ones][smem]
            const uint32_t v_addr = sRing + stage * kTileBytes;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            if (!(a.dbg & 2) && elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk)           // keys 32q..32q+31 (role q) in P columns [32q, 32q+16)
                    umma_f16_ts(tmem_base + kColO4, tmem_base + b * 128 + (kk >> 1) * 32 + 8 * (kk & 1),
                                dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
            }
            __syncwarp();
        };
        auto commit = [&](uint32_t bar) {
            if (elect_one()) umma_commit(bar);
            __syncwarp();
        };
        int cnt = 0;
        auto next_slot = [&]() {
            const int st = cnt % kRing4;
            mbar_wait(bar_full + 8 * st, (uint32_t)(cnt / kRing4) & 1u);
            ++cnt;
            return st;
        };
        mbar_wait(bar_q_ready, 0);
        for (int b = 0; b < 2 && b < nkv; ++b) {                             // prologue: S(0) -> buffer 0, S(1) -> buffer 1
            const int sk = next_slot();
            tc_fence_after();
            issue_qk(b, sk);
            commit(bar_s_full + 8 * b);
        }
        for (int j = 0; j < nkv; ++j) {
            const int b = j & 1;
            const bool more2 = (j + 2 < nkv);
            const int sv = next_slot();                                      // V_j
            mbar_wait(bar_p_full + 8 * b, (uint32_t)(j >> 1) & 1u);          // P(j) in TMEM, every read of S(j) done
            trace_stamp(a, j, 0);
            tc_fence_after();
            issue_pv(b, sv, j == 0);
            if (more2) {
                const int sk = next_slot();                                  // K_{j+2}
                tc_fence_after();
                issue_qk(b, sk);                                             // behind PV(j) in issue order: overwrites S[b] | P[b]
            }
            commit(bar_s_full + 8 * b);                                      // "S index j+2" (virtual past the end): PV(j) retired
            trace_stamp(a, j, 1);
        }
    } else {
        // ================= softmax: thread (role q, row) owns keys 32q..32q+31 of its row in every step =================
        const int role = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS0 = tmem_base + lane_sel + role * 32, tO = tmem_base + kColO4 + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pQ = gen + (sQ - base) + row * 128;
        float* smax = reinterpret_cast<float*>(gen + (sMax - base));
        const bool releaser = (warp == 2);                       // its elected lane hands ring slots back to the TMA producer
        // ---- Q row: TMA tile -> registers -> TMEM (roles 0 and 1 copy half a row each)
        if (role < 2) {
            mbar_wait(bar_q, 0);
            uint32_t r[16];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const uint4 u = *reinterpret_cast<const uint4*>(pQ + (((uint32_t)(role * 4 + ch) ^ sw) << 4));
                r[4 * ch] = u.x; r[4 * ch + 1] = u.y; r[4 * ch + 2] = u.z; r[4 * ch + 3] = u.w;
            }
            tmem_st_32x32b_x16(tmem_base + kColQ4 + role * 16 + lane_sel, r);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_ready);
        }
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int b = j & 1;
            const int valid = a.S - j * kKV;
            const uint32_t tS = tS0 + b * 128;
            mbar_wait(bar_s_full + 8 * b, (uint32_t)(j >> 1) & 1u);          // S index j: also PV(j-2) and QK^T(j) have retired
            if (releaser) {
                if (elect_one()) {
                    mbar_arrive(bar_empty + 8 * (ring_pos_k(j, nkv) % kRing4));                      // K_j
                    if (j >= 2) mbar_arrive(bar_empty + 8 * (ring_pos_v(j - 2, nkv) % kRing4));      // V_{j-2}
                }
                __syncwarp();
                trace_stamp(a, j, 2);
            }
            tc_fence_after();
            uint32_t r[32];
            float mx;
            if (a.dbg & 1) {
                mx = 1.f;
            } else {
                tmem_ld_32x32b_x32(tS, r);
                tmem_ld_wait();
                if (valid < kKV) mask_chunk(r, role * 32, valid);
                mx = row_max_chunk(r, -INFINITY);
            }
            // partial row maxima through shared memory (double-buffered by step parity), one barrier for the 16 softmax warps
            float* sm = smax + b * 512;
            sm[role * 128 + row] = mx;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            mx = fmaxf(fmaxf(sm[row], sm[128 + row]), fmaxf(sm[256 + row], sm[384 + row]));
            if (releaser) trace_stamp(a, j, 3);
            const bool grow = (mx - m_ref) * c > kRescaleThreshold;
            const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);       // same rows, same maxima: identical in all 4 roles
            if (any_grow) {
                const float m_new = fmaxf(m_ref, mx);
                const float alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                m_ref = m_new;
                if (j > 0) {
                    // O must be idle: "S index j+1" was committed behind PV(j-1).  Role q rescales output columns [16q, 16q+16),
                    // role 0 also the row-sum columns.
                    mbar_wait(bar_s_full + 8 * (b ^ 1), (uint32_t)((j + 1) >> 1) & 1u);
                    tc_fence_after();
                    uint32_t o[16];
                    tmem_ld_32x32b_x16(tO + role * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                    tmem_st_32x32b_x16(tO + role * 16, o);
                    if (role == 0) {
                        tmem_ld_32x32b_x16(tO + 64, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st_32x32b_x16(tO + 64, o);
                    }
                }
            }
            if (!(a.dbg & 1)) {
                uint32_t p[16];
                exp_pack<POLY>(r, c, m_ref * c, p);
                tmem_st_32x32b_x16(tS, p);                       // over the first 16 of this thread's own 32 score columns
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * b);
            if (releaser) trace_stamp(a, j, 4);
        }
        // ---- epilogue: "S index nkv+1" == PV(nkv-1) retired; role q converts output columns [16q, 16q+16)
        mbar_wait(bar_s_full + 8 * ((nkv + 1) & 1), (uint32_t)((nkv + 1) >> 1) & 1u);
        tc_fence_after();
        {
            uint32_t l[16], o[16];
            tmem_ld_32x32b_x16(tO + 64, l);                      // row sum l = (P @ ones), fp32 from the tensor core
            tmem_ld_32x32b_x16(tO + role * 16, o);
            tmem_ld_wait();
            const float inv = 1.f / __uint_as_float(l[0]);
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(o[v8 * 8 + 2 * e]) * inv, __uint_as_float(o[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)(role * 2 + v8);
                *reinterpret_cast<uint4*>(pQ + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 512;" ::: "memory");
        if (warp == 2 && lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sQ), "r"(head * kD), "r"(q0), "r"(img) : "memory");
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


// =====================================================================================================================
// ss2: the shared-memory-operand kernel (attn_spatial_kernel: Q, P in smem, QK^T of the next KV tile issued as soon as the scores
// are read) with TWO threads per query row: the primary thread (warps 2-5 / 10-13) runs pass 1 over all 128 scores, decides the
// reference max / lazy rescale and exponentiates keys 0..63; the helper thread (warps 6-9 / 14-17) is released through a named
// barrier and exponentiates keys 64..127 into the second half of the P tile.  Timeline of the TS variants
// (profiles/r02_attn_trace_ts2.txt): the exponential phase of a tile shrinks from ~1 800 to ~1 220 cycles with the split, but
// aliasing P on S delays the next QK^T behind PV (+600 cycles); here P stays in shared memory so both gains combine.
constexpr int kAttnSmemSS2 = kAttnSmem + 1024;

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads2, 1)
attn_spatial_ss2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                        const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRing * kTileBytes;
    const uint32_t sP = sOnes + kTileBytes;                      // tile t: 2 x 16 KB at sP + t*32 KB; reused for O staging
    const uint32_t sMref = sP + 4 * kTileBytes;                  // float [2][128]
    const uint32_t bars = sMref + 1024;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing]
    const uint32_t bar_empty = bar_full + 8 * kRing;             // [kRing]
    const uint32_t bar_s_full = bar_empty + 8 * kRing;           // [2]
    const uint32_t bar_s_free = bar_s_full + 16;                 // [2]  256 arrivals
    const uint32_t bar_p_full = bar_s_free + 16;                 // [2]  256 arrivals
    const uint32_t bar_pv_done = bar_p_full + 16;                // [2]
    const uint32_t tmem_slot = bar_pv_done + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s_full + 8 * t, 1);
            mbar_init(bar_s_free + 8 * t, 256);
            mbar_init(bar_p_full + 8 * t, 256);
            mbar_init(bar_pv_done + 8 * t, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads2)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_qkv, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        for (int i = 0; i < 2 * nkv; ++i) {                     // item 2j = K_j, 2j+1 = V_j
            const int st = i % kRing;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(i / kRing) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int col = ((i & 1) ? 2 * C : C) + head * kD;
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, (i >> 1) * kKV, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (as attn_spatial_kernel) ================================
        const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);
        auto issue_qk = [&](int t, int stage) {
            const uint64_t dq = make_desc_sw128(sQ + t * kTileBytes), dk = make_desc_sw128(sRing + stage * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ss(tmem_base + t * 128, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                umma_commit(bar_s_full + 8 * t);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int stage, bool first) {
            const uint32_t v_addr = sRing + stage * kTileBytes;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            const uint64_t dp0 = make_desc_sw128(sP + t * 2 * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk) {
                    const uint64_t dp = dp0 + (uint64_t)((kk >> 2) * (kTileBytes >> 4) + 2 * (kk & 3));
                    umma_f16_ss(tmem_base + 256 + t * 128, dp, dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
                }
                umma_commit(bar_pv_done + 8 * t);
            }
            __syncwarp();
        };
        mbar_wait(bar_q, 0);
        mbar_wait(bar_full + 0, 0);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        if (elect_one()) umma_commit(bar_empty + 0);
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
            const bool more = (j + 1 < nkv);
            const int ik = 2 * j + 2, iv = 2 * j + 1;
            const int sk = ik % kRing, sv = iv % kRing;
            if (more) mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRing) & 1u);
            mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRing) & 1u);
            for (int t = 0; t < 2; ++t) {
                if (more) {
                    mbar_wait(bar_s_free + 8 * t, (uint32_t)j & 1u);      // both threads of every row have read S_t(j)
                    tc_fence_after();
                    issue_qk(t, sk);
                }
                mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);
                tc_fence_after();
                issue_pv(t, sv, j == 0);
            }
            if (elect_one()) {
                if (more) umma_commit(bar_empty + 8 * sk);
                umma_commit(bar_empty + 8 * sv);
            }
            __syncwarp();
        }
    } else {
        // ================= softmax: tile t, primary (hf = 0) / helper (hf = 1) thread of one query row =================
        const int sidx = warp - 2;
        const int t = sidx >> 3;
        const int hf = (sidx >> 2) & 1;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_sel, tO = tmem_base + 256 + t * 128 + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pP = gen + (sP - base) + t * 2 * kTileBytes + row * 128;
        volatile float* mref = reinterpret_cast<volatile float*>(gen + (sMref - base)) + t * 128 + row;
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int valid = a.S - j * kKV;
            const bool masked = valid < kKV;
            float mc, alpha = 1.f;
            bool rescale = false;
            if (hf == 0) {
                mbar_wait(bar_s_full + 8 * t, (uint32_t)j & 1u);
                tc_fence_after();
                float mx = -INFINITY;
                if (!masked) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(tS, ra);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 32, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 64, ra);
                    mx = row_max_chunk(rb, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 96, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    mx = row_max_chunk(rb, mx);
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tS + ch * 32, r);
                        tmem_ld_wait();
                        mask_chunk(r, ch * 32, valid);
                        mx = row_max_chunk(r, mx);
                    }
                }
                const bool grow = (mx - m_ref) * c > kRescaleThreshold;
                const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
                if (any_grow) {
                    const float m_new = fmaxf(m_ref, mx);
                    alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                    m_ref = m_new;
                    rescale = (j > 0);
                }
                mc = m_ref * c;
                *mref = mc;
                tc_fence_before();
                if (t == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
                else        asm volatile("bar.arrive 2, 256;" ::: "memory");
            } else {
                if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                else        asm volatile("bar.sync 2, 256;" ::: "memory");
                tc_fence_after();
                mc = *mref;
            }
            if (j > 0) {
                // the P_t buffer and O_t belong to the tensor core until PV of tile j-1 has retired
                mbar_wait(bar_pv_done + 8 * t, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (rescale) {
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                    tmem_st_wait();
                }
            }
            // this thread's 64 keys: two 32-score chunks -> packed fp16 -> its half of the P tile (K-major, SWIZZLE_128B)
            {
                uint32_t ra[32], rb[32];
                uint8_t* dst = pP + hf * kTileBytes;
                tmem_ld_32x32b_x32(tS + hf * 64, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + hf * 64 + 32, rb);
                if (masked) mask_chunk(ra, hf * 64, valid);
                exp_chunk<POLY>(ra, 0, c, mc, dst, sw);
                tmem_ld_wait();
                tc_fence_before();                               // last read of S_t by this thread: release it for the next QK
                mbar_arrive(bar_s_free + 8 * t);
                if (masked) mask_chunk(rb, hf * 64 + 32, valid);
                exp_chunk<POLY>(rb, 1, c, mc, dst, sw);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * t);
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128, first half of this tile's P buffer) -> TMA store; 32 columns per thread
        mbar_wait(bar_pv_done + 8 * t, (uint32_t)(nkv - 1) & 1u);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + (hf * 2 + ch) * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)((hf * 2 + ch) * 2 + v8);
                *reinterpret_cast<uint4*>(pP + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
        else        asm volatile("bar.sync 4, 256;" ::: "memory");
        if (hf == 0 && qd == 2 && lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sP + t * 2 * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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

// =====================================================================================================================
// ss3 = ss2 + ONE MMA WARP PER Q TILE + ring slots released by the softmax threads.
// The MMA warp of the kernels above serves both tiles in a fixed order and pays, per KV tile, 6 tcgen05.commit (~200 cycles of
// issue time each, measured on the barrier-only skeletons: profiles/r02_attn_ts3_bounds.txt) and 6 mbarrier waits (~90 each when
// already complete) on top of the throttled issue of 24 MMAs: ~2 900 cycles, the iteration time of EVERY variant.  Here each tile's
// QK^T / PV chain has its own issuing warp (2 commits + 4 waits per KV tile) and the K/V ring is handed back by a softmax thread of
// each tile once it has observed the commit that covers the slot's last reader.
// ss2: the shared-memory-operand kernel (attn_spatial_kernel: Q, P in smem, QK^T of the next KV tile issued as soon as the scores
// are read) with TWO threads per query row: the primary thread (warps 2-5 / 10-13) runs pass 1 over all 128 scores, decides the
// reference max / lazy rescale and exponentiates keys 0..63; the helper thread (warps 6-9 / 14-17) is released through a named
// barrier and exponentiates keys 64..127 into the second half of the P tile.  Timeline of the TS variants
// (profiles/r02_attn_trace_ts2.txt): the exponential phase of a tile shrinks from ~1 800 to ~1 220 cycles with the split, but
// aliasing P on S delays the next QK^T behind PV (+600 cycles); here P stays in shared memory so both gains combine.
constexpr int kAttnThreads3 = 19 * 32;

template <int POLY>
__global__ void __launch_bounds__(kAttnThreads3, 1)
attn_spatial_ss3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                        const __grid_constant__ AttnArgs a, int C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                    // 2 tiles
    const uint32_t sRing = sQ + 2 * kTileBytes;
    const uint32_t sOnes = sRing + kRing * kTileBytes;
    const uint32_t sP = sOnes + kTileBytes;                      // tile t: 2 x 16 KB at sP + t*32 KB; reused for O staging
    const uint32_t sMref = sP + 4 * kTileBytes;                  // float [2][128]
    const uint32_t bars = sMref + 1024;
    const uint32_t bar_q = bars;
    const uint32_t bar_full = bars + 8;                          // [kRing]
    const uint32_t bar_empty = bar_full + 8 * kRing;             // [kRing]
    const uint32_t bar_s_full = bar_empty + 8 * kRing;           // [2]
    const uint32_t bar_s_free = bar_s_full + 16;                 // [2]  256 arrivals
    const uint32_t bar_p_full = bar_s_free + 16;                 // [2]  256 arrivals
    const uint32_t bar_pv_done = bar_p_full + 16;                // [2]
    const uint32_t tmem_slot = bar_pv_done + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kQ, head = blockIdx.y, img = blockIdx.z;
    const int nkv = a.nkv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2); }   // one release per tile
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s_full + 8 * t, 1);
            mbar_init(bar_s_free + 8 * t, 256);
            mbar_init(bar_p_full + 8 * t, 256);
            mbar_init(bar_pv_done + 8 * t, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kTileBytes / 16; i += kAttnThreads3)
        *reinterpret_cast<uint4*>(gen + (sOnes - base) + i * 16) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_3d(sQ, &tm_qkv, bar_q, head * kD, q0, img);
            tma_load_3d(sQ + kTileBytes, &tm_qkv, bar_q, head * kD, q0 + kQ, img);
        }
        __syncwarp();
        for (int i = 0; i < 2 * nkv; ++i) {                     // item 2j = K_j, 2j+1 = V_j
            const int st = i % kRing;
            mbar_wait(bar_empty + 8 * st, ((uint32_t)(i / kRing) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * st, kTileBytes);
                const int col = ((i & 1) ? 2 * C : C) + head * kD;
                tma_load_3d(sRing + st * kTileBytes, &tm_qkv, bar_full + 8 * st, col, (i >> 1) * kKV, img);
            }
            __syncwarp();
        }
    } else if (warp <= 2) {
        // ================================ MMA issuer of tile t = warp - 1 ================================
        const int t = warp - 1;
        const uint32_t idesc_qk = make_idesc_f16(kKV, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(kOCols, 0, 1);
        auto issue_qk = [&](int stage) {
            const uint64_t dq = make_desc_sw128(sQ + t * kTileBytes), dk = make_desc_sw128(sRing + stage * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_f16_ss(tmem_base + t * 128, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
                umma_commit(bar_s_full + 8 * t);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int stage, bool first) {
            const uint32_t v_addr = sRing + stage * kTileBytes;
            const uint64_t dv = (uint64_t((v_addr & 0x3FFFF) >> 4)) | (uint64_t((sOnes - v_addr) >> 4) << 16) |
                                (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
            const uint64_t dp0 = make_desc_sw128(sP + t * 2 * kTileBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk) {
                    const uint64_t dp = dp0 + (uint64_t)((kk >> 2) * (kTileBytes >> 4) + 2 * (kk & 3));
                    umma_f16_ss(tmem_base + 256 + t * 128, dp, dv + (uint64_t)kk * (2048 >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
                }
                umma_commit(bar_pv_done + 8 * t);
            }
            __syncwarp();
        };
        mbar_wait(bar_q, 0);
        mbar_wait(bar_full + 0, 0);                                           // K_0 in ring slot 0
        tc_fence_after();
        issue_qk(0);
        for (int j = 0; j < nkv; ++j) {
            const bool more = (j + 1 < nkv);
            const int ik = 2 * j + 2, iv = 2 * j + 1;
            const int sk = ik % kRing, sv = iv % kRing;
            if (more) {
                mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRing) & 1u);    // K_{j+1}
                mbar_wait(bar_s_free + 8 * t, (uint32_t)j & 1u);              // both threads of every row have read S_t(j)
                tc_fence_after();
                issue_qk(sk);
            }
            mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRing) & 1u);        // V_j
            mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);                  // P_t(j) is in smem
            tc_fence_after();
            issue_pv(sv, j == 0);
        }
    } else {
        // ================= softmax: tile t, primary (hf = 0) / helper (hf = 1) thread of one query row =================
        const int sidx = warp - 3;
        const int t = sidx >> 3;
        const int hf = (sidx >> 2) & 1;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_sel, tO = tmem_base + 256 + t * 128 + lane_sel;
        const uint32_t sw = (uint32_t)row & 7u;
        uint8_t* pP = gen + (sP - base) + t * 2 * kTileBytes + row * 128;
        volatile float* mref = reinterpret_cast<volatile float*>(gen + (sMref - base)) + t * 128 + row;
        float m_ref = -INFINITY;
        const float c = a.scale_log2;
        const bool releaser = (sidx & 7) == 0;                   // first primary warp of the tile: hands ring slots back to the TMA producer
        for (int j = 0; j < nkv; ++j) {
            const int valid = a.S - j * kKV;
            const bool masked = valid < kKV;
            float mc, alpha = 1.f;
            bool rescale = false;
            if (hf == 0) {
                mbar_wait(bar_s_full + 8 * t, (uint32_t)j & 1u);
                if (releaser) {                                  // QK^T(t, j) has retired: this tile is done with K_j
                    if (elect_one()) mbar_arrive(bar_empty + 8 * ((2 * j) % kRing));
                    __syncwarp();
                }
                tc_fence_after();
                float mx = -INFINITY;
                if (!masked) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(tS, ra);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 32, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 64, ra);
                    mx = row_max_chunk(rb, mx);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(tS + 96, rb);
                    mx = row_max_chunk(ra, mx);
                    tmem_ld_wait();
                    mx = row_max_chunk(rb, mx);
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tS + ch * 32, r);
                        tmem_ld_wait();
                        mask_chunk(r, ch * 32, valid);
                        mx = row_max_chunk(r, mx);
                    }
                }
                const bool grow = (mx - m_ref) * c > kRescaleThreshold;
                const bool any_grow = __any_sync(0xffffffffu, grow) || (j == 0);
                if (any_grow) {
                    const float m_new = fmaxf(m_ref, mx);
                    alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * c);
                    m_ref = m_new;
                    rescale = (j > 0);
                }
                mc = m_ref * c;
                *mref = mc;
                tc_fence_before();
                if (t == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
                else        asm volatile("bar.arrive 2, 256;" ::: "memory");
            } else {
                if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                else        asm volatile("bar.sync 2, 256;" ::: "memory");
                tc_fence_after();
                mc = *mref;
            }
            if (j > 0) {
                // the P_t buffer and O_t belong to the tensor core until PV of tile j-1 has retired
                mbar_wait(bar_pv_done + 8 * t, (uint32_t)(j - 1) & 1u);
                if (releaser) {                                  // PV(t, j-1) has retired: this tile is done with V_{j-1}
                    if (elect_one()) mbar_arrive(bar_empty + 8 * ((2 * j - 1) % kRing));
                    __syncwarp();
                }
                tc_fence_after();
                if (rescale) {
#pragma unroll
                    for (int ch = 0; ch < kOCols / 16; ++ch) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(tO + ch * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
                        tmem_st_32x32b_x16(tO + ch * 16, r);
                    }
                    tmem_st_wait();
                }
            }
            // this thread's 64 keys: two 32-score chunks -> packed fp16 -> its half of the P tile (K-major, SWIZZLE_128B)
            {
                uint32_t ra[32], rb[32];
                uint8_t* dst = pP + hf * kTileBytes;
                tmem_ld_32x32b_x32(tS + hf * 64, ra);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(tS + hf * 64 + 32, rb);
                if (masked) mask_chunk(ra, hf * 64, valid);
                exp_chunk<POLY>(ra, 0, c, mc, dst, sw);
                tmem_ld_wait();
                tc_fence_before();                               // last read of S_t by this thread: release it for the next QK
                mbar_arrive(bar_s_free + 8 * t);
                if (masked) mask_chunk(rb, hf * 64 + 32, valid);
                exp_chunk<POLY>(rb, 1, c, mc, dst, sw);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * t);
        }
        // ---- epilogue: O / l -> fp16 -> smem (SW128, first half of this tile's P buffer) -> TMA store; 32 columns per thread
        mbar_wait(bar_pv_done + 8 * t, (uint32_t)(nkv - 1) & 1u);
        tc_fence_after();
        float inv;
        {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + 64, r);
            tmem_ld_wait();
            inv = 1.f / __uint_as_float(r[0]);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tO + (hf * 2 + ch) * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 2; ++v8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    hh[e] = __floats2half2_rn(__uint_as_float(r[v8 * 8 + 2 * e]) * inv, __uint_as_float(r[v8 * 8 + 2 * e + 1]) * inv);
                const uint32_t chunk16 = (uint32_t)((hf * 2 + ch) * 2 + v8);
                *reinterpret_cast<uint4*>(pP + ((chunk16 ^ sw) << 4)) = u;
            }
        }
        fence_proxy_async_smem();
        if (t == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
        else        asm volatile("bar.sync 4, 256;" ::: "memory");
        if (releaser && lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(sP + t * 2 * kTileBytes), "r"(head * kD), "r"(q0 + t * kQ), "r"(img) : "memory");
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

// development aid: wiw_debug_attn_trace(1) allocates a device buffer and makes the next launches stamp CTA 0's pipeline events
// into it; wiw_debug_attn_trace(0) copies the stamps out (host buffer of kTraceIters * kTraceSlots int64) -- see scripts/attn_trace.py
static long long* g_attn_trace = nullptr;
extern "C" int wiw_debug_attn_trace(int enable, long long* host_out) {
    if (enable) {
        if (!g_attn_trace) WIW_CHECK_CUDA(cudaMalloc(&g_attn_trace, sizeof(long long) * kTraceIters * kTraceSlots));
        WIW_CHECK_CUDA(cudaMemset(g_attn_trace, 0, sizeof(long long) * kTraceIters * kTraceSlots));
        return 0;
    }
    if (!g_attn_trace || !host_out) return 1;
    WIW_CHECK_CUDA(cudaMemcpy(host_out, g_attn_trace, sizeof(long long) * kTraceIters * kTraceSlots, cudaMemcpyDeviceToHost));
    WIW_CHECK_CUDA(cudaFree(g_attn_trace));
    g_attn_trace = nullptr;
    return 0;
}

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
    CUtensorMap tm_kv;
    {
        uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)S, (uint64_t)images};
        uint64_t strides[2] = {(uint64_t)3 * C * 2, (uint64_t)S * 3 * C * 2};
        uint32_t box[3] = {64, (uint32_t)kKV3, 1};
        if (encode_tmap(&tm_kv, qkv, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    AttnArgs a;
    a.S = S;
    a.nkv = (S + kKV - 1) / kKV;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.trace = g_attn_trace;
    static const int attn_dbg = getenv("WIW_ATTN_DBG") ? atoi(getenv("WIW_ATTN_DBG")) : 0;
    a.dbg = attn_dbg;
    static unsigned long long attr_set = 0;
    if (first_call_on_device(&attr_set)) {
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS2));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS2));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS3));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS3));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS4));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ts4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemTS4));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ss2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemSS2));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ss2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemSS2));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_ss3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemSS2));
    }
    // development switches for same-box A/B runs: WIW_ATTN_TS=0 -> the shared-memory-operand kernel; WIW_ATTN_POLY = exponential
    // pairs (of every 4) evaluated on the FMA pipe instead of MUFU (see exp2_poly)
    static const int ts = getenv("WIW_ATTN_TS") ? atoi(getenv("WIW_ATTN_TS")) : 6;        // 0 = SS kernel, 1 = TS 8 softmax warps, 2 = TS 16, 3 = TS 16 + double-buffered 64-key scores
    static const int poly = getenv("WIW_ATTN_POLY") ? atoi(getenv("WIW_ATTN_POLY")) : 1;
    dim3 grid((S + 2 * kQ - 1) / (2 * kQ), C / 64, images);
    if (ts == 6) {
        attn_spatial_ss3_kernel<1><<<grid, kAttnThreads3, kAttnSmemSS2, stream>>>(tm_qkv, tm_out, a, C);
    } else if (ts == 5) {
        if (poly == 2) attn_spatial_ss2_kernel<2><<<grid, kAttnThreads2, kAttnSmemSS2, stream>>>(tm_qkv, tm_out, a, C);
        else attn_spatial_ss2_kernel<1><<<grid, kAttnThreads2, kAttnSmemSS2, stream>>>(tm_qkv, tm_out, a, C);
    } else if (ts == 4) {
        dim3 grid4((S + kQ - 1) / kQ, C / 64, images);
        if (poly == 2) attn_spatial_ts4_kernel<2><<<grid4, kAttnThreads4, kAttnSmemTS4, stream>>>(tm_qkv, tm_out, a, C);
        else attn_spatial_ts4_kernel<1><<<grid4, kAttnThreads4, kAttnSmemTS4, stream>>>(tm_qkv, tm_out, a, C);
    } else if (ts == 3) {
        a.nkv = (S + kKV3 - 1) / kKV3;
        if (poly == 2) attn_spatial_ts3_kernel<2><<<grid, kAttnThreads2, kAttnSmemTS3, stream>>>(tm_qkv, tm_kv, tm_out, a, C);
        else attn_spatial_ts3_kernel<1><<<grid, kAttnThreads2, kAttnSmemTS3, stream>>>(tm_qkv, tm_kv, tm_out, a, C);
    } else if (!ts) attn_spatial_kernel<1><<<grid, kAttnThreads, kAttnSmem, stream>>>(tm_qkv, tm_out, a, C);
    else if (ts == 2 && poly == 2) attn_spatial_ts2_kernel<2><<<grid, kAttnThreads2, kAttnSmemTS2, stream>>>(tm_qkv, tm_out, a, C);
    else if (ts == 2) attn_spatial_ts2_kernel<1><<<grid, kAttnThreads2, kAttnSmemTS2, stream>>>(tm_qkv, tm_out, a, C);
    else if (poly == 2) attn_spatial_ts_kernel<2><<<grid, kAttnThreads, kAttnSmemTS, stream>>>(tm_qkv, tm_out, a, C);
    else attn_spatial_ts_kernel<1><<<grid, kAttnThreads, kAttnSmemTS, stream>>>(tm_qkv, tm_out, a, C);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
