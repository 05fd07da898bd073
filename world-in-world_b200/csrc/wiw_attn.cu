// Spatial self-attention for the SVD transformer blocks on sm_100a:
//   out = softmax(Q K^T / sqrt(64)) V   per (image, head), head_dim 64, no mask, S up to 9216
// (reference: F.scaled_dot_product_attention in AttnProcessor2_0, DP/models/attention_processor.py:2383,
//  called from BasicTransformerBlock.attn1, DP/models/attention.py:507).
//
// One CTA (one per SM) = 256 query rows (two 128-row tiles) of one (image, head); the two tiles
// ping-pong so the tensor core works on one while the other is in softmax.  Per KV tile of 128 keys:
//   warp 0     TMA: Q0,Q1 once, then K_j / V_j tiles ([128][64] fp16, SWIZZLE_128B) through a 4-slot ring
//   warps 1,2  tcgen05.mma issuers, ONE PER Q TILE (each tile's QK^T / PV chain waits only on its own softmax warps):
//              S_t = Q_t K_j^T (M128 N128 K64, K-major x K-major) -> TMEM,
//              O_t += P_t V_j (M128 N64 K128, A = P_t from smem, B = V consumed MN-major from its TMA tile)
//   warps 3-6  softmax of tile 0, warps 7-10 softmax of tile 1 (one thread per query row):
//              ONE tcgen05.ld sweep: the row's 128 scores stay in registers (max, then exp2 from registers) and S_t goes back to the
//              tensor core immediately, so QK^T of the next KV tile overlaps this tile's exponentials (+2.8 % at S = 9216, +5.6 % at
//              S = 576 over the two-sweep version, profiles/r02_perf_attn_onepass_ab.txt); P -> fp16 -> swizzled smem, lazy rescale
//              of O in TMEM when some row max of the warp grew by more than 2^8.  The row sum l is NOT accumulated by the
//              softmax threads: the PV MMA runs with N = 80 where B columns 64..79 come from a tile of ones, so
//              O[:, 64] = sum_k P[:, k] in fp32 for free.  Final O / l -> fp16 -> TMA store.
// This kernel does not use setmaxnreg, so the mbarrier-timeout diagnostic may stay a real (noinline, printf) call: with the inline trap
// the softmax loop scheduled measurably worse (profiles/r02_attn_trap_call_ab.txt).
#define WIW_TRAP_PRINTF 1
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

constexpr int kQ = 128, kKV = 128, kD = 64;
constexpr int kTileBytes = 128 * 64 * 2;        // 16 KB: a Q / K / V tile and each 64-key half of P
constexpr int kRing = 4;
constexpr int kAttnThreads = 11 * 32;
constexpr int kAttnSmem = 1024 + kTileBytes * (2 + kRing + 1 + 4) + 256;   // Q0,Q1 | K/V ring | ones | P0,P1
constexpr uint32_t kTmemCols = 512;              // S0 [0,128) S1 [128,256) O0 [256,336) O1 [384,464)
constexpr int kOCols = 80;                       // 64 output columns + 16 copies of the row sum l (P @ ones)
constexpr int kStaggerCycles = 1000;
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
};

// keys at or beyond `valid` (zero-filled by TMA past the image's S rows) -> -inf before max / exp
__device__ __forceinline__ void mask_chunk(uint32_t* r, int base, int valid) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
        if (base + e >= valid) r[e] = 0xff800000u;
}

// The softmax warps are bound by instruction issue inside their dependency chains, so the per-score work uses the
// instructions that do two things at once: 3-input max (FMNMX3) for the row maximum, packed f32x2 FMA / ADD (FFMA2, FADD2: the pipe
// time of two scalar instructions, the issue cost of one) for the scaling and the polynomial.  Same operations, same rounding: the
// results are bit-identical to the scalar forms.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

__device__ __forceinline__ float row_max_chunk(const uint32_t* r, float mx) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) mx = max3(mx, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
    return mx;
}

// exp2_poly on a pair (same clamp, split and polynomial; the exponent insertion stays scalar integer work)
__device__ __forceinline__ void exp2_poly2(f32x2 x, float& y0, float& y1) {
    float x0, x1;
    upk2(x, x0, x1);
    x = pk2(fmaxf(x0, -26.0f), fmaxf(x1, -26.0f));
    const f32x2 xr = add2(x, pk2(12582912.0f, 12582912.0f));
    const f32x2 f = fma2(add2(xr, pk2(-12582912.0f, -12582912.0f)), pk2(-1.0f, -1.0f), x);          // x - (xr - 1.5 * 2^23), in [-0.5, 0.5]
    f32x2 p = fma2(pk2(0.05517167f, 0.05517167f), f, pk2(0.24261113f, 0.24261113f));
    p = fma2(p, f, pk2(0.69326097f, 0.69326097f));
    p = fma2(p, f, pk2(0.99992806f, 0.99992806f));
    float p0, p1, r0, r1;
    upk2(p, p0, p1);
    upk2(xr, r0, r1);
    y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
    y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// 32 scores -> 32 fp16 probabilities p = 2^(s*c - mc) written to the swizzled P tile
// (fp32 MUFU.EX2 + one F2FP pack per pair: cheaper than ex2.f16x2, which splits into two MUFU + a PRMT)
template <int kPolyPairs>
__device__ __forceinline__ void exp_chunk(const uint32_t* r, int ch, float c, float mc, uint8_t* dst, uint32_t sw) {
    const f32x2 c2 = pk2(c, c), nmc2 = pk2(-mc, -mc);
#pragma unroll
    for (int v8 = 0; v8 < 4; ++v8) {
        uint32_t p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const f32x2 a = fma2(pk2u(r[v8 * 8 + 2 * e], r[v8 * 8 + 2 * e + 1]), c2, nmc2);      // s * c - m * c, two scores per instruction
            float y0, y1;
            if (e < kPolyPairs) exp2_poly2(a, y0, y1);
            else {
                float a0, a1;
                upk2(a, a0, a1);
                y0 = fast_exp2(a0); y1 = fast_exp2(a1);
            }
            p[e] = cvt_f16x2(y0, y1);
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
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2); }
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
                umma_commit(bar_empty + 8 * stage);
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
                umma_commit(bar_empty + 8 * stage);
            }
            __syncwarp();
        };
        mbar_wait(bar_q, 0);
        mbar_wait(bar_full + 0, 0);
        // The two tiles' chains are independent; started together they run in lockstep and leave the MUFU / issue pipes idle
        // during both tiles' load-max-wait phases.  Tile 1 starts ~a third of a KV-tile period late, which the chains then keep
        // (same-box A/B: 820 TFLOP/s without the offset, 851 with 900-1700 cycles; profiles/r02_perf_attn_mma2_stagger_ab.txt)
        if (t == 1) { const long long t0 = clock64(); while (clock64() - t0 < kStaggerCycles) {} }
        tc_fence_after();
        issue_qk(0);
        for (int j = 0; j < nkv; ++j) {
            const bool more = (j + 1 < nkv);
            const int ik = 2 * j + 2, iv = 2 * j + 1;
            const int sk = ik % kRing, sv = iv % kRing;
            if (more) {
                mbar_wait(bar_full + 8 * sk, (uint32_t)(ik / kRing) & 1u);
                mbar_wait(bar_s_free + 8 * t, (uint32_t)j & 1u);
                tc_fence_after();
                issue_qk(sk);
            }
            mbar_wait(bar_full + 8 * sv, (uint32_t)(iv / kRing) & 1u);
            mbar_wait(bar_p_full + 8 * t, (uint32_t)j & 1u);
            tc_fence_after();
            issue_pv(sv, j == 0);
        }
    } else {
        // ================= softmax / correction / epilogue: one thread per query row of tile t =================
        const int t = (warp - 3) >> 2;
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
            // ONE sweep: the 128 scores of this row go to registers and stay there; S_t is handed back to the tensor core at once,
            // so QK^T of the next KV tile runs under this tile's exponentials (the registers are the second score buffer)
            uint32_t sc[128];
            tmem_ld_32x32b_x32(tS, sc);
            tmem_ld_32x32b_x32(tS + 32, sc + 32);
            tmem_ld_32x32b_x32(tS + 64, sc + 64);
            tmem_ld_32x32b_x32(tS + 96, sc + 96);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_s_free + 8 * t);
            if (masked) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) mask_chunk(sc + 32 * ch, ch * 32, valid);
            }
            float mx = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) mx = row_max_chunk(sc + 32 * ch, mx);
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
            // p = 2^((s - m_ref) * c) in packed fp16 -> P smem (K-major, SWIZZLE_128B, two 64-key halves)
            const float mc = m_ref * c;
            exp_chunk<POLY>(sc, 0, c, mc, pP, sw);
            exp_chunk<POLY>(sc + 32, 1, c, mc, pP, sw);
            exp_chunk<POLY>(sc + 64, 2, c, mc, pP + kTileBytes, sw);
            exp_chunk<POLY>(sc + 96, 3, c, mc, pP + kTileBytes, sw);
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
        if (((warp - 3) & 3) == 0 && lane == 0) {
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
    static unsigned long long attr_set = 0;
    if (first_call_on_device(&attr_set)) {
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        WIW_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    }
    static const int poly = getenv("WIW_ATTN_POLY") ? atoi(getenv("WIW_ATTN_POLY")) : 1;      // development aid for A/B runs
    // 1 of every 4 exponential pairs runs on the FMA pipe (exp2_poly2): with the packed f32x2 softmax 898 TFLOP/s at S = 9216 against 782
    // all-MUFU and 755 with 2 of 4 (profiles/r02_attn_poly_share_packed_ab.txt).  Round 2 rebuilt this kernel six ways (A operands in tensor memory, P aliased on S,
    // two / four threads per row, double-buffered scores, one MMA warp per Q tile): all are correct, none beats this one by more
    // than 2 % -- DESIGN.md section 4 and profiles/r02_attn_*.txt; the variants live in the git history (commits c837173..33a0137).
    dim3 grid((S + 2 * kQ - 1) / (2 * kQ), C / 64, images);
    if (poly == 0) attn_spatial_kernel<0><<<grid, kAttnThreads, kAttnSmem, stream>>>(tm_qkv, tm_out, a, C);
    else if (poly == 2) attn_spatial_kernel<2><<<grid, kAttnThreads, kAttnSmem, stream>>>(tm_qkv, tm_out, a, C);
    else attn_spatial_kernel<1><<<grid, kAttnThreads, kAttnSmem, stream>>>(tm_qkv, tm_out, a, C);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
